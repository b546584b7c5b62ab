"""The *_host_multi entry points: one host batch split over the devices of this process with a host-side gather.
On a one-GPU box this exercises the slab logic with the single device (and with the same device given twice)."""
import ctypes

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import libsafecrypto_b200 as sc  # noqa: E402
from libsafecrypto_b200 import binding as B  # noqa: E402


def test_multi_device_host_batches_match_the_oracle():
    q, n = 12289, 512
    w, r = O.tables(q, n, 16)
    ps = B.NttPlanSet(n, q, sc.REFERENCE, w, r)
    assert ps.ndev == torch.cuda.device_count()
    rng = np.random.default_rng(9)
    count = 3001                                     # ragged: slabs of unequal size
    a = rng.integers(0, q, size=(count, n)).astype(np.int32)
    b = rng.integers(0, q, size=(count, n)).astype(np.int32)
    exp = O.port().ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, 16, a, b, w, r)
    out = np.zeros_like(a)
    ps.polymul_host(out, a, b)
    assert np.array_equal(out, exp)
    # exact op through the same scatter / gather, shared second operand
    key = rng.integers(0, q, size=n).astype(np.int16)
    out2 = np.zeros_like(a)
    ps.batch_host(sc.OP_TRIPLE16, out2, a, key)
    assert np.array_equal(out2, O.port().ntt_batch(O.REFERENCE, O.OP_TRIPLE16, n, q, 16, a, key, w, r))
    # more slabs than devices: the same device serves several slabs concurrently (its plan serialises them)
    handles = (ctypes.c_void_p * 3)(ps.handles[0], ps.handles[0], ps.handles[ps.ndev - 1])
    out3 = np.zeros_like(a)
    st = sc.lib().scgpu_polymul_batch_host_multi(handles, 3, out3.ctypes.data, a.ctypes.data, b.ctypes.data, n, count)
    assert st == 0 and np.array_equal(out3, exp)
    # count smaller than the number of slabs
    out4 = np.zeros((2, n), dtype=np.int32)
    st = sc.lib().scgpu_polymul_batch_host_multi(handles, 3, out4.ctypes.data, a.ctypes.data, b.ctypes.data, n, 2)
    assert st == 0 and np.array_equal(out4, exp[:2])
