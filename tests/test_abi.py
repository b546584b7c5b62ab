"""CPU-only checks of the boundary: libscgpu.so loads, exports every symbol include/*.h declares, and its
struct layouts match the compiled reference.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    """Function prototypes at file scope: `type name(args);` -- function-pointer members are skipped."""
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = set()
    for m in re.finditer(r"^(?:extern\s+)?(?:const\s+)?[A-Za-z_]\w*(?:\s+|\s*\*+\s*)+([A-Za-z_]\w*)\s*\(", text, flags=re.M):
        if "(*" not in m.group(0) and not m.group(0).lstrip().startswith("typedef"):
            names.add(m.group(1))
    return names


@pytest.fixture(scope="module")
def lib():
    import libsafecrypto_b200 as sc
    if not os.path.exists(sc.lib_path()):
        import __graft_entry__ as ge
        ge.build()
    return sc.lib()


def test_header_symbols_exported(lib):
    for header in ("scgpu.h", "scgpu_dropin.h"):
        names = declared_symbols(header)
        assert len(names) >= 15, header
        for name in sorted(names):
            assert hasattr(lib, name), "%s declares %s but libscgpu.so does not export it" % (header, name)
    for must in ("scgpu_polymul_batch", "scgpu_ntt_batch", "scgpu_gauss_streams", "utils_arith_ntt", "init_reduce",
                 "create_sampler", "get_vector_32", "prng_create", "scgpu_matvec_batch", "roots_of_unity_s16"):
        assert must in declared_symbols("scgpu.h") | declared_symbols("scgpu_dropin.h")


def test_init_reduce_and_layout(lib):
    import libsafecrypto_b200 as sc
    buf = sc.make_params(512, 12289)
    raw = bytes(buf.raw)
    import struct
    q_dbl, inv_q_dbl, inv_q_flt = struct.unpack_from("<ddf", raw, 0)
    n, = struct.unpack_from("<Q", raw, 20)
    q, q_inv, m, k = struct.unpack_from("<iIii", raw, 28)
    assert (q_dbl, n, q, k) == (12289.0, 512, 12289, 30) and m == (1 << 30) // 12289 == 87374
    assert inv_q_dbl == 1.0 / 12289.0
    if O.ref_available():
        R = O.ref().lib
        R.ref_sizeof_ntt_params.restype = ctypes.c_size_t
        R.ref_sizeof_ntt_table.restype = ctypes.c_size_t
        assert R.ref_sizeof_ntt_params() == sc.binding.PARAMS_SIZE == 60
        assert R.ref_sizeof_ntt_table() == 77 * 8
        rbuf = ctypes.create_string_buffer(64)
        R.init_reduce(rbuf, ctypes.c_size_t(512), 12289)
        assert bytes(rbuf.raw)[:60] == raw[:60]


def test_table_has_77_members_and_roots(lib):
    lib.utils_arith_ntt.restype = ctypes.POINTER(ctypes.c_void_p * 77)
    for variant in range(6):
        tab = lib.utils_arith_ntt(variant).contents
        assert all(tab[i] for i in range(77))
    # distinct tables per variant, unknown type -> reference table (arith.c:360-396)
    addr = lambda v: ctypes.addressof(lib.utils_arith_ntt(v).contents)  # noqa: E731
    assert len({addr(v) for v in range(6)}) == 6 and addr(11) == addr(0)
    # run-time twiddle generation reproduces the reference's generated tables
    import numpy as np
    lib.roots_of_unity_s16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int32]
    lib.roots_of_unity_s32.argtypes = lib.roots_of_unity_s16.argtypes
    for tw, q, n in O.TABLE_PARAMS:
        dt = np.int16 if tw == 16 else np.int32
        w = np.zeros(n, dtype=dt)
        r = np.zeros(n, dtype=dt)
        fn = lib.roots_of_unity_s16 if tw == 16 else lib.roots_of_unity_s32
        assert fn(w.ctypes.data, r.ctypes.data, n, q, 0, 0) == 0
        ew, er = O.tables(q, n, tw)
        assert np.array_equal(w, ew) and np.array_equal(r, er)


def test_no_gpu_means_loud_failure(lib):
    """Without a device the product must refuse, not fall back."""
    import libsafecrypto_b200 as sc
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    w, r = O.tables(12289, 512, 16)
    with pytest.raises(sc.ScgpuError):
        sc.NttPlan(512, 12289, sc.REFERENCE, w, r)
    with pytest.raises(sc.ScgpuError):
        sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 215.0)


def test_product_does_not_touch_oracle():
    """Nothing under libsafecrypto_b200/ may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "libsafecrypto_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("libscoracle", "libscref", "sc_oracle.h", "_oracle", "orc_"):
                    if needle == "_oracle" and f.endswith((".cu", ".cuh", ".h")):
                        continue
                    assert needle not in text.replace("oracle/sc_oracle_ntt.c lane_quotient", ""), (f, needle)
