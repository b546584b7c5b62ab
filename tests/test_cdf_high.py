"""High-precision CDF table construction (gauss_cdf_create_high_precision, gaussian_cdf.c:192-318) -- host set-up of
the product (libsafecrypto_b200/csrc/cdf_hp.cu), no GPU involved.

The reference runs this on MPFR (every operation truncated to `precision` bits, sc_mpf.c:38); the reference build that is
possible in this container has no MPFR and yields a degenerate table, so the pin is an INDEPENDENT evaluation of the same
operation sequence over mpmath's correctly rounded arithmetic (tests/golden/make_cdf_high.py -> cdf_high_v1.npz), plus
two cross-checks against quantities that ARE pinned to the compiled reference: the 64-bit table of gauss_cdf_create_64
(long double path) and an exact (non-truncating) evaluation of the formula."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

import _oracle as O
import libsafecrypto_b200 as sc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cdf_high_v1.npz"))
CASES = ["p128_s215", "p128_s215_blind", "p192_s19", "p256_s3", "p128_s4"]


def build(precision, blinding, tail, sigma):
    L = sc.lib()
    n = ctypes.c_size_t(0)
    assert L.scgpu_gauss_cdf_table_high(None, 0, ctypes.byref(n), precision, blinding, tail, sigma) == 0
    t = np.zeros((n.value, precision // 64), dtype=np.uint64)
    assert L.scgpu_gauss_cdf_table_high(t.ctypes.data, n.value, ctypes.byref(n), precision, blinding, tail, sigma) == 0
    return t


def as_int(row):
    return sum(int(w) << (64 * j) for j, w in enumerate(row))


@pytest.mark.parametrize("name", CASES)
def test_table_equals_the_truncating_evaluation(name):
    prec, blind, tail, sigma = G[name + "_params"]
    t = build(int(prec), int(blind), float(tail), float(sigma))
    if name in G.files:
        assert np.array_equal(t, G[name])
    else:
        assert list(t.shape) == list(G[name + "_shape"])
        assert np.array_equal(t[G[name + "_rows"]], G[name + "_sample"])
        assert hashlib.sha256(t.tobytes()).digest() == bytes(G[name + "_sha256"])
    # shape of a CDF table: zero first, all-ones last, non-decreasing in between
    assert not t[0].any() and (t[-1] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    vals = [as_int(r) for r in t]
    assert all(a <= b for a, b in zip(vals, vals[1:]))


def test_top_words_agree_with_the_pinned_64bit_table():
    """gauss_cdf_create_64 (long double arithmetic, pinned against the compiled reference in test_oracle_vs_ref.py)
    computes the same sums at 64 bits: the top word of the 128-bit table must agree to within the long double path's
    accumulated rounding (a few thousand units in the last place over 2885 additions)."""
    t128 = build(128, 0, 13.42, 215.0)
    t64 = np.asarray(O.port().cdf_table(64, 0, 13.42, 215.0)).astype(np.uint64)
    assert t64.shape[0] == t128.shape[0]
    live = (t64 != np.uint64(0xFFFFFFFFFFFFFFFF)) & (t64 != 0)
    assert live.sum() > 2000
    diff = np.abs(t128[live, 1].astype(np.float64) - t64[live].astype(np.float64))
    assert diff.max() < 2.0 ** 20, diff.max()                      # of 2^64: agreement to ~2^-44 relative


@pytest.mark.parametrize("precision,blinding,tail,sigma", [(128, 0, 13.42, 215.0), (192, 1, 10.0, 19.53), (256, 0, 9.42, 3.33)])
def test_truncation_error_against_the_exact_formula(precision, blinding, tail, sigma):
    """The exact evaluation of the same recurrence (tests/_oracle.py: high_precision_cdf_table, decimal arithmetic with
    40 guard digits) bounds the effect of the per-operation truncations: every truncation loses less than one unit of
    the precision-bit mantissa, there are about six per entry, and they all point the same way."""
    t = build(precision, blinding, tail, sigma)
    ex = O.high_precision_cdf_table(precision, tail, sigma, blinding)
    assert t.shape == ex.shape
    worst = 0
    for i in range(1, t.shape[0] - 1):
        a, b = as_int(t[i]), as_int(ex[i])
        if b == (1 << precision) - 1:
            continue
        assert a <= b + 1                                            # truncation never pushes a value up
        worst = max(worst, b - a)
    assert worst < 64 * t.shape[0], worst                            # units of 2^0 on a 2^precision scale


def test_arguments():
    L = sc.lib()
    n = ctypes.c_size_t(0)
    assert L.scgpu_gauss_cdf_table_high(None, 0, ctypes.byref(n), 64, 0, 13.42, 215.0) < 0
    assert L.scgpu_gauss_cdf_table_high(None, 0, ctypes.byref(n), 128, 0, 13.42, -1.0) < 0
    assert L.scgpu_gauss_cdf_table_high(None, 0, None, 128, 0, 13.42, 215.0) < 0
    assert L.scgpu_gauss_cdf_table_high(None, 0, ctypes.byref(n), 128, 0, 13.42, 215.0) == 0 and n.value == 4096
    small = np.zeros((8, 2), dtype=np.uint64)
    assert L.scgpu_gauss_cdf_table_high(small.ctypes.data, 8, ctypes.byref(n), 128, 0, 13.42, 215.0) < 0
