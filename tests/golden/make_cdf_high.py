"""Golden high-precision CDF tables (tests/golden/cdf_high_v1.npz).

gauss_cdf_create_high_precision (src/utils/sampling/gaussian_cdf.c:192-318) runs on the reference's sc_mpf layer, which
in a build with MPFR maps every call to the mpfr_* function with MPFR_RNDZ at `precision` bits (src/utils/arith/
sc_mpf.c:38, :42-54); the build WITHOUT MPFR -- the only one possible in this container -- has empty bodies for
sc_mpf_exp / sc_mpf_get_pi (:99-106, :822-828) and yields a degenerate table, so the compiled reference cannot pin this
function here.  This script restates the function's operation sequence over mpmath's correctly rounded arithmetic
(libmp, rounding 'd' = toward zero; exp and pi are evaluated with 80 guard bits and then truncated), i.e. it produces
what the MPFR build of the reference produces.  The product's own construction (libsafecrypto_b200/csrc/cdf_hp.cu, plain
integer arithmetic, no mpmath) is compared with these tables bit for bit in tests/test_cdf_high.py.

usage: python tests/golden/make_cdf_high.py            (needs mpmath; writes tests/golden/cdf_high_v1.npz)"""
import hashlib
import os
import struct

import numpy as np
from mpmath import libmp

RZ = "d"


def from_double(x):
    return libmp.from_float(float(x))              # exact


def trunc(x, prec):
    return libmp.mpf_pos(x, prec, RZ)


def table(precision, blinding, tail, sigma):
    """(entries, words) uint64, word 0 least significant -- the reference's cdf_128 / cdf_192 / cdf_256 array."""
    p = precision
    nw = p // 64
    guard = p + 80
    sigma = struct.unpack("f", struct.pack("f", sigma))[0]          # FLOAT is float in the reference's default build
    tail = struct.unpack("f", struct.pack("f", tail))[0]
    x = int(np.float32(tail) * np.float32(sigma))                   # sc_ceil_log2((size_t)(tail * sigma)), gaussian_cdf.c:340
    bits = x.bit_length() - 1 + (1 if x & (x - 1) else 0)
    size = 1 << bits
    pi = trunc(libmp.mpf_pi(guard), p)                              # mpfr_const_pi(RNDZ)
    t1 = libmp.mpf_shift(pi, 1)                                     # mul_2exp: exact
    t0 = libmp.mpf_sqrt(t1, p, RZ)
    two_sqrt_2pi = libmp.mpf_div(from_double(2.0), t0, p, RZ)
    sqrt_1_2 = libmp.mpf_sqrt(from_double(0.5), p, RZ)
    s128 = from_double(sigma)
    half = from_double(0.5)
    if blinding:
        s128 = libmp.mpf_mul(s128, sqrt_1_2, p, RZ)
    t0 = libmp.mpf_shift(from_double(1.0), p)                       # 2^precision
    t1 = libmp.mpf_div(t0, s128, p, RZ)
    d = libmp.mpf_mul(t1, two_sqrt_2pi, p, RZ)
    t0 = libmp.mpf_mul(s128, s128, p, RZ)
    e = libmp.mpf_neg(libmp.mpf_div(half, t0, p, RZ))
    s = libmp.mpf_mul(d, half, p, RZ)
    out = np.zeros((size, nw), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i in range(1, size - 1):
        ip = libmp.to_int(s, RZ)                                    # the limb-by-limb mpfr_get_ui(RNDZ) extraction
        assert 0 <= ip < (1 << p)
        for j in range(nw):
            out[i, j] = (ip >> (64 * j)) & mask
        t0 = libmp.mpf_mul(e, libmp.from_int(i * i), p, RZ)
        t1 = trunc(libmp.mpf_exp(t0, guard, "n"), p)
        t0 = libmp.mpf_mul(d, t1, p, RZ)
        s = libmp.mpf_add(s, t0, p, RZ)
    out[size - 1, :] = np.uint64(mask)
    return out


CASES = [  # name, precision, blinding, tail, sigma
    ("p128_s215", 128, 0, 13.42, 215.0),           # BLISS-B I (func_alg_bliss_b.c:118 asks for the 128-bit CDF)
    ("p128_s215_blind", 128, 1, 13.42, 215.0),
    ("p192_s19", 192, 0, 10.0, 19.53),
    ("p256_s3", 256, 0, 9.42, 3.33),
    ("p128_s4", 128, 0, 12.0, 4.5),
]

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    data = {}
    for name, prec, blind, tail, sigma in CASES:
        t = table(prec, blind, tail, sigma)
        data[name + "_params"] = np.array([prec, blind, tail, sigma], dtype=np.float64)
        if t.shape[0] > 512:                        # large tables: digest + every 37th row
            data[name + "_sha256"] = np.frombuffer(hashlib.sha256(t.tobytes()).digest(), dtype=np.uint8)
            data[name + "_rows"] = np.arange(0, t.shape[0], 37, dtype=np.int64)
            data[name + "_sample"] = t[::37].copy()
            data[name + "_shape"] = np.array(t.shape, dtype=np.int64)
        else:
            data[name] = t
        print(name, t.shape, hex(int(t[1, -1])), hex(int(t[t.shape[0] // 2, -1])))
    np.savez_compressed(os.path.join(here, "cdf_high_v1.npz"), **data)
