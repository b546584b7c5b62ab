"""Latency of single drop-in calls (one polynomial / one vector per call, host buffers): what a scheme that was only
re-linked against libscgpu.so pays per call.  usage: python tools/dropin_latency.py"""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O

L = sc.lib()
L.prng_create.restype = ctypes.c_void_p
L.prng_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
L.prng_set_entropy.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
L.prng_init.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
L.prng_destroy.argtypes = [ctypes.c_void_p]
L.create_sampler.restype = ctypes.c_void_p
L.create_sampler.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int32, ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_float]
L.get_vector_32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float]
seed = bytes(((i * 7 + 3) & 0xFF) for i in range(64))
for pname, prng in (("aes_ctr_drbg", O.PRNG_AES_CTR_DRBG), ("chacha20", O.PRNG_CHACHA)):
    ctx = L.prng_create(5, prng, 0, 0x00100000)
    L.prng_set_entropy(ctx, seed, len(seed)); L.prng_init(ctx, b"SAFEcrypto nonce", 16)
    for stype, sname in ((0, "cdf64"), (1, "knuth_yao64"), (5, "bernoulli64")):
        smp = L.create_sampler(stype, 64, 0, 512, 0, ctx, 13.42, 215.0)
        v = np.zeros(512, dtype=np.int32)
        for _ in range(3):
            L.get_vector_32(smp, v.ctypes.data, 512, 0.0)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            L.get_vector_32(smp, v.ctypes.data, 512, 0.0)
        dt = (time.perf_counter() - t0) / reps
        print("get_vector_32(512) %-12s %-12s %8.1f us per call  (%.3g samples/s)" % (sname, pname, dt * 1e6, 512 / dt))
    L.prng_destroy(ctx)

# one polynomial per call through the batch entry point with host buffers (what a table member does: plan lookup, H2D,
# one launch, D2H, stream synchronisation)
for q, n, tw in ((12289, 512, 16), (8380417, 256, 32)):
    w, r = O.tables(q, n, tw)
    a = np.random.default_rng(1).integers(0, q, size=(1, n)).astype(np.int32)
    b = np.random.default_rng(2).integers(0, q, size=(1, n)).astype(np.int32)
    out = np.zeros((1, n), dtype=np.int32)
    for vname, v in (("reference", sc.REFERENCE), ("avx", sc.AVX)):
        pl = sc.NttPlan(n, q, v, w, r)
        for label, fn in (("fwd_ntt (exact)", lambda: pl.batch_host(sc.OP_FWD, out, a)),
                          ("polymul (fused)", lambda: pl.polymul_host(out, a, b))):
            for _ in range(5):
                fn()
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
            dt = (time.perf_counter() - t0) / 200
            print("count=1 %-16s n=%d q=%d %-9s %7.1f us per call" % (label, n, q, vname, dt * 1e6))
