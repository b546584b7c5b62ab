"""A/B of the work distribution of the persistent warp-local kernels: static stride (SCGPU_STATIC_SCHED=1) against
the global work counter (default), same inputs, outputs compared.  usage: python tools/ab_sched.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e-3


def ab(name, units, fn, out):
    res = {}
    for mode in ("1", "0"):
        os.environ["SCGPU_STATIC_SCHED"] = mode
        out.zero_()
        t = timeit(fn)
        res[mode] = (t, out.clone())
    os.environ["SCGPU_STATIC_SCHED"] = "0"
    same = torch.equal(res["1"][1], res["0"][1])
    print("%-40s static %.4g/s  dynamic %.4g/s  (%+.1f %%)  outputs %s" % (
        name, units / res["1"][0], units / res["0"][0], (res["1"][0] / res["0"][0] - 1) * 100, "equal" if same else "DIFFER"))
    return same


def rnd(q, shape):
    return torch.randint(0, q, shape, dtype=torch.int32, device=dev, generator=g)


ok = True
for n, q, tw in ((512, 12289, 16), (1024, 12289, 16), (256, 7681, 16), (256, 8380417, 32)):
    B = (1 << 29) // (4 * n)
    w, r = O.tables(q, n, tw)
    pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    a, b, o = rnd(q, (B, n)), rnd(q, (B, n)), torch.empty((B, n), dtype=torch.int32, device=dev)
    ok &= ab("polymul n=%d q=%d" % (n, q), B, lambda: pl.polymul(o, a, b), o)
    ok &= ab("canonical fwd n=%d q=%d" % (n, q), B, lambda: pl.ntt_canonical(o, a), o)
    ok &= ab("canonical inv n=%d q=%d" % (n, q), B, lambda: pl.ntt_canonical(o, a, inverse=True), o)
    if tw == 16:
        key = rnd(q, (n,)).to(torch.int16)
        ok &= ab("key product n=%d" % n, B, lambda: pl.mul_key(o, a, key), o)
    # an odd, unaligned-count batch exercises the tail
    ok &= ab("polymul n=%d ragged (100003 rows)" % n, 100003, lambda: pl.polymul(o[:100003], a[:100003], b[:100003]), o[:100003])
    if n == 256:
        for k, l in ((2, 2), (3, 3), (4, 4)) if q == 7681 else ((5, 4),):
            Bm = 1 << 16
            A, s = rnd(q, (Bm, k * l, n)), torch.randint(-4, 5, (Bm, l, n), dtype=torch.int32, device=dev, generator=g)
            om = torch.empty((Bm, k, n), dtype=torch.int32, device=dev)
            ok &= ab("mat-vec q=%d k=%d l=%d" % (q, k, l), Bm, lambda: pl.matvec(om, A, s, k, l), om)
            del A, s, om
    del a, b, o
ns = 1 << 18
seeds = torch.randint(0, 256, (ns, 40), dtype=torch.uint8, device=dev, generator=g)
smp = torch.empty((ns, 512), dtype=torch.int32, device=dev)
for prec in (64, 32):
    gp = sc.GaussPlan(sc.SAMPLER_CDF, prec, 0, 13.42, 215.0)
    ok &= ab("CDF-%d ChaCha20 samples" % prec, ns * 512, lambda: gp.streams(sc.PRNG_CHACHA, seeds, 512, smp), smp)
sys.exit(0 if ok else 1)
