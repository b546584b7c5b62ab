"""A/B of the mat-vec kernels: 32-bit stash (k_matvec_w32, SCGPU_MATVEC16_MINL=99) against the 16-bit stash
(k_matvec16_w32), Kyber shapes, outputs compared.  usage: python tools/matvec_stash_ab.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc, _oracle as O
dev = torch.device("cuda", 0); g = torch.Generator(device=dev).manual_seed(1)
q, n = 7681, 256
w, r = O.tables(q, n, 16); pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e-3
for k, l in ((2, 2), (2, 3), (3, 2), (1, 1), (3, 3)):
    B = 1 << 17
    A = torch.randint(0, q, (B, k * l, n), dtype=torch.int32, device=dev, generator=g)
    s = torch.randint(-4, 5, (B, l, n), dtype=torch.int32, device=dev, generator=g)
    o = torch.empty((B, k, n), dtype=torch.int32, device=dev)
    res = {}
    for minl in ("99", "1"):
        os.environ["SCGPU_MATVEC16_MINL"] = minl
        t = timeit(lambda: pl.matvec(o, A, s, k, l)); res[minl] = (t, o.clone())
    by = 4 * n * (k * l + l + k)
    print("k=%d l=%d  stash32 %.4g/s (%.3f)  stash16 %.4g/s (%.3f)  equal=%s" % (k, l, B / res["99"][0], B / res["99"][0] * by / 6.55e12, B / res["1"][0], B / res["1"][0] * by / 6.55e12, torch.equal(res["99"][1], res["1"][1])))
