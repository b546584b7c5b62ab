"""A/B of the fused two-operand product: full-length transforms + pointwise products (SCGPU_NO_BASEMUL=1 at plan
creation) against transforms that stop two stages early + degree-3 base multiplication (default), and the latter
without the range vote (SCGPU_PLAN_INPUTS_IN_RANGE).  Same inputs, outputs compared.
usage: python tools/ab_basemul.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
PEAK = 6550.4e9


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


ok = True
for n, q in ((512, 12289), (1024, 12289), (256, 7681)):
    B = (1 << 31) // (4 * n)
    w, r = O.tables(q, n, 16)
    os.environ["SCGPU_NO_BASEMUL"] = "1"
    p_full = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    del os.environ["SCGPU_NO_BASEMUL"]
    p_bm = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    p_flag = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    p_flag.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    a = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
    b = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
    outs = []
    line = "polymul n=%d q=%d, %d pairs:" % (n, q, B)
    for name, p in (("full", p_full), ("basemul", p_bm), ("basemul+in-range", p_flag)):
        o = torch.empty((B, n), dtype=torch.int32, device=dev)
        t = timeit(lambda: p.polymul(o, a, b))
        outs.append(o)
        line += "  %s %.4g/s (%.3f of HBM peak)" % (name, B / t, 12 * n * B / t / PEAK)
    same = torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ok &= same
    print(line, " outputs", "equal" if same else "DIFFER", flush=True)
    key = torch.randint(0, q, (n,), dtype=torch.int32, device=dev, generator=g).to(torch.int16)
    o1, o2 = torch.empty((B, n), dtype=torch.int32, device=dev), torch.empty((B, n), dtype=torch.int32, device=dev)
    t1 = timeit(lambda: p_bm.mul_key(o1, a, key)); t2 = timeit(lambda: p_flag.mul_key(o2, a, key))
    same = torch.equal(o1, o2); ok &= same
    print("key product n=%d: checked %.4g/s (%.3f)  in-range %.4g/s (%.3f)  outputs %s" % (
        n, B / t1, 8 * n * B / t1 / PEAK, B / t2, 8 * n * B / t2 / PEAK, "equal" if same else "DIFFER"), flush=True)
    del a, b, outs, o1, o2
for q, tw, shapes in ((7681, 16, ((2, 2), (3, 3), (4, 4))), (8380417, 32, ((5, 4),))):
    n = 256
    w, r = O.tables(q, n, tw)
    p_def = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    p_flag = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    p_flag.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    for k, l in shapes:
        Bm = 1 << 17
        A = torch.randint(0, q, (Bm, k * l, n), dtype=torch.int32, device=dev, generator=g)
        sv = torch.randint(-4, 5, (Bm, l, n), dtype=torch.int32, device=dev, generator=g)
        o1 = torch.empty((Bm, k, n), dtype=torch.int32, device=dev); o2 = torch.empty_like(o1)
        t1 = timeit(lambda: p_def.matvec(o1, A, sv, k, l)); t2 = timeit(lambda: p_flag.matvec(o2, A, sv, k, l))
        same = torch.equal(o1, o2); ok &= same
        bpi = 4 * n * (k * l + l + k)
        print("mat-vec q=%d k=%d l=%d: checked %.4g/s (%.3f)  in-range %.4g/s (%.3f)  outputs %s" % (
            q, k, l, Bm / t1, bpi * Bm / t1 / PEAK, Bm / t2, bpi * Bm / t2 / PEAK, "equal" if same else "DIFFER"), flush=True)
        del A, sv, o1, o2
sys.exit(0 if ok else 1)
