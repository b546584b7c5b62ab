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
    del a, b, outs
sys.exit(0 if ok else 1)
