"""bench.py's order of legs around the canonical transforms, repeated, to find what made them slow there."""
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
n, q, B = 512, 12289, 1 << 20
w, r = O.tables(q, n, 16)
plan = sc.NttPlan(n, q, sc.REFERENCE, w, r); plan.set_flags(sc.PLAN_INPUTS_IN_RANGE)
chk = sc.NttPlan(n, q, sc.REFERENCE, w, r)
a = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
b = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(a)
key = torch.randint(0, q, (n,), dtype=torch.int32, device=dev, generator=g).to(torch.int16)
def show(tag, t): print("%-34s %.4g /s  %.3f" % (tag, B / t, B / t * 8 * n / 1e9 / 6550.4), flush=True)
for rnd in range(3):
    show("polymul inrange (12n)", timeit(lambda: plan.polymul(out, a, b)))
    show("key checked", timeit(lambda: chk.mul_key(out, a, key)))
    show("key inrange", timeit(lambda: plan.mul_key(out, a, key)))
    show("fwd canonical inrange", timeit(lambda: plan.ntt_canonical(out, a)))
    show("inv canonical inrange", timeit(lambda: plan.ntt_canonical(out, a, inverse=True)))
    show("fwd canonical checked", timeit(lambda: chk.ntt_canonical(out, a)))
    show("inv canonical checked", timeit(lambda: chk.ntt_canonical(out, a, inverse=True)))
    show("fwd canonical inrange out=b", timeit(lambda: plan.ntt_canonical(b, a)))
    show("fwd canonical inrange again", timeit(lambda: plan.ntt_canonical(out, a)))
    show("fwd canonical inrange 100 reps", timeit(lambda: plan.ntt_canonical(out, a), reps=100))
