"""Groups per work-counter claim (SCGPU_CLAIM_CHUNK = 1 / 2 / 4), interleaved in one process on one box."""
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
def rnd(q, shape): return torch.randint(0, q, shape, dtype=torch.int32, device=dev, generator=g)
legs = []
n, q, B = 512, 12289, 1 << 20
w, r = O.tables(q, n, 16)
plan = sc.NttPlan(n, q, sc.REFERENCE, w, r); plan.set_flags(sc.PLAN_INPUTS_IN_RANGE)
chk = sc.NttPlan(n, q, sc.REFERENCE, w, r)
a, b = rnd(q, (B, n)), rnd(q, (B, n)); out = torch.empty_like(a)
key = rnd(q, (n,)).to(torch.int16)
legs += [("polymul n512 inrange", B, 12 * n, lambda: plan.polymul(out, a, b)),
         ("polymul n512 checked", B, 12 * n, lambda: chk.polymul(out, a, b)),
         ("key product n512 inrange", B, 8 * n, lambda: plan.mul_key(out, a, key)),
         ("canonical fwd n512", B, 8 * n, lambda: plan.ntt_canonical(out, a)),
         ("canonical inv n512", B, 8 * n, lambda: plan.ntt_canonical(out, a, inverse=True))]
pe = sc.NttPlan(n, q, sc.AVX, w, r)
legs += [("exact fwd n512 avx", B, 8 * n, lambda: pe.batch(sc.OP_FWD, out, a)), ("exact inv n512 avx", B, 8 * n, lambda: pe.batch(sc.OP_INV, out, a))]
a1, b1, o1 = a.view(B // 2, 1024), b.view(B // 2, 1024), out.view(B // 2, 1024)
w1, r1 = O.tables(q, 1024, 16)
p1 = sc.NttPlan(1024, q, sc.REFERENCE, w1, r1); p1.set_flags(sc.PLAN_INPUTS_IN_RANGE)
legs += [("polymul n1024 inrange", B // 2, 12 * 1024, lambda: p1.polymul(o1, a1, b1))]
w2, r2 = O.tables(7681, 256, 16)
p2 = sc.NttPlan(256, 7681, sc.REFERENCE, w2, r2); p2.set_flags(sc.PLAN_INPUTS_IN_RANGE)
a2, b2 = rnd(7681, (B, 256)), rnd(7681, (B, 256)); o2 = torch.empty_like(a2)
legs += [("polymul n256 q7681 inrange", B, 12 * 256, lambda: p2.polymul(o2, a2, b2))]
k = 3; inst = 1 << 17
A = rnd(7681, (inst, k * k, 256)); sv = rnd(7681, (inst, k, 256)); to = torch.empty_like(sv)
legs += [("kyber matvec k3 inrange", inst, 4 * 256 * (k * k + 2 * k), lambda: p2.matvec(to, A, sv, k, k))]
w3, r3 = O.tables(8380417, 256, 32)
p3 = sc.NttPlan(256, 8380417, sc.REFERENCE, w3, r3); p3.set_flags(sc.PLAN_INPUTS_IN_RANGE)
a3, b3 = rnd(8380417, (B, 256)), rnd(8380417, (B, 256)); o3 = torch.empty_like(a3)
legs += [("polymul n256 q8380417 inrange", B, 12 * 256, lambda: p3.polymul(o3, a3, b3))]
kk, ll, inst3 = 5, 4, 1 << 15
A3 = rnd(8380417, (inst3, kk * ll, 256)); s3 = rnd(8380417, (inst3, ll, 256)); t3 = torch.empty((inst3, kk, 256), dtype=torch.int32, device=dev)
legs += [("dilithium matvec k5 l4 inrange", inst3, 4 * 256 * (kk * ll + kk + ll), lambda: p3.matvec(t3, A3, s3, kk, ll))]
res = {}
for rep in range(3):
    for name, units, bpu, fn in legs:
        for ck in ("1", "2", "4"):
            os.environ["SCGPU_CLAIM_CHUNK"] = ck
            t = timeit(fn)
            res.setdefault((name, ck), []).append(units * bpu / t / 1e9 / 6550.4)
# mat-vec with the 32-bit stash: one warp per output row (default) against one warp per instance group
for rep in range(3):
    for name, units, bpu, fn in legs:
        if "dilithium" not in name:
            continue
        for tag, val in (("rows", "0"), ("one_warp", "1")):
            os.environ["SCGPU_MATVEC_ONE_WARP"] = val
            os.environ["SCGPU_CLAIM_CHUNK"] = "2"
            t = timeit(fn)
            res.setdefault((name, tag), []).append(units * bpu / t / 1e9 / 6550.4)
os.environ["SCGPU_MATVEC_ONE_WARP"] = "0"
for name, units, bpu, fn in legs:
    if (name, "rows") in res:
        print("%-32s" % name + "".join("  %s: %s" % (tag, " ".join("%.3f" % x for x in res[(name, tag)])) for tag in ("rows", "one_warp")), flush=True)
for name, units, bpu, fn in legs:
    print("%-32s" % name + "".join("  chunk %s: %s" % (ck, " ".join("%.3f" % x for x in res[(name, ck)])) for ck in ("1", "2", "4")), flush=True)
