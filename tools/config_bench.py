"""Device-resident throughput of the BASELINE.json configs beyond the headline (C2 n=1024, C3 Kyber mat-vec,
C4 Dilithium, key products, variant-exact transforms) with the HBM fraction of each.
usage: python tools/config_bench.py [--json out.json]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O

PEAK = 6550.4
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
res = {}


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


def report(name, units, secs, bytes_per_unit, unit="polymul"):
    rate = units / secs
    res[name] = {"per_s": rate, "unit": unit + "/s", "GBps": rate * bytes_per_unit / 1e9, "hbm_frac": rate * bytes_per_unit / 1e9 / PEAK}
    print("%-44s %10.4g %s/s  %7.0f GB/s  %.3f of HBM peak" % (name, rate, unit, rate * bytes_per_unit / 1e9, rate * bytes_per_unit / 1e9 / PEAK))


# SCGPU_BENCH_IN_RANGE=1: the fused plans carry SCGPU_PLAN_INPUTS_IN_RANGE (no range votes; the operands here are canonical)
IN_RANGE = os.environ.get("SCGPU_BENCH_IN_RANGE", "0") not in ("", "0")


def rnd(q, shape):
    return torch.randint(0, q, shape, dtype=torch.int32, device=dev, generator=g)


# C2: polymul n = 512 / 1024, q = 12289; BLISS-style key product (shared SINT16 key)
for n in (512, 1024):
    q, B = 12289, (1 << 29) // (4 * n)
    w, r = O.tables(q, n, 16)
    pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    if IN_RANGE: pl.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    a, b, o = rnd(q, (B, n)), rnd(q, (B, n)), torch.empty((B, n), dtype=torch.int32, device=dev)
    report("C2 polymul n=%d q=12289" % n, B, timeit(lambda: pl.polymul(o, a, b)), 12 * n)
    key = rnd(q, (n,)).to(torch.int16)
    report("C2 key product n=%d (shared SINT16 key)" % n, B, timeit(lambda: pl.mul_key(o, a, key)), 8 * n)
    report("C2 canonical fwd NTT n=%d (normalize o fwd_ntt)" % n, B, timeit(lambda: pl.ntt_canonical(o, a)), 8 * n, "ntt")
    report("C2 canonical inv NTT n=%d" % n, B, timeit(lambda: pl.ntt_canonical(o, a, inverse=True)), 8 * n, "ntt")
    for v, vn in ((sc.REFERENCE, "reference"), (sc.BARRETT, "barrett"), (sc.FP, "fp"), (sc.AVX, "avx")):
        pe = sc.NttPlan(n, q, v, w, r)
        report("exact fwd_ntt_32_16 n=%d %s" % (n, vn), B, timeit(lambda: pe.batch(sc.OP_FWD, o, a)), 8 * n, "ntt")
        report("exact inv_ntt_32_16 n=%d %s" % (n, vn), B, timeit(lambda: pe.batch(sc.OP_INV, o, a)), 8 * n, "ntt")
    report("exact normalize_32 n=%d" % n, B, timeit(lambda: pe.batch(sc.OP_NORMALIZE, o, a)), 8 * n, "poly")
    report("exact mul_32_pointwise n=%d" % n, B, timeit(lambda: pe.batch(sc.OP_PW, o, a, b)), 12 * n, "poly")
    del a, b, o
# C3: Kyber mat-vec q = 7681 n = 256, k = l = 2, 3, 4
q, n = 7681, 256
w, r = O.tables(q, n, 16)
pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
if IN_RANGE: pl.set_flags(sc.PLAN_INPUTS_IN_RANGE)
for k in (2, 3, 4):
    B = 1 << 17
    A, s = rnd(q, (B, k * k, n)), torch.randint(-4, 5, (B, k, n), dtype=torch.int32, device=dev, generator=g)
    o = torch.empty((B, k, n), dtype=torch.int32, device=dev)
    report("C3 Kyber mat-vec k=%d" % k, B, timeit(lambda: pl.matvec(o, A, s, k, k)), 4 * n * (k * k + 2 * k), "instance")
    del A, s, o
a, b = rnd(q, (1 << 20, n)), rnd(q, (1 << 20, n))
o = torch.empty_like(a)
report("C3 polymul n=256 q=7681", 1 << 20, timeit(lambda: pl.polymul(o, a, b)), 12 * n)
# C4: Dilithium q = 8380417 n = 256
q = 8380417
w, r = O.tables(q, n, 32)
pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
if IN_RANGE: pl.set_flags(sc.PLAN_INPUTS_IN_RANGE)
a, b = rnd(q, (1 << 20, n)), rnd(q, (1 << 20, n))
report("C4 polymul n=256 q=8380417 (Shoup, warp-local)", 1 << 20, timeit(lambda: pl.polymul(o, a, b)), 12 * n)
report("C4 canonical fwd NTT n=256 q=8380417", 1 << 20, timeit(lambda: pl.ntt_canonical(o, a)), 8 * n, "ntt")
report("C4 canonical inv NTT n=256 q=8380417", 1 << 20, timeit(lambda: pl.ntt_canonical(o, a, inverse=True)), 8 * n, "ntt")
for v, vn in ((sc.REFERENCE, "reference"), (sc.FP, "fp")):
    pe = sc.NttPlan(n, q, v, w, r)
    report("C4 exact fwd_ntt_32_32 %s" % vn, 1 << 20, timeit(lambda: pe.batch(sc.OP_FWD, o, a)), 8 * n, "ntt")
    report("C4 exact inv_ntt_32_32 %s" % vn, 1 << 20, timeit(lambda: pe.batch(sc.OP_INV, o, a)), 8 * n, "ntt")
A, s = rnd(q, (1 << 15, 20, n)), torch.randint(-2, 3, (1 << 15, 4, n), dtype=torch.int32, device=dev, generator=g)
o5 = torch.empty((1 << 15, 5, n), dtype=torch.int32, device=dev)
report("C4 Dilithium mat-vec k=5 l=4", 1 << 15, timeit(lambda: pl.matvec(o5, A, s, 5, 4)), 4 * n * (20 + 4 + 5), "instance")
A6, s6 = rnd(q, (1 << 14, 30, n)), torch.randint(-2, 3, (1 << 14, 5, n), dtype=torch.int32, device=dev, generator=g)
o6 = torch.empty((1 << 14, 6, n), dtype=torch.int32, device=dev)
report("C4 Dilithium mat-vec k=6 l=5", 1 << 14, timeit(lambda: pl.matvec(o6, A6, s6, 6, 5)), 4 * n * (30 + 5 + 6), "instance")
del A, s, o5, A6, s6, o6, a, b, o
# the other moduli of the reference's parameter sets (ENS / DLP, Ring-TESLA): 32-bit tables; 51750913 (26 bits) is beyond
# both fused arithmetics and runs the Montgomery kernels of round 1
for q, n in ((5767169, 512), (10223617, 1024), (51750913, 512), (51750913, 1024)):
    B = (1 << 29) // (4 * n)
    w, r = O.tables(q, n, 32)
    pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    if IN_RANGE: pl.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    a, b, o = rnd(q, (B, n)), rnd(q, (B, n)), torch.empty((B, n), dtype=torch.int32, device=dev)
    report("polymul n=%d q=%d" % (n, q), B, timeit(lambda: pl.polymul(o, a, b)), 12 * n)
    report("canonical fwd NTT n=%d q=%d" % (n, q), B, timeit(lambda: pl.ntt_canonical(o, a)), 8 * n, "ntt")
    pe = sc.NttPlan(n, q, sc.AVX, w, r)
    report("exact fwd_ntt_32_32 n=%d q=%d avx" % (n, q), B, timeit(lambda: pe.batch(sc.OP_FWD, o, a)), 8 * n, "ntt")
    del a, b, o
if len(sys.argv) > 2 and sys.argv[1] == "--json":
    json.dump(res, open(sys.argv[2], "w"), indent=1)
