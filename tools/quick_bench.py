"""Device-resident polymul timing only (no e2e / CPU legs): python tools/quick_bench.py [n] [q] [log2_batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
q = int(sys.argv[2]) if len(sys.argv) > 2 else 12289
lb = int(sys.argv[3]) if len(sys.argv) > 3 else 20
B = 1 << lb
w, r = O.tables(q, n, 16 if q < 32768 else 32)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
a = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
b = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(a)
plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
for mode, name in ((0, "auto"), (4, "shoup"), (3, "fq8"), (2, "barrett32")):
    sc.lib().scgpu_set_fast_arith(mode)
    for _ in range(3):
        plan.polymul(out, a, b)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        plan.polymul(out, a, b)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print("n=%d q=%d %-10s %.3f ms  %.4g polymul/s  %.0f GB/s" % (n, q, name, ms, B / ms * 1e3, 12 * n * B / ms / 1e6))
sc.lib().scgpu_set_fast_arith(0)
