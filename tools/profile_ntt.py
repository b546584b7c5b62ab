"""A few launches of the canonical single-transform kernels for ncu: python tools/profile_ntt.py [n] [q] [log2_batch]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
q = int(sys.argv[2]) if len(sys.argv) > 2 else 12289
B = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 20)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
w, r = O.tables(q, n, 16 if q < 32768 else 32)
pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
a = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
o = torch.empty_like(a)
for _ in range(3):
    pl.ntt_canonical(o, a)
for _ in range(3):
    pl.ntt_canonical(a, o, inverse=True)
torch.cuda.synchronize()
