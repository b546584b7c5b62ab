import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O
k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
q, n, B = 7681, 256, 1 << 16
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
w, r = O.tables(q, n, 16)
pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
A = torch.randint(0, q, (B, k * k, n), dtype=torch.int32, device=dev, generator=g)
s = torch.randint(-4, 5, (B, k, n), dtype=torch.int32, device=dev, generator=g)
o = torch.empty((B, k, n), dtype=torch.int32, device=dev)
for _ in range(4):
    pl.matvec(o, A, s, k, k)
torch.cuda.synchronize()
