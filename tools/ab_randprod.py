"""A/B of the on-device matrix generation for the module product (scgpu_rand_product_csprng_batch), AES-CTR-DRBG:
one warp per instance (SCGPU_GEN_AES_WARP=1) against DRBG set-up per thread + one counter-addressed block per thread over
bank-replicated tables (default).  Same seeds, outputs compared.  usage: python tools/ab_randprod.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e-3


ok = True
for q, tw, qbits, shapes in ((7681, 16, 13, ((2, 2), (3, 3), (4, 4))), (8380417, 32, 23, ((5, 4),))):
    n = 256
    w, r = O.tables(q, n, tw)
    pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    for k, l in shapes:
        inst = 1 << 17
        sv = torch.randint(-4, 5, (inst, l, n), dtype=torch.int32, device=dev, generator=g)
        sd = torch.randint(0, 256, (inst, 32), dtype=torch.uint8, device=dev, generator=g)
        outs, line = [], "rand product q=%d k=%d l=%d:" % (q, k, l)
        for name, prng, env in (("AES warp-per-instance", sc.PRNG_AES_CTR_DRBG, "1"), ("AES block-per-thread", sc.PRNG_AES_CTR_DRBG, "0"),
                                ("ChaCha20", sc.PRNG_CHACHA, "0")):
            os.environ["SCGPU_GEN_AES_WARP"] = env
            to = torch.zeros((inst, k, n), dtype=torch.int32, device=dev)
            t = timeit(lambda: pl.rand_product(to, sv, sd, prng, qbits, k, l))
            outs.append(to)
            line += "  %s %.4g inst/s" % (name, inst / t)
        same = torch.equal(outs[0], outs[1])
        ok &= same
        print(line, " AES outputs", "equal" if same else "DIFFER", flush=True)
os.environ["SCGPU_GEN_AES_WARP"] = "0"
sys.exit(0 if ok else 1)
