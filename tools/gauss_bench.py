"""Device-resident Gaussian sampling throughput: python tools/gauss_bench.py [log2_streams] [n]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
lb = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
ns = 1 << lb
seeds = torch.randint(0, 256, (ns, 40), dtype=torch.uint8, device=dev, generator=g)
smp = torch.empty((ns, n), dtype=torch.int32, device=dev)
for prec in (64, 32):
    gp = sc.GaussPlan(sc.SAMPLER_CDF, prec, 0, 13.42, 215.0)
    for name, prng in (("aes_ctr_drbg", sc.PRNG_AES_CTR_DRBG), ("chacha20", sc.PRNG_CHACHA)):
        for _ in range(2):
            gp.streams(prng, seeds, n, smp)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            gp.streams(prng, seeds, n, smp)
        e.record(); torch.cuda.synchronize()
        print("cdf%d %-13s %.4g samples/s" % (prec, name, 5 * ns * n / (s.elapsed_time(e) * 1e-3)))
