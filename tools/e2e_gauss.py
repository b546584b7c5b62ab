import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import torch, libsafecrypto_b200 as sc
gp = sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 215.0)
ns, n = 1 << 18, 512
hs = torch.randint(0, 256, (ns, 40), dtype=torch.uint8).pin_memory()
ho = torch.empty((ns, n), dtype=torch.int32).pin_memory()
for name, prng in (("aes", 0), ("chacha", 2)):
    gp.streams_host(prng, hs, n, ho)
    t0 = time.perf_counter()
    for _ in range(3): gp.streams_host(prng, hs, n, ho)
    dt = (time.perf_counter() - t0) / 3
    print(name, "%.3g samples/s  %.1f ms" % (ns * n / dt, dt * 1e3))
