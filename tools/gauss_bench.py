"""Device-resident Gaussian sampling throughput: python tools/gauss_bench.py [log2_streams] [n] [log2_streams_sequential]"""
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
for prec in (64, 32, -64):
    fixed = prec < 0
    prec = abs(prec)
    sc.lib().scgpu_set_fixed_probe_search(1 if fixed else 0)
    if fixed:
        print("fixed probe sequence (scgpu_set_fixed_probe_search(1)):")
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

sc.lib().scgpu_set_fixed_probe_search(0)
# sequential-semantics kernels (k_stream_seq, k_ber_lanes): shuffle / blinding / discard wrappers, Knuth-Yao, Bernoulli,
# Micciancio-Walter bootstrap; one lane per stream, so the rate depends on the stream count
for lb2 in ((int(sys.argv[3]),) if len(sys.argv) > 3 else (16, 18)):
    ns2 = 1 << lb2
    print("%d streams x %d samples:" % (ns2, n))
    seeds2 = seeds[:ns2] if ns2 <= ns else torch.randint(0, 256, (ns2, 40), dtype=torch.uint8, device=dev, generator=g)
    smp2 = torch.empty((ns2, n), dtype=torch.int32, device=dev)
    cases = [("cdf64 shuffle", sc.SAMPLER_CDF, 64, 2, 0), ("cdf64 blinding", sc.SAMPLER_CDF, 64, 1, 0), ("cdf64 normal discard=4", sc.SAMPLER_CDF, 64, 0, 4),
             ("knuth-yao 64", sc.SAMPLER_KNUTH_YAO, 64, 0, 0), ("knuth-yao 128", sc.SAMPLER_KNUTH_YAO, 128, 0, 0), ("knuth-yao 64 shuffle", sc.SAMPLER_KNUTH_YAO, 64, 2, 0),
             ("bernoulli 64", sc.SAMPLER_BERNOULLI, 64, 0, 0), ("bernoulli 64 discard=4", sc.SAMPLER_BERNOULLI, 64, 0, 4), ("mw bootstrap sigma 215", None, 64, 0, 0)]
    for label, smpl, prec, bl, disc in cases:
        gp = sc.GaussPlan(smpl, prec, bl, 13.42, 215.0) if smpl is not None else sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 0.0, mw=True)
        for name, prng in (("aes_ctr_drbg", sc.PRNG_AES_CTR_DRBG), ("chacha20", sc.PRNG_CHACHA)):
            run = (lambda: gp.streams(prng, seeds2, n, smp2, discard=disc)) if smpl is not None else (lambda: gp.mw_streams(prng, seeds2, n, smp2, 215.0, 0.25))
            run()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            run()
            e.record(); torch.cuda.synchronize()
            print("  %-24s %-13s %.4g samples/s" % (label, name, ns2 * n / (s.elapsed_time(e) * 1e-3)))
