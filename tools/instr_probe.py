"""Launches every hot kernel ONCE at a known size (for tools/instr_counts.py, which runs this under ncu and divides the
executed warp-instructions by the units printed here)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc  # noqa: E402
import _oracle as O  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
units = []          # (kernel name regex, key, units) in launch order

n, q, B = 512, 12289, 1 << 18
w, r = O.tables(q, n, 16)
a = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
b = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(a)
plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
plan.polymul(out, a, b); units.append(("k_polymul_w32", "k_polymul_w32_n512", B))
plan.set_flags(sc.PLAN_INPUTS_IN_RANGE)
plan.polymul(out, a, b); units.append(("k_polymul_w32", "k_polymul_w32_n512_inrange", B))
plan.set_flags(0)
key = torch.randint(0, q, (n,), dtype=torch.int32, device=dev, generator=g).to(torch.int16)
plan.mul_key(out, a, key); units.append(("k_polymul_w32", "k_polymul_w32_key16_n512", B))
plan.ntt_canonical(out, a); units.append(("k_ntt_w32", "k_ntt_w32_fwd_n512", B))
for v, vn in ((sc.REFERENCE, "reference"), (sc.BARRETT, "barrett"), (sc.AVX, "avx")):
    pe = sc.NttPlan(n, q, v, w, r)
    pe.batch(sc.OP_FWD, out, a); units.append(("k_exact_w32", "k_exact_w32_fwd_n512_%s" % vn, B))
    pe.batch(sc.OP_INV, out, a); units.append(("k_exact_w32", "k_exact_w32_inv_n512_%s" % vn, B))
gp = sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 215.0)
ns = 1 << 16
seeds = torch.randint(0, 256, (ns, 40), dtype=torch.uint8, device=dev, generator=g)
smp = torch.empty((ns, 512), dtype=torch.int32, device=dev)
for mode, fixed in (("fixed_probe", 1), ("guided", 0)):
    sc.lib().scgpu_set_fixed_probe_search(fixed)
    gp.streams(sc.PRNG_AES_CTR_DRBG, seeds, 512, smp); units.append(("k_cdf_aes", "k_cdf_aes_%s" % mode, ns * 512))
    gp.streams(sc.PRNG_CHACHA, seeds, 512, smp); units.append(("k_cdf_chacha", "k_cdf_chacha_%s" % mode, ns * 512))
sc.lib().scgpu_set_fixed_probe_search(1)
nk = 1 << 14
for sid, name, kern in ((sc.SAMPLER_KNUTH_YAO, "k_stream_seq_ky64", "k_stream_seq"), (sc.SAMPLER_BERNOULLI, "k_ber_lanes", "k_ber_lanes")):
    gx = sc.GaussPlan(sid, 64, 0, 13.42, 215.0)
    sx = torch.empty((nk, 128), dtype=torch.int32, device=dev)
    gx.streams(sc.PRNG_CHACHA, seeds[:nk], 128, sx); units.append((kern, name, nk * 128))
torch.cuda.synchronize()
json.dump(units, open(os.path.join(ROOT, "gpurun_out", "instr_probe_units.json"), "w"))
print("probe done:", len(units), "launches")
