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
# the other shapes of bench.py's `other_shapes` (plans without range votes, as quoted there)
plan.set_flags(sc.PLAN_INPUTS_IN_RANGE)
plan.mul_key(out, a, key); units.append(("k_polymul_w32", "k_polymul_w32_key16_n512_inrange", B))
plan.ntt_canonical(out, a); units.append(("k_ntt_w32", "k_ntt_w32_fwd_n512_inrange", B))
plan.ntt_canonical(out, a, inverse=True); units.append(("k_ntt_w32", "k_ntt_w32_inv_n512_inrange", B))
plan.set_flags(0)
w1, r1 = O.tables(q, 1024, 16)
p1 = sc.NttPlan(1024, q, sc.REFERENCE, w1, r1); p1.set_flags(sc.PLAN_INPUTS_IN_RANGE)
a1, b1, o1 = a.view(B // 2, 1024), b.view(B // 2, 1024), out.view(B // 2, 1024)
p1.polymul(o1, a1, b1); units.append(("k_polymul_w32", "k_polymul_w32_n1024_inrange", B // 2))
w2, r2 = O.tables(7681, 256, 16)
p2 = sc.NttPlan(256, 7681, sc.REFERENCE, w2, r2); p2.set_flags(sc.PLAN_INPUTS_IN_RANGE)
a2 = torch.randint(0, 7681, (B, 256), dtype=torch.int32, device=dev, generator=g)
b2 = torch.randint(0, 7681, (B, 256), dtype=torch.int32, device=dev, generator=g)
o2 = torch.empty_like(a2)
p2.polymul(o2, a2, b2); units.append(("k_polymul_w32", "k_polymul_w32_n256_q7681_inrange", B))
kk, inst = 3, 1 << 15
A2 = torch.randint(0, 7681, (inst, kk * kk, 256), dtype=torch.int32, device=dev, generator=g)
s2 = torch.randint(-4, 5, (inst, kk, 256), dtype=torch.int32, device=dev, generator=g)
t2 = torch.empty((inst, kk, 256), dtype=torch.int32, device=dev)
p2.matvec(t2, A2, s2, kk, kk); units.append(("k_matvec16_w32", "k_matvec16_w32_kyber_k3_inrange", inst))
w3, r3 = O.tables(8380417, 256, 32)
p3 = sc.NttPlan(256, 8380417, sc.REFERENCE, w3, r3); p3.set_flags(sc.PLAN_INPUTS_IN_RANGE)
a3 = torch.randint(0, 8380417, (B, 256), dtype=torch.int32, device=dev, generator=g)
b3 = torch.randint(0, 8380417, (B, 256), dtype=torch.int32, device=dev, generator=g)
p3.polymul(o2, a3, b3); units.append(("k_polymul_w32", "k_polymul_w32_n256_q8380417_inrange", B))
inst3 = 1 << 14
A3 = torch.randint(0, 8380417, (inst3, 20, 256), dtype=torch.int32, device=dev, generator=g)
s3 = torch.randint(-2, 3, (inst3, 4, 256), dtype=torch.int32, device=dev, generator=g)
t3 = torch.empty((inst3, 5, 256), dtype=torch.int32, device=dev)
p3.matvec(t3, A3, s3, 5, 4); units.append(("k_matvec_rows_w32", "k_matvec_w32_dilithium_k5_l4_inrange", inst3))
del a2, b2, o2, a3, b3, A2, s2, t2, A3, s3, t3
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
