"""Small driver for ncu captures: a few launches of each hot kernel at benchmark shape.
Usage (on the GPU box): ncu ... python tools/profile_run.py [polymul|fwd|gauss|all] [log2_batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc  # noqa: E402
import _oracle as O  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 18
n, q, B = 512, 12289, 1 << lb
w, r = O.tables(q, n, 16)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
a = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
b = torch.randint(0, q, (B, n), dtype=torch.int32, device=dev, generator=g)
out = torch.empty_like(a)
if what in ("polymul", "polymul_inrange", "all"):
    plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    if what == "polymul_inrange":
        plan.set_flags(sc.PLAN_INPUTS_IN_RANGE)          # the bench's headline configuration: no range vote
    for _ in range(5):
        plan.polymul(out, a, b)
if what == "keyproduct":
    plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    key = torch.randint(0, q, (n,), dtype=torch.int32, device=dev, generator=g).to(torch.int16)
    for _ in range(5):
        plan.mul_key(out, a, key)
if what in ("fwd", "all"):
    for v in (sc.REFERENCE, sc.FP, sc.BARRETT):
        p = sc.NttPlan(n, q, v, w, r)
        for _ in range(3):
            p.batch(sc.OP_FWD, out, a)
        p.batch(sc.OP_INV, out, a)
        p.batch(sc.OP_NORMALIZE, out, a)
if what in ("gauss", "all"):
    gp = sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 215.0)
    seeds = torch.randint(0, 256, (B // 2, 40), dtype=torch.uint8, device=dev, generator=g)
    smp = torch.empty((B // 2, n), dtype=torch.int32, device=dev)
    for prng in (sc.PRNG_AES_CTR_DRBG, sc.PRNG_CHACHA):
        for _ in range(3):
            gp.streams(prng, seeds, n, smp)
if what in ("exact_avx",):
    p = sc.NttPlan(n, q, sc.AVX, w, r)
    for _ in range(2):
        p.batch(sc.OP_FWD, out, a)
    p.batch(sc.OP_INV, out, a)
if what in ("ber", "ky"):
    sid = sc.SAMPLER_BERNOULLI if what == "ber" else sc.SAMPLER_KNUTH_YAO
    gp = sc.GaussPlan(sid, 64, 0, 13.42, 215.0)
    ns = 1 << 16
    seeds = torch.randint(0, 256, (ns, 40), dtype=torch.uint8, device=dev, generator=g)
    smp = torch.empty((ns, 128), dtype=torch.int32, device=dev)
    for _ in range(2):
        gp.streams(sc.PRNG_CHACHA, seeds, 128, smp)
if what in ("randprod",):
    q3, k = 7681, 3
    w3, r3 = O.tables(q3, 256, 16)
    p3 = sc.NttPlan(256, q3, sc.REFERENCE, w3, r3)
    inst = 1 << 15
    sd = torch.randint(0, 256, (inst, 32), dtype=torch.uint8, device=dev, generator=g)
    sv = torch.randint(-4, 5, (inst, k, 256), dtype=torch.int32, device=dev, generator=g)
    to = torch.empty((inst, k, 256), dtype=torch.int32, device=dev)
    for prng in (sc.PRNG_CHACHA, sc.PRNG_AES_CTR_DRBG):
        for _ in range(2):
            p3.rand_product(to, sv, sd, prng, 13, k, k)
torch.cuda.synchronize()
print("done")
if what in ("dilithium",):
    # C4: q = 8380417, n = 256, Shoup-policy fused product (k_polymul_w32<ArSh, 8, ...>)
    qq, nn = 8380417, 256
    ww, rr = O.tables(qq, nn, 32)
    pd = sc.NttPlan(nn, qq, sc.REFERENCE, ww, rr)
    pd.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    xa = torch.randint(0, qq, (B, nn), dtype=torch.int32, device=dev, generator=g)
    xb = torch.randint(0, qq, (B, nn), dtype=torch.int32, device=dev, generator=g)
    xo = torch.empty_like(xa)
    for _ in range(5):
        pd.polymul(xo, xa, xb)
if what in ("exact_barrett",):
    p = sc.NttPlan(n, q, sc.BARRETT, w, r)
    for _ in range(3):
        p.batch(sc.OP_INV, out, a)
if what in ("dil_matvec",):
    qq, nn, kk, ll = 8380417, 256, 5, 4
    inst = 1 << 15
    ww, rr = O.tables(qq, nn, 32)
    pd = sc.NttPlan(nn, qq, sc.REFERENCE, ww, rr)
    pd.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    A3 = torch.randint(0, qq, (inst, kk * ll, nn), dtype=torch.int32, device=dev, generator=g)
    s3 = torch.randint(0, qq, (inst, ll, nn), dtype=torch.int32, device=dev, generator=g)
    t3 = torch.empty((inst, kk, nn), dtype=torch.int32, device=dev)
    for _ in range(5):
        pd.matvec(t3, A3, s3, kk, ll)
if what in ("polymul1024", "key1024"):
    w1, r1 = O.tables(q, 1024, 16)
    p1 = sc.NttPlan(1024, q, sc.REFERENCE, w1, r1)
    p1.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    a1, b1, o1 = a.view(B // 2, 1024), b.view(B // 2, 1024), out.view(B // 2, 1024)
    key1 = torch.randint(0, q, (1024,), dtype=torch.int32, device=dev, generator=g).to(torch.int16)
    for _ in range(5):
        if what == "polymul1024":
            p1.polymul(o1, a1, b1)
        else:
            p1.mul_key(o1, a1, key1)
if what in ("kyber_matvec",):
    qq, nn, kk = 7681, 256, 3
    inst = 1 << 16
    ww, rr = O.tables(qq, nn, 16)
    pk = sc.NttPlan(nn, qq, sc.REFERENCE, ww, rr)
    pk.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    A2 = torch.randint(0, qq, (inst, kk * kk, nn), dtype=torch.int32, device=dev, generator=g)
    s2 = torch.randint(-4, 5, (inst, kk, nn), dtype=torch.int32, device=dev, generator=g)
    t2 = torch.empty((inst, kk, nn), dtype=torch.int32, device=dev)
    for _ in range(5):
        pk.matvec(t2, A2, s2, kk, kk)
