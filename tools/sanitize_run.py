"""Small run of every fused kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O
dev = "cuda:0"
rng = np.random.default_rng(5)
ok = True
for q, n, tw in ((12289, 512, 16), (12289, 1024, 16), (7681, 256, 16), (8380417, 256, 32)):
    w, r = O.tables(q, n, tw)
    pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    rows = 13000                      # more than one grid-full for every n: the work-counter path
    a = rng.integers(0, q, (rows, n)).astype(np.int32); b = rng.integers(-q, q, (rows, n)).astype(np.int32)
    out = torch.empty((rows, n), dtype=torch.int32, device=dev)
    pl.polymul(out, torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
    torch.cuda.synchronize()
    ok &= np.array_equal(out.cpu().numpy()[:64], O.port().ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, tw, a[:64], b[:64], w, r))
    # round 2: shared-key product (k_key_residues + the residue-table kernel), per-row keys, and the kernels without range
    # votes (SCGPU_PLAN_INPUTS_IN_RANGE)
    ta = torch.from_numpy(a).to(dev)
    key = torch.from_numpy(rng.integers(0, q, n).astype(np.int16 if tw == 16 else np.int32)).to(dev)
    pl.mul_key(out, ta, key)
    torch.cuda.synchronize()
    if tw == 16:
        ok &= np.array_equal(out.cpu().numpy()[:64], O.port().ntt_batch(O.REFERENCE, O.OP_TRIPLE16, n, q, tw, a[:64], key.cpu().numpy(), w, r))
    keys = torch.from_numpy(rng.integers(0, q, (rows, n)).astype(np.int32)).to(dev)
    pl.mul_key(out, ta, keys)
    pl.set_flags(sc.PLAN_INPUTS_IN_RANGE)
    pl.polymul(out, ta, torch.from_numpy(b).to(dev))
    pl.mul_key(out, ta, key)
    pl.ntt_canonical(out, ta)
    pl.ntt_canonical(out, ta, inverse=True)
    torch.cuda.synchronize()
    ok &= np.array_equal(out.cpu().numpy()[:16], O.port().ntt_batch(O.REFERENCE, O.OP_INV, n, q, tw, a[:16], None, w, r))
    pl.set_flags(0)
    if n == 256:
        k = 3
        A = rng.integers(0, q, (9000, k * k, n)).astype(np.int32); s = rng.integers(-4, 5, (9000, k, n)).astype(np.int32)
        o = torch.empty((9000, k, n), dtype=torch.int32, device=dev)
        pl.matvec(o, torch.from_numpy(A).to(dev), torch.from_numpy(s).to(dev), k, k)
        torch.cuda.synchronize()
# chunked work-counter claims: more than 10 grid-fulls of groups, so that claims of several groups AND the single-group
# tail both run (warp32.cuh: Claim); fused product, canonical and variant-exact transforms
for chunk in ("2", "4"):
    os.environ["SCGPU_CLAIM_CHUNK"] = chunk
    q, n, tw = 7681, 256, 16
    w, r = O.tables(q, n, tw)
    pl = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    rows = 130001
    a = rng.integers(0, q, (rows, n)).astype(np.int32)
    ta = torch.from_numpy(a).to(dev); out = torch.empty_like(ta)
    pl.polymul(out, ta, ta)
    pl.ntt_canonical(out, ta)
    back = torch.empty_like(ta)
    pl.ntt_canonical(back, out, inverse=True)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(back, ta))
    pe = sc.NttPlan(n, q, sc.AVX, w, r)
    pe.batch(sc.OP_FWD, out, ta)
    torch.cuda.synchronize()
    ok &= np.array_equal(out.cpu().numpy()[-33:], O.port().ntt_batch(O.AVX, O.OP_FWD, n, q, tw, a[-33:], None, w, r))
del os.environ["SCGPU_CLAIM_CHUNK"]
gp = sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 215.0)
seeds = torch.from_numpy(rng.integers(0, 256, (4000, 40)).astype(np.uint8)).to(dev)
smp = torch.empty((4000, 64), dtype=torch.int32, device=dev)
for prng in (sc.PRNG_CHACHA, sc.PRNG_AES_CTR_DRBG):
    gp.streams(prng, seeds, 64, smp)
torch.cuda.synchronize()
print("sanitize run done, parity", ok)
