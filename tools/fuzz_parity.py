"""Randomised parity soak: random parameter sets, batch sizes, alignments, input ranges and arithmetic modes of
the fused kernels, canonical transforms, mat-vec and samplers against the CPU oracle port.
usage: python tools/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import libsafecrypto_b200 as sc
import _oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
dev = "cuda:0"
P = O.port()
PARAMS = [(12289, 512, 16), (12289, 1024, 16), (7681, 256, 16), (8380417, 256, 32), (18433, 512, 16), (8399873, 512, 32),
          (12289, 256, 16), (18433, 1024, 16), (51750913, 512, 32), (51750913, 1024, 32), (5767169, 1024, 32), (10223617, 512, 32)]
plans = {}


def plan(q, n, tw):
    key = (q, n)
    if key not in plans:
        w, r = O.tables(q, n, tw)
        flagged = sc.NttPlan(n, q, sc.REFERENCE, w, r)
        flagged.set_flags(sc.PLAN_INPUTS_IN_RANGE)          # no range votes: only used with operands inside +-4q
        plans[key] = (sc.NttPlan(n, q, sc.REFERENCE, w, r), w, r, flagged)
    return plans[key]


def inputs(kind, q, shape):
    if kind == 0: return rng.integers(0, q, shape).astype(np.int32)
    if kind == 1: return rng.integers(-q + 1, q, shape).astype(np.int32)
    if kind == 2: return rng.integers(-4 * q, 4 * q + 1, shape).astype(np.int32)
    if kind == 3: return rng.integers(-2**31, 2**31, shape).astype(np.int32)
    x = rng.integers(-5, 6, shape).astype(np.int32)
    x[rng.random(shape) < 0.01] = 2**31 - 1
    return x


def dev_rows(x, off):
    """device copy of x whose first element sits `off` int32 words after a 256-byte boundary"""
    buf = torch.zeros(x.size + 8, dtype=torch.int32, device=dev)
    buf[off:off + x.size] = torch.from_numpy(x.reshape(-1)).to(dev)
    return buf[off:off + x.size].view(*x.shape)


t0, iters, fails = time.time(), 0, 0
counts = {}
while time.time() - t0 < budget:
    q, n, tw = PARAMS[rng.integers(len(PARAMS))]
    pl0, w, r, plf = plan(q, n, tw)
    pl = pl0
    mode = int(rng.choice([0, 0, 0, 1, 2, 3, 4]))
    sc.lib().scgpu_set_fast_arith(mode)
    what = int(rng.integers(5))
    rows = int(rng.choice([1, 2, 3, 5, 8, 31, 64, 257, 1000, 4099, 4099, 13001]))      # 13001: beyond one grid-full for every n (work counter)
    offa, offb = int(rng.choice([0, 0, 0, 1, 2, 4])), int(rng.choice([0, 0, 1, 4]))
    name = ""
    try:
        if what == 0:
            name = "polymul"
            ka, kb = int(rng.integers(5)), int(rng.integers(5))
            a, b = inputs(ka, q, (rows, n)), inputs(kb, q, (rows, n))
            if ka <= 2 and kb <= 2 and rng.random() < 0.5:
                pl, name = plf, "polymul_inrange"
            out = torch.empty((rows, n), dtype=torch.int32, device=dev)
            if rng.random() < 0.2:
                b = b[:1]
                pl.polymul(out, dev_rows(a, offa), dev_rows(b[0], offb))
                bb = np.tile(b, (rows, 1))
            else:
                pl.polymul(out, dev_rows(a, offa), dev_rows(b, offb))
                bb = b
            exp = P.ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, tw, a, bb, w, r)
        elif what == 1 and tw == 16:
            name = "key16"
            ka = int(rng.integers(5))
            a = inputs(ka, q, (rows, n))
            key = rng.integers(-32768, 32768, (rows, n)).astype(np.int16) if rng.random() < 0.4 else rng.integers(-32768, 32768, (n,)).astype(np.int16)
            if ka <= 2 and rng.random() < 0.5:
                pl, name = plf, "key16_inrange"
            out = torch.empty((rows, n), dtype=torch.int32, device=dev)
            pl.mul_key(out, dev_rows(a, offa), torch.from_numpy(key).to(dev))
            exp = P.ntt_batch(O.REFERENCE, O.OP_TRIPLE16, n, q, tw, a, key if key.ndim == 2 else np.tile(key, (rows, 1)), w, r)
        elif what == 2:
            name = "fwd_canonical"
            a = inputs(rng.integers(5), q, (rows, n))
            out = torch.empty((rows, n), dtype=torch.int32, device=dev)
            pl.ntt_canonical(out, dev_rows(a, offa))
            exp = P.ntt_batch(O.REFERENCE, O.OP_NORMALIZE, n, q, tw, P.ntt_batch(O.REFERENCE, O.OP_FWD, n, q, tw, a, None, w, r))
        elif what == 3:
            name = "inv_canonical"
            a = inputs(rng.integers(3), q, (rows, n))
            out = torch.empty((rows, n), dtype=torch.int32, device=dev)
            pl.ntt_canonical(out, dev_rows(a, offa), inverse=True)
            exp = P.ntt_batch(O.REFERENCE, O.OP_INV, n, q, tw, a, None, w, r)
        elif what == 4 and n == 256:
            name = "matvec"
            k, l = int(rng.integers(1, 6)), int(rng.integers(1, 5))
            rows = min(rows, 300)
            A = inputs(rng.integers(3), q, (rows, k * l, n))
            ks = int(rng.choice([0, 1, 4]))
            s = inputs(ks, q, (rows, l, n))
            if ks <= 1 and rng.random() < 0.5:
                pl, name = plf, "matvec_inrange"
            out = torch.empty((rows, k, n), dtype=torch.int32, device=dev)
            pl.matvec(out, dev_rows(A, offa), dev_rows(s, offb), k, l)
            sh = P.ntt_batch(O.REFERENCE, O.OP_FWD, n, q, tw, s.reshape(-1, n), None, w, r).reshape(rows, l, n)
            exp = np.zeros((rows, k, n), dtype=np.int32)
            for i in range(k):
                acc = np.zeros((rows, n), dtype=np.int64)
                for j in range(l):
                    acc += P.ntt_batch(O.REFERENCE, O.OP_PW, n, q, tw, A[:, i * l + j], sh[:, j])
                t = P.ntt_batch(O.REFERENCE, O.OP_NORMALIZE, n, q, tw, np.mod(acc, q).astype(np.int32))
                t = P.ntt_batch(O.REFERENCE, O.OP_INV, n, q, tw, t, None, w, r)
                exp[:, i] = P.ntt_batch(O.REFERENCE, O.OP_NORMALIZE, n, q, tw, t)
        else:
            continue
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        ok = np.array_equal(got, exp)
    except Exception as ex:            # noqa: BLE001
        ok = False
        print("EXCEPTION", name, q, n, mode, rows, repr(ex))
    iters += 1
    counts[name] = counts.get(name, 0) + 1
    if not ok:
        fails += 1
        print("MISMATCH", name, "q=%d n=%d mode=%d rows=%d offa=%d offb=%d" % (q, n, mode, rows, offa, offb))
        if fails > 10:
            break
sc.lib().scgpu_set_fast_arith(0)
# samplers: random seeds / lengths / modes against the port
gi = 0
while time.time() - t0 < budget * 1.25:
    prec = int(rng.choice([32, 64])); bl = int(rng.choice([0, 0, 1, 2])); prng = int(rng.choice([sc.PRNG_CHACHA, sc.PRNG_AES_CTR_DRBG]))
    ns, n, calls = int(rng.choice([1, 3, 33, 200])), int(rng.choice([1, 7, 64, 255, 512, 1000])), int(rng.choice([1, 2]))
    disc = int(rng.choice([0, 0, 2, 4, 6])); centre = int(rng.integers(-3, 4)); sl = int(rng.choice([32, 40, 64]))
    seeds = rng.integers(0, 256, (ns, sl)).astype(np.uint8)
    gp = sc.GaussPlan(sc.SAMPLER_CDF, prec, bl, 13.42, 215.0)
    out = torch.empty((ns, n * calls), dtype=torch.int32, device=dev)
    gp.streams(prng, torch.from_numpy(seeds).to(dev), n, out, calls=calls, centre=centre, discard=disc)
    torch.cuda.synchronize()
    exp = P.gauss_streams(O.SAMPLER_CDF, prec, bl, prng, 13.42, 215.0, seeds, n, discard=disc, centre=centre, calls=calls)
    gi += 1
    if not np.array_equal(out.cpu().numpy(), exp):
        fails += 1
        print("MISMATCH gauss prec=%d bl=%d prng=%d ns=%d n=%d calls=%d disc=%d" % (prec, bl, prng, ns, n, calls, disc))
print("fuzz: %d NTT cases %s, %d sampler cases, %d failures, %.0f s" % (iters, counts, gi, fails, time.time() - t0))
sys.exit(1 if fails else 0)
