"""CPU checks of the fused kernels' static design: the shared-memory swizzle is a permutation that is
bank-conflict-free for every pass's access pattern, the pass schedule covers every stage exactly once, and
the merged-twiddle index maps the fused transform's output slot i to the reference's NTT coefficient brv(i)."""
import numpy as np
import pytest

import _oracle as O


def swz(idx, logn):        # libsafecrypto_b200/csrc/fast_common.cuh
    if logn == 8:
        return idx ^ ((idx >> 5) & 7) ^ (((idx >> 5) & 3) << 3)
    if logn == 9:
        return idx ^ ((idx >> 5) & 7) ^ (((idx >> 6) & 3) << 3)
    return idx ^ ((idx >> 5) & 7) ^ (((idx >> 5) & 1) << 3) ^ (((idx >> 7) & 1) << 4)


def passes(logn):
    out, p = [], 0
    while 3 * p < logn:
        j = min(3, logn - 3 * p)
        out.append((3 * p, j, ((1 << logn) >> (3 * p + 3)) if j == 3 else 1))
        p += 1
    return out


@pytest.mark.parametrize("logn", [8, 9, 10])
def test_swizzle_is_conflict_free_permutation(logn):
    n, t = 1 << logn, (1 << logn) // 8
    assert sorted(swz(i, logn) for i in range(n)) == list(range(n))
    for _, _, d in passes(logn):
        for warp in range(max(1, t // 32)):
            for m in range(8):
                banks = set()
                for lane in range(32):
                    tau = warp * 32 + lane
                    banks.add(swz((tau // d) * 8 * d + tau % d + m * d, logn) % 32)
                assert len(banks) == 32


@pytest.mark.parametrize("logn", [8, 9, 10])
def test_pass_schedule_covers_all_stages(logn):
    n = 1 << logn
    seen = []
    for s0, j, d in passes(logn):
        for delta_log in range(j - 1, -1, -1):
            seen.append(d << delta_log)          # butterfly distance in elements
        # every thread's 8 elements are distinct and the threads tile the polynomial
        idx = sorted((tau // d) * 8 * d + tau % d + m * d for tau in range(n // 8) for m in range(8))
        assert idx == list(range(n))
    assert seen == [n >> (s + 1) for s in range(logn)]


@pytest.mark.parametrize("q,n", [(12289, 512), (7681, 256), (12289, 1024)])
def test_merged_twiddle_ordering_matches_reference_domain(q, n):
    """Cooley-Tukey with zeta[k] = w[brv(k)] leaves the reference's NTT coefficient brv(i) in slot i
    (what lets key / matrix operands given in the reference's layout be read with a bit-reversed index)."""
    logn = n.bit_length() - 1
    w, r = O.tables(q, n, 16)
    brv = lambda i: int(format(i, "0%db" % logn)[::-1], 2)  # noqa: E731
    zeta = [0] + [int(w[brv(k)]) for k in range(1, n)]
    a = np.random.default_rng(1).integers(0, q, n)
    x, k, ln = [int(v) for v in a], 1, n // 2
    while ln >= 1:
        for start in range(0, n, 2 * ln):
            z = zeta[k]
            k += 1
            for j in range(start, start + ln):
                t = z * x[j + ln] % q
                x[j + ln] = (x[j] - t) % q
                x[j] = (x[j] + t) % q
        ln //= 2
    ref = O.port().ntt_batch(O.REFERENCE, O.OP_FWD, n, q, 16, a.astype(np.int32), None, w, r)[0].astype(np.int64) % q
    assert all(x[i] == ref[brv(i)] for i in range(n))


# ---- warp-local 32-coefficient schedule (libsafecrypto_b200/csrc/warp32.cuh) ---------------------------------

def pos32(e):
    return e + 4 * (e >> 5)


def slot32(logn, s, tau, r):
    n = 1 << logn
    t, ln = n // 32, n >> (s + 1)
    g = 16 // ln
    v = g if g < 4 else 4
    return (1 << s) + ((r // v) * t + tau) * v + (r % v)


@pytest.mark.parametrize("logn", [8, 9, 10])
def test_warp32_tile_layout(logn):
    """Padded tile: (1) injective; (2) pass-0 accesses (element tau + T m, 32-bit) of the warp's 32 / T polynomials,
    whose tiles are TS words apart, hit 32 distinct banks; (3) pass-1 accesses (128-bit at 36 tau + 4 k) are
    conflict-free per quarter-warp and 16-byte aligned; (4) both passes tile the polynomial."""
    n = 1 << logn
    t = n // 32
    pw = 32 // t
    ts = n + n // 8 + (t & 31)
    assert len({pos32(e) for e in range(n)}) == n and max(pos32(e) for e in range(n)) < ts
    for m in range(32):
        banks = {(p * ts + tau + pos32(t * m)) % 32 for p in range(pw) for tau in range(t)}
        assert len(banks) == 32
        assert all(pos32(tau + t * m) == tau + pos32(t * m) for tau in range(t))       # thread base + immediate
    for k in range(8):
        for quarter in range(4):
            lanes = range(8 * quarter, 8 * quarter + 8)
            groups = set()
            for lane in lanes:
                p, tau = lane // t, lane % t
                addr = p * ts + 36 * tau + 4 * k
                assert addr % 4 == 0 and all(pos32(32 * tau + 4 * k + j) == 36 * tau + 4 * k + j for j in range(4))
                groups.add((addr % 32) // 4)
            assert len(groups) == 8
    assert sorted(tau + t * m for tau in range(t) for m in range(32)) == list(range(n))
    assert sorted(32 * tau + i for tau in range(t) for i in range(32)) == list(range(n))


@pytest.mark.parametrize("logn", [8, 9, 10])
def test_warp32_twiddle_table_is_a_permutation_with_contiguous_vectors(logn):
    """Thread-major pass-1 table: a permutation inside every stage's region [2^s, 2^(s+1)), and the 4 (2, 1)
    consecutive entries a thread loads with one 128 (64, 32)-bit access are contiguous and aligned."""
    n = 1 << logn
    t = n // 32
    for s in range(5, logn):
        g = 16 // (n >> (s + 1))
        v = min(g, 4)
        slots = [slot32(logn, s, tau, r) for tau in range(t) for r in range(g)]
        assert sorted(slots) == list(range(1 << s, 2 << s))
        for tau in range(t):
            for r0 in range(0, g, v):
                base = slot32(logn, s, tau, r0)
                assert base % v == 0
                assert [slot32(logn, s, tau, r0 + i) for i in range(v)] == list(range(base, base + v))
        # lanes of one vector load are contiguous across the polynomial's threads
        assert [slot32(logn, s, tau, 0) for tau in range(t)] == list(range(1 << s, (1 << s) + v * t, v))


@pytest.mark.parametrize("logn", [8, 9, 10])
def test_warp32_ntt_index_is_the_bit_reversal(logn):
    """Element e of thread tau in the pass-1 layout (position 32 tau + e of the bit-reversed order) is the
    reference's coefficient brev5(e) * (n / 32) + brev(tau): what key / matrix gathers use."""
    n = 1 << logn
    brv = lambda i, bits: int(format(i, "0%db" % bits)[::-1], 2)  # noqa: E731
    for tau in range(n // 32):
        for e in range(32):
            assert brv(32 * tau + e, logn) == (brv(e, 5) << (logn - 5)) | brv(tau, logn - 5)
