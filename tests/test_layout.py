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
