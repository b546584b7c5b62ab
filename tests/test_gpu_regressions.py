"""Regressions for the round-1 review findings: concurrent sampler calls on one plan, host staging sized per call,
launches inside a CUDA graph capture."""
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import libsafecrypto_b200 as sc  # noqa: E402

DEV = "cuda:0"


def test_two_streams_share_one_sampler_plan():
    """scgpu_gauss_streams on the AES fast path from two CUDA streams at once (and back to back without a sync): the
    DRBG round keys of a call live in stream-ordered scratch, not in a buffer of the plan."""
    plan = sc.GaussPlan(sc.SAMPLER_CDF, 64, 0, 13.42, 215.0)
    rng = np.random.default_rng(3)
    seeds = [rng.integers(0, 256, size=(4096, 40)).astype(np.uint8) for _ in range(4)]
    d_seeds = [torch.from_numpy(s).to(DEV) for s in seeds]
    outs = [torch.zeros((4096, 512), dtype=torch.int32, device=DEV) for _ in range(4)]
    streams = [torch.cuda.Stream(device=DEV) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(4):
            plan.streams(sc.PRNG_AES_CTR_DRBG, d_seeds[i], 512, outs[i], stream=streams[i % 2])
    torch.cuda.synchronize()
    for i in range(4):
        exp = O.port().gauss_streams(O.SAMPLER_CDF, 64, 0, O.PRNG_AES_CTR_DRBG, 13.42, 215.0, seeds[i][:64], 512)
        assert np.array_equal(outs[i][:64].cpu().numpy(), exp), i
    # whole outputs agree with a serial re-run
    ref = torch.zeros_like(outs[0])
    for i in range(4):
        plan.streams(sc.PRNG_AES_CTR_DRBG, d_seeds[i], 512, ref)
        torch.cuda.synchronize()
        assert torch.equal(ref, outs[i])


def test_host_staging_follows_the_widest_row_of_each_call():
    """A *_host call with padded second-operand rows (b_stride = 2n) after one with b_stride = n on the same plan."""
    q, n = 12289, 512
    w, r = O.tables(q, n, 16)
    plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    rng = np.random.default_rng(4)
    count = 20000
    a = rng.integers(0, q, size=(count, n)).astype(np.int32)
    b = rng.integers(0, q, size=(count, n)).astype(np.int32)
    out = np.zeros_like(a)
    plan.batch_host(sc.OP_PW, out, a, b)
    exp = O.port().ntt_batch(O.REFERENCE, O.OP_PW, n, q, 16, a, b, w, r)
    assert np.array_equal(out, exp)
    bpad = np.zeros((count, 2 * n), dtype=np.int32)
    bpad[:, :n] = b
    out2 = np.zeros_like(a)
    plan.batch_host(sc.OP_PW, out2, a, bpad, b_stride=2 * n)
    assert np.array_equal(out2, exp)
    out3 = np.zeros_like(a)
    plan.polymul_host(out3, a, b)
    assert np.array_equal(out3[:50], O.port().ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, 16, a[:50], b[:50], w, r))


def test_launches_inside_a_cuda_graph_capture():
    """Captured launches keep the static stride (no work-counter slot is baked into the graph); replays agree."""
    q, n = 12289, 512
    w, r = O.tables(q, n, 16)
    plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    g = torch.Generator(device=DEV).manual_seed(2)
    count = 1 << 16                                   # more than one grid-full: the eager launch uses the counter
    a = torch.randint(0, q, (count, n), dtype=torch.int32, device=DEV, generator=g)
    b = torch.randint(0, q, (count, n), dtype=torch.int32, device=DEV, generator=g)
    eager = torch.empty_like(a)
    plan.polymul(eager, a, b)
    torch.cuda.synchronize()
    out = torch.zeros_like(a)
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(device=DEV)
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            plan.polymul(out, a, b, stream=s)
            plan.batch(sc.OP_FWD, out, out, stream=s)
    ref = torch.empty_like(a)
    plan.batch(sc.OP_FWD, ref, eager)
    for _ in range(3):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)


@pytest.mark.parametrize("chunk", ["1", "2", "4"])
def test_chunked_work_claims_cover_every_row_once(chunk, monkeypatch):
    """The persistent kernels claim their groups from a counter, several groups per claim while plenty of work is left
    and single groups at the end (warp32.cuh: Claim).  A ragged batch that goes through both phases: whole output of
    the fused product, the canonical transforms and a variant-exact transform against the checker, for every chunk
    size (SCGPU_CLAIM_CHUNK overrides the launchers' choice)."""
    monkeypatch.setenv("SCGPU_CLAIM_CHUNK", chunk)
    q, n = 7681, 256
    w, r = O.tables(q, n, 16)
    count = 200003                                   # 50001 groups of 4 rows > 10 grid-fulls of 2960 warps
    g = torch.Generator(device=DEV).manual_seed(int(chunk))
    a = torch.randint(0, q, (count, n), dtype=torch.int32, device=DEV, generator=g)
    b = torch.randint(0, q, (count, n), dtype=torch.int32, device=DEV, generator=g)
    out = torch.full_like(a, -1)
    chk = O.ref() if O.ref_available() else O.port()
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    for flags in (0, sc.PLAN_INPUTS_IN_RANGE):
        plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
        if flags:
            plan.set_flags(flags)
        out.fill_(-1)
        plan.polymul(out, a, b)
        exp = chk.ntt_batch(O.REFERENCE, O.OP_POLYMUL, n, q, 16, an, bn, w, r, threads=8)
        assert np.array_equal(out.cpu().numpy(), exp)
        out.fill_(-1)
        plan.ntt_canonical(out, a)
        fwd = chk.ntt_batch(O.REFERENCE, O.OP_FWD, n, q, 16, an, None, w, r, threads=8)
        assert np.array_equal(out.cpu().numpy(), np.mod(fwd, q))
        back = torch.full_like(a, -1)
        plan.ntt_canonical(back, out, inverse=True)
        assert torch.equal(back, a)
    pe = sc.NttPlan(n, q, sc.BARRETT, w, r)
    out.fill_(-1)
    pe.batch(sc.OP_FWD, out, a)
    exp = chk.ntt_batch(O.BARRETT, O.OP_FWD, n, q, 16, an, None, w, r, threads=8)
    assert np.array_equal(out.cpu().numpy(), exp)


@pytest.mark.parametrize("k,l", [(2, 2), (3, 2), (4, 3), (5, 4), (6, 4), (6, 5), (8, 1), (2, 4), (1, 5)])
def test_matvec_one_warp_per_output_row(k, l, monkeypatch):
    """Module product for a 23-bit modulus (32-bit stash): the kernel whose k warps share the transformed vectors of an
    instance group (k_matvec_rows_w32: CTA barriers, work counter per CTA, two matrix rows in flight per warp) against
    the one-warp kernel on a ragged batch larger than one grid-full, and against the checker's composition
    (module_lwe.c:588-748) on the first and last instances."""
    q, n, tw = 8380417, 256, 32
    w, r = O.tables(q, n, tw)
    count = 7001
    g = torch.Generator(device=DEV).manual_seed(100 * k + l)
    A = torch.randint(0, q, (count, k * l, n), dtype=torch.int32, device=DEV, generator=g)
    s = torch.randint(-5, 6, (count, l, n), dtype=torch.int32, device=DEV, generator=g)
    outs = {}
    for flags in (0, sc.PLAN_INPUTS_IN_RANGE):
        plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
        if flags:
            plan.set_flags(flags)
        for mode in ("0", "1"):
            monkeypatch.setenv("SCGPU_MATVEC_ONE_WARP", mode)
            o = torch.full((count, k, n), -1, dtype=torch.int32, device=DEV)
            plan.matvec(o, A, s, k, l)
            torch.cuda.synchronize()
            outs[(flags, mode)] = o
    ref_out = outs[(0, "1")]
    for key, o in outs.items():
        assert torch.equal(o, ref_out), key
    P = O.port()
    sel = np.r_[0:6, count - 6:count]
    An, sn = A[sel].cpu().numpy().reshape(len(sel), k, l, n), s[sel].cpu().numpy()
    sh = P.ntt_batch(O.REFERENCE, O.OP_FWD, n, q, tw, sn.reshape(-1, n), None, w, r).reshape(len(sel), l, n)
    for i in range(k):
        acc = np.zeros((len(sel), n), dtype=np.int64)
        for j in range(l):
            acc += P.ntt_batch(O.REFERENCE, O.OP_PW, n, q, tw, np.ascontiguousarray(An[:, i, j]), np.ascontiguousarray(sh[:, j]))
        t = P.ntt_batch(O.REFERENCE, O.OP_NORMALIZE, n, q, tw, acc.astype(np.int32))
        t = P.ntt_batch(O.REFERENCE, O.OP_INV, n, q, tw, t, None, w, r)
        exp = P.ntt_batch(O.REFERENCE, O.OP_NORMALIZE, n, q, tw, t)
        assert np.array_equal(ref_out[sel][:, i].cpu().numpy(), exp), i


def test_matvec_rows_kernel_small_batches_and_graph_capture(monkeypatch):
    """k_matvec_rows_w32 at the edges: fewer instances than one group, exactly one group, one more; and inside a CUDA
    graph capture, where the launch keeps the static stride (no counter slot baked into the graph)."""
    q, n, tw, k, l = 8380417, 256, 32, 5, 4
    w, r = O.tables(q, n, tw)
    plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    g = torch.Generator(device=DEV).manual_seed(9)
    big = 9001
    A = torch.randint(0, q, (big, k * l, n), dtype=torch.int32, device=DEV, generator=g)
    s = torch.randint(-2, 3, (big, l, n), dtype=torch.int32, device=DEV, generator=g)
    monkeypatch.setenv("SCGPU_MATVEC_ONE_WARP", "1")
    ref = torch.empty((big, k, n), dtype=torch.int32, device=DEV)
    plan.matvec(ref, A, s, k, l)
    torch.cuda.synchronize()
    monkeypatch.setenv("SCGPU_MATVEC_ONE_WARP", "0")
    for count in (1, 3, 4, 5, 8, 9):
        o = torch.full((count, k, n), -1, dtype=torch.int32, device=DEV)
        plan.matvec(o, A[:count], s[:count], k, l)
        torch.cuda.synchronize()
        assert torch.equal(o, ref[:count]), count
    out = torch.zeros_like(ref)
    graph = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream(device=DEV)
    with torch.cuda.stream(st):
        with torch.cuda.graph(graph, stream=st):
            plan.matvec(out, A, s, k, l, stream=st)
    for _ in range(2):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
