"""create_rand_product_{16,32}_csprng with the matrix sampled on the device (scgpu_rand_product_csprng_batch) against
the compiled reference calling its own module_lwe.c functions, one CSPRNG per instance (oracle/ref_driver.c:
ref_rand_product), and against committed fixtures generated from it (tests/golden/golden_v3.npz)."""
import os

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import libsafecrypto_b200 as sc  # noqa: E402
from libsafecrypto_b200 import binding as B  # noqa: E402

DEV = "cuda:0"
G3 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v3.npz")


def gpu_product(q, n, tw, k, l, q_bits, prng, seeds, y, transpose, variant=sc.REFERENCE, host=False):
    w, r = O.tables(q, n, tw)
    plan = sc.NttPlan(n, q, variant, w, r)
    count = seeds.shape[0]
    if host:
        t = np.zeros((count, k, n), dtype=np.int32)
        plan.rand_product_host(t, np.ascontiguousarray(y), np.ascontiguousarray(seeds), prng, q_bits, k, l, transpose)
        return t
    t = torch.full((count, k, n), -7, dtype=torch.int32, device=DEV)
    plan.rand_product(t, torch.from_numpy(y).to(DEV), torch.from_numpy(seeds).to(DEV), prng, q_bits, k, l, transpose)
    torch.cuda.synchronize()
    return t.cpu().numpy()


def gpu_matrix(q, n, k, l, q_bits, prng, seeds, transpose):
    A = torch.full((seeds.shape[0], k, l, n), -1, dtype=torch.int32, device=DEV)
    B.rand_matrix(A, torch.from_numpy(seeds).to(DEV), prng, q, q_bits, n, k, l, transpose)
    torch.cuda.synchronize()
    return A.cpu().numpy()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libscref.so not built")
@pytest.mark.parametrize("prng", [O.PRNG_CHACHA, O.PRNG_AES_CTR_DRBG])
def test_matrix_is_what_the_reference_draws(prng):
    """uniform_random_ring_q_csprng ring by ring (512 bytes of prng_mem per 256 coefficients), Kyber and Dilithium
    shapes; the transposed destination is ring j k + i -> A[i][j]."""
    rng = np.random.default_rng(3 + prng)
    for q, n, tw, k, l, q_bits in ((7681, 256, 16, 3, 3, 13), (7681, 256, 16, 2, 2, 13), (8380417, 256, 32, 5, 4, 23), (7681, 256, 16, 2, 3, 13)):
        seeds = rng.integers(0, 256, size=(21, 32)).astype(np.uint8)
        y = np.zeros((21, l, n), dtype=np.int32)
        w, r = O.tables(q, n, tw)
        _, A = O.ref().rand_product(tw, O.REFERENCE, n, q, q_bits, k, l, False, prng, seeds, y, w, r, want_matrix=True)
        got = gpu_matrix(q, n, k, l, q_bits, prng, seeds, False)
        assert np.array_equal(got.reshape(21, k * l, n), A), (q, k, l)
        got_t = gpu_matrix(q, n, k, l, q_bits, prng, seeds, True)
        for i in range(k):
            for j in range(l):
                assert np.array_equal(got_t[:, i, j], A[:, j * k + i])
    with pytest.raises(sc.ScgpuError):       # n = 512: the reference's ring sampler never advances its output pointer
        gpu_matrix(12289, 512, 2, 2, 14, prng, seeds, False)


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libscref.so not built")
@pytest.mark.parametrize("prng", [O.PRNG_CHACHA, O.PRNG_AES_CTR_DRBG])
@pytest.mark.parametrize("k", [2, 3, 4])
def test_kyber_products_against_the_reference(prng, k):
    """q = 7681, n = 256, k = l: keygen's non-transposed product and encapsulation's transposed one
    (module_lwe.c:1233-1237, 1338-1340), reference / avx / barrett variants of the reference all give these residues."""
    q, n = 7681, 256
    rng = np.random.default_rng(10 * k + prng)
    seeds = rng.integers(0, 256, size=(45, 32)).astype(np.uint8)
    y = rng.integers(-4, 5, size=(45, k, n)).astype(np.int32)
    w, r = O.tables(q, n, 16)
    for transpose in (False, True):
        got = gpu_product(q, n, 16, k, k, 13, prng, seeds, y, transpose)
        for variant in (O.REFERENCE, O.AVX, O.BARRETT):
            exp = O.ref().rand_product(16, variant, n, q, 13, k, k, transpose, prng, seeds, y, w, r)
            assert np.array_equal(got, exp), (transpose, variant)
        assert got.min() >= 0 and got.max() < q
    assert np.array_equal(gpu_product(q, n, 16, k, k, 13, prng, seeds, y, False, host=True),
                          gpu_product(q, n, 16, k, k, 13, prng, seeds, y, False))


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libscref.so not built")
@pytest.mark.parametrize("prng", [O.PRNG_CHACHA, O.PRNG_AES_CTR_DRBG])
@pytest.mark.parametrize("k,l", [(5, 4), (6, 5), (3, 2)])
def test_dilithium_product_against_the_reference(prng, k, l):
    """q = 8380417, n = 256, the (k, l) of Dilithium's parameter sets incl. the largest, k = 6, l = 5; 32-bit tables
    (dilithium.c:877-887): the matrix coefficients are 16-bit values (uniform_random_ring_q_csprng reads UINT16
    whatever q_bits is) -- reproduced, not corrected."""
    q, n = 8380417, 256
    rng = np.random.default_rng(77 + prng + 10 * k)
    seeds = rng.integers(0, 256, size=(33, 32)).astype(np.uint8)
    y = rng.integers(-5, 6, size=(33, l, n)).astype(np.int32)
    w, r = O.tables(q, n, 32)
    got = gpu_product(q, n, 32, k, l, 23, prng, seeds, y, False)
    for variant in (O.REFERENCE, O.FP):
        assert np.array_equal(got, O.ref().rand_product(32, variant, n, q, 23, k, l, False, prng, seeds, y, w, r)), variant
    wq, rq = O.tables(q, n, 32)
    plan = sc.NttPlan(n, q, sc.REFERENCE, wq, rq)
    with pytest.raises(sc.ScgpuError):                       # the reference's own transposed 32-bit branch is broken
        plan.rand_product(torch.zeros((1, k, n), dtype=torch.int32, device=DEV), torch.zeros((1, l, n), dtype=torch.int32, device=DEV),
                          torch.zeros((1, 32), dtype=torch.uint8, device=DEV), O.PRNG_CHACHA, 23, k, l, True)


def test_large_batch_spans_several_l2_chunks_and_matches_matvec():
    """2^15 Kyber k = 3 instances (several 48 MB matrix chunks): equal to the mat-vec entry point fed with the matrix
    the generator entry point returns."""
    q, n, k = 7681, 256, 3
    g = torch.Generator(device=DEV).manual_seed(5)
    count = 1 << 15
    seeds = torch.randint(0, 256, (count, 32), dtype=torch.uint8, device=DEV, generator=g)
    y = torch.randint(-4, 5, (count, k, n), dtype=torch.int32, device=DEV, generator=g)
    w, r = O.tables(q, n, 16)
    plan = sc.NttPlan(n, q, sc.REFERENCE, w, r)
    t1 = torch.empty((count, k, n), dtype=torch.int32, device=DEV)
    plan.rand_product(t1, y, seeds, O.PRNG_CHACHA, 13, k, k)
    A = torch.empty((count, k, k, n), dtype=torch.int32, device=DEV)
    B.rand_matrix(A, seeds, O.PRNG_CHACHA, q, 13, n, k, k)
    t2 = torch.empty_like(t1)
    plan.matvec(t2, A, y, k, k)
    torch.cuda.synchronize()
    assert torch.equal(t1, t2)
    assert int(A.max()) < q and int(A.min()) >= 0


def test_golden_rand_products():
    """Fixtures generated from the compiled reference (tests/golden/make_golden.py), for boxes without libscref."""
    G = np.load(G3)
    for name, (q, n, tw, k, l, q_bits) in (("kyber3", (7681, 256, 16, 3, 3, 13)), ("dil", (8380417, 256, 32, 5, 4, 23))):
        for pname, prng in (("chacha", O.PRNG_CHACHA), ("aes", O.PRNG_AES_CTR_DRBG)):
            seeds, y = G["rp_seeds"], G["rp_y_%s" % name]
            assert np.array_equal(gpu_product(q, n, tw, k, l, q_bits, prng, seeds, y, False), G["rp_t_%s_%s" % (name, pname)])
            if name == "kyber3":
                assert np.array_equal(gpu_product(q, n, tw, k, l, q_bits, prng, seeds, y, True), G["rp_tT_%s_%s" % (name, pname)])
