"""The drop-in surface on the GPU, called the way the reference's scheme code calls it: through the function-pointer
table utils_arith_ntt() returns, and through prng_* / create_sampler, with host buffers -- against the unmodified
reference (oracle/_ref/libscref.so) doing the same calls, and against the port."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import libsafecrypto_b200 as sc  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS_SRC = os.path.join(ROOT, "tests", "harness", "table_harness.c")
HARNESS_BIN = os.path.join(ROOT, "tests", "harness", "build", "table_harness")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libscref.so")

# member order of utils_arith_ntt_t (ntt.h:217-297): 18 SINT16-data members, then the 26 SINT32-data ones
MEMBERS_32 = ["modn_32", "muln_32", "sqrn_32", "mul_32_sparse", "mul_32_sparse_16", "mul_32_pointwise",
              "mul_32_pointwise_16", "mul_32_scalar", "fft_32_32", "fft_32_32_large", "fft_32_16", "fft_32_16_large",
              "pwr_32", "invert_32", "div_32", "flip_32", "center_32", "normalize_32", "fwd_ntt_32_32", "inv_ntt_32_32",
              "fwd_ntt_32_32_large", "inv_ntt_32_32_large", "fwd_ntt_32_16", "inv_ntt_32_16", "fwd_ntt_32_16_large",
              "inv_ntt_32_16_large"]
vp = ctypes.c_void_p


def build_harness():
    if not os.path.exists(HARNESS_BIN) or os.path.getmtime(HARNESS_BIN) < os.path.getmtime(HARNESS_SRC):
        os.makedirs(os.path.dirname(HARNESS_BIN), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-Wall", "-o", HARNESS_BIN, HARNESS_SRC, "-ldl", "-lpthread"])
    return HARNESS_BIN


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libscref.so not built")
def test_all_26_members_through_the_table_against_the_reference():
    """tests/harness/table_harness.c: every *_32 member of utils_arith_ntt(v), q in {12289 (n = 512, 1024), 7681,
    8380417}, every live variant, normalize_32 on 3n / 4n / 5n words, in-place inverse transforms, the invert / div
    failure codes -- libscgpu.so against libscref.so word for word, then 8 threads hammering one table."""
    exe = build_harness()
    res = subprocess.run([exe, sc.lib_path(), REF_SO, "--threads", "8", "--rounds", "3"], capture_output=True, text=True,
                         timeout=900)
    print(res.stdout[-3000:])
    print(res.stderr[-3000:])
    assert res.returncode == 0, res.stderr[-2000:]
    last = res.stdout.strip().splitlines()[-1]
    assert last.startswith("SUMMARY") and last.endswith("mismatches=0")
    assert int(last.split("member_calls=")[1].split()[0]) >= 1000
    assert "8 threads concurrently, 0 mismatches" in res.stdout


class Table:
    """utils_arith_ntt(variant) of one library as Python callables."""

    def __init__(self, lib, variant):
        lib.utils_arith_ntt.restype = ctypes.POINTER(vp * 77)
        lib.utils_arith_ntt.argtypes = [ctypes.c_int]
        self.ptrs = lib.utils_arith_ntt(variant).contents
        lib.init_reduce.argtypes = [vp, ctypes.c_size_t, ctypes.c_int32]
        lib.init_reduce.restype = None
        self.lib = lib

    def params(self, n, q):
        buf = ctypes.create_string_buffer(64)
        self.lib.init_reduce(buf, n, q)
        return buf

    def fn(self, name, restype, *argtypes):
        return ctypes.CFUNCTYPE(restype, *argtypes)(self.ptrs[18 + MEMBERS_32.index(name)])


def polymul_via_table(lib, variant, n, q, a, b, w, r):
    """fwd_ntt_32_16(a), fwd_ntt_32_16(b), mul_32_pointwise, inv_ntt_32_16 -- func_ntt-style, one pair."""
    T = Table(lib, variant)
    p = T.params(n, q)
    fwd = T.fn("fwd_ntt_32_16", None, vp, vp, vp, vp)
    pw = T.fn("mul_32_pointwise", None, vp, vp, vp, vp)
    inv = T.fn("inv_ntt_32_16", None, vp, vp, vp, vp, vp)
    a, b, w, r = O.aligned(a), O.aligned(b), O.aligned(w), O.aligned(r)
    fa, fb = O.aligned(np.zeros(n, np.int32)), O.aligned(np.zeros(n, np.int32))
    fwd(fa.ctypes.data, p, a.ctypes.data, w.ctypes.data)
    fwd(fb.ctypes.data, p, b.ctypes.data, w.ctypes.data)
    fa_copy = fa.copy()
    pw(fa.ctypes.data, p, fa.ctypes.data, fb.ctypes.data)
    inv(fa.ctypes.data, p, fa.ctypes.data, w.ctypes.data, r.ctypes.data)
    return fa_copy, fa.copy()


def test_config_c1_single_pair_through_the_count_1_shim():
    """BASELINE config #1 / SURVEY 8d C1: one pair, a[i], b[i] ~ U[0, q) from default_rng(20261017), n = 512,
    q = 12289, through the count = 1 drop-in members and through libscref's same members: identical, for every
    variant; and equal to the schoolbook negacyclic product (congruent to it for Barrett, whose reference
    output is itself not canonical on this composition)."""
    q, n = 12289, 512
    rng = np.random.default_rng(20261017)
    a = rng.integers(0, q, size=n).astype(np.int32)
    b = rng.integers(0, q, size=n).astype(np.int32)
    w, r = O.tables(q, n, 16)
    # schoolbook negacyclic product
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))
    school = full[:n].copy()
    school[:n - 1] -= full[n:]
    school %= q
    for variant in (O.REFERENCE, O.BARRETT, O.FP, O.AVX):
        fwd_g, out_g = polymul_via_table(sc.lib(), variant, n, q, a, b, w, r)
        if variant == O.BARRETT:
            # unnormalised x unnormalised pointwise products leave Barrett (k = 30) only congruent: the
            # reference's own output is not canonical here, and it is matched word for word below
            assert np.array_equal(out_g.astype(np.int64) % q, school), variant
        else:
            assert np.array_equal(out_g, school), variant
        assert np.array_equal(fwd_g, O.port().ntt_batch(variant, O.OP_FWD, n, q, 16, a, None, w, r)[0])
        if O.ref_available():
            fwd_r, out_r = polymul_via_table(O.ref().lib, variant, n, q, a, b, w, r)
            assert np.array_equal(fwd_g, fwd_r) and np.array_equal(out_g, out_r), variant


# ---- PRNG front end ------------------------------------------------------------------------------------------------

def bind_prng(L):
    L.prng_create.restype = vp
    L.prng_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
    L.prng_set_entropy.argtypes = [vp, vp, ctypes.c_size_t]
    L.prng_init.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
    L.prng_destroy.argtypes = [vp]
    L.prng_reset.argtypes = [vp]
    L.prng_reset.restype = None
    for name, res in (("prng_32", ctypes.c_uint32), ("prng_64", ctypes.c_uint64), ("prng_16", ctypes.c_uint16),
                      ("prng_8", ctypes.c_uint8), ("prng_bit", ctypes.c_int32), ("prng_float", ctypes.c_float),
                      ("prng_double", ctypes.c_double), ("prng_get_csprng_bytes", ctypes.c_uint64),
                      ("prng_get_out_bytes", ctypes.c_uint64)):
        getattr(L, name).restype = res
        getattr(L, name).argtypes = [vp]
    L.prng_var.restype = ctypes.c_uint32
    L.prng_var.argtypes = [vp, ctypes.c_size_t]
    L.prng_mem.argtypes = [vp, vp, ctypes.c_int32]
    L.prng_mem.restype = ctypes.c_int32

    class U128(ctypes.Structure):                      # unsigned __int128 comes back in RAX:RDX like a 2 x u64 struct
        _fields_ = [("lo", ctypes.c_uint64), ("hi", ctypes.c_uint64)]
    L.prng_128.restype = U128
    L.prng_128.argtypes = [vp]
    return L


def run_script(L, ctx, script):
    """The (kind, arg) script of oracle/sc_oracle.h on a live prng_ctx_t of library L."""
    out = []
    for kind, arg in script:
        if kind == 32:
            out.append(L.prng_32(ctx))
        elif kind == 64:
            x = L.prng_64(ctx)
            out += [x >> 32, x & 0xFFFFFFFF]
        elif kind == 8:
            out.append(L.prng_8(ctx))
        elif kind == 1:
            out.append(L.prng_bit(ctx) & 0xFFFFFFFF)
        elif kind == 16:
            out.append(L.prng_16(ctx))
        elif kind == 128:
            x = L.prng_128(ctx)
            out += [x.hi >> 32, x.hi & 0xFFFFFFFF, x.lo >> 32, x.lo & 0xFFFFFFFF]
        elif kind == 2:
            out.append(int(np.float32(L.prng_float(ctx)).view(np.uint32)))
        elif kind == 3:
            bits = int(np.float64(L.prng_double(ctx)).view(np.uint64))
            out += [bits & 0xFFFFFFFF, bits >> 32]
        elif kind == 4:
            buf = np.zeros((arg + 3) // 4, dtype=np.uint32)
            assert L.prng_mem(ctx, buf.ctypes.data, arg) == 0
            out += [int(x) for x in buf]
        elif kind == 5:
            L.prng_reset(ctx)
        elif kind == 6:
            out += [L.prng_get_csprng_bytes(ctx) & 0xFFFFFFFF, L.prng_get_out_bytes(ctx) & 0xFFFFFFFF]
        else:
            out.append(L.prng_var(ctx, arg))
    return np.array(out, dtype=np.uint64).astype(np.uint32)


def random_script(rng, prng, length):
    script = []
    for _ in range(length):
        k = int(rng.choice([32, 64, 8, 1, 16, 0, 128, 2, 3, 4, 4, 6] + ([5] if prng == O.PRNG_AES_CTR_DRBG else [])))
        arg = 0
        if k == 0:
            arg = int(rng.integers(1, 33))
        if k == 4:
            arg = int(rng.choice([1, 7, 64, 65, 512, 1000, 4096]))
        script.append((k, arg))
    return script


@pytest.mark.parametrize("prng", [O.PRNG_CHACHA, O.PRNG_AES_CTR_DRBG])
def test_prng_front_end_scripts(prng):
    """prng_32/64/128/16/8/bit/var/float/double/mem/reset and the byte counters of the drop-in context, in random
    interleavings (prng_mem draws behind the 4096-word pool; prng_reset leaves the DRBG's stale buffer), against the
    port and the compiled reference.  User-provided entropy, small and default reseed periods."""
    L = bind_prng(sc.lib())
    rng = np.random.default_rng(77 + prng)
    for trial in range(6):
        seed = rng.integers(0, 256, size=int(rng.integers(36, 80))).astype(np.uint8)
        script = random_script(rng, prng, 250)
        period = int(rng.choice([0x00100000, 64, 4096, 0x10000]))
        ctx = L.prng_create(5, prng, 0, period)
        assert ctx and L.prng_set_entropy(ctx, seed.ctypes.data, seed.size) == 0
        assert L.prng_init(ctx, b"SAFEcrypto nonce", 16) == 0
        got = run_script(L, ctx, script)
        L.prng_destroy(ctx)
        exp = O.port().prng_script(prng, seed, script, period)
        assert np.array_equal(got, exp), (trial, np.nonzero(got != exp)[0][:4])
        if O.ref_available():
            assert np.array_equal(exp, O.ref().prng_script(prng, seed, script, period))


def test_prng_front_end_golden_fixture():
    """The same front-end script against outputs of the compiled reference committed in tests/golden/golden_v3.npz."""
    G3 = np.load(os.path.join(ROOT, "tests", "golden", "golden_v3.npz"))
    L = bind_prng(sc.lib())
    script = [tuple(int(v) for v in row) for row in G3["prng_script"]]
    seed = np.ascontiguousarray(G3["prng_script_seed"])
    for prng, key, scr in ((O.PRNG_AES_CTR_DRBG, "prng_script_aes", script),
                           (O.PRNG_CHACHA, "prng_script_chacha", [(k, a) for k, a in script if k != 5])):
        ctx = L.prng_create(5, prng, 0, 4096)
        assert ctx and L.prng_set_entropy(ctx, seed.ctypes.data, seed.size) == 0 and L.prng_init(ctx, b"SAFEcrypto nonce", 16) == 0
        got = run_script(L, ctx, scr)
        L.prng_destroy(ctx)
        assert np.array_equal(got, G3[key])


ENTROPY_CB = ctypes.CFUNCTYPE(None, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint8))


def counting_callback():
    """A stateful deterministic entropy source: byte k of everything it ever handed out is f(k)."""
    state = {"pos": 0, "calls": []}

    def cb(n, data):
        for i in range(n):
            z = (state["pos"] + i + 1) * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF       # splitmix64 of the byte index
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
            data[i] = (z ^ (z >> 31)) & 0xFF
        state["pos"] += n
        state["calls"].append(n)
    return ENTROPY_CB(cb), state


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libscref.so not built")
@pytest.mark.parametrize("prng,period,nbytes", [(O.PRNG_CHACHA, 4096, 80 * 4096), (O.PRNG_AES_CTR_DRBG, 0x1000 << 4, 9 << 20)])
def test_callback_entropy_is_fresh_at_every_reseed(prng, period, nbytes):
    """SC_ENTROPY_CALLBACK: the reference pulls new entropy at every reseed (chacha20_csprng.c:21-29,
    ctr_drbg.c:128-147).  ChaCha20 with a 4 KiB period crosses 80 reseeds -- more than any pre-drawn ring of the
    previous implementation held -- and must still equal the reference fed by the same stateful source; the DRBG
    crosses its first two reseeds.  Every reseed consumes new callback bytes, in the reference's request sizes."""
    G, R = bind_prng(sc.lib()), bind_prng(O.ref().lib)
    streams = []
    for L in (G, R):
        cb, state = counting_callback()
        L.prng_set_entropy_callback.argtypes = [ENTROPY_CB]
        assert L.prng_set_entropy_callback(cb) == 0
        ctx = L.prng_create(4, prng, 0, period)               # SC_ENTROPY_CALLBACK
        assert ctx and L.prng_init(ctx, b"SAFEcrypto nonce", 16) == 0
        buf = np.zeros(nbytes, dtype=np.uint8)
        chunk = 1 << 16
        for off in range(0, nbytes, chunk):
            assert L.prng_mem(ctx, buf[off:].ctypes.data, min(chunk, nbytes - off)) == 0
        words = [L.prng_32(ctx) for _ in range(64)]
        L.prng_destroy(ctx)
        streams.append((buf, words, state))
    (gb, gw, gs), (rb, rw, rs) = streams
    assert np.array_equal(gb, rb) and gw == rw
    reseeds = nbytes // period if prng == O.PRNG_CHACHA else nbytes // 1024 // 0x1000
    per_seed = 40 if prng == O.PRNG_CHACHA else 36
    assert rs["pos"] >= (reseeds + 1) * per_seed and gs["pos"] >= rs["pos"]
    # same request pattern: 40-byte requests for ChaCha20, 4 + 32 for the DRBG
    assert set(gs["calls"]) == set(rs["calls"]) == ({40} if prng == O.PRNG_CHACHA else {4, 32})
    # the keystream never repeats: no 4 KiB block of the output occurs twice
    blocks = {gb[i:i + 4096].tobytes() for i in range(0, min(nbytes, 2 << 20), 4096)}
    assert len(blocks) == min(nbytes, 2 << 20) // 4096


def test_os_entropy_contexts_differ_and_run():
    """SC_ENTROPY_DEV_URANDOM: two contexts give different streams, reseeds draw new bytes (statistical smoke)."""
    L = bind_prng(sc.lib())
    outs = []
    for _ in range(2):
        ctx = L.prng_create(2, O.PRNG_CHACHA, 0, 1024)
        assert ctx and L.prng_init(ctx, b"SAFEcrypto nonce", 16) == 0
        buf = np.zeros(1 << 16, dtype=np.uint8)
        assert L.prng_mem(ctx, buf.ctypes.data, buf.size) == 0
        outs.append(buf)
        L.prng_destroy(ctx)
    assert not np.array_equal(outs[0], outs[1])
    assert abs(float(np.unpackbits(outs[0]).mean()) - 0.5) < 0.01
    assert len({outs[0][i:i + 1024].tobytes() for i in range(0, 1 << 16, 1024)}) == 64
