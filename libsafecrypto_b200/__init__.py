"""libsafecrypto_b200 -- B200-native lattice hot path behind libsafecrypto's own C interface.

The product is the C-ABI shared library ``libscgpu.so`` (sources in ``csrc/``, interface in
``include/scgpu.h`` and ``include/scgpu_dropin.h``).  This package is only the thin Python
harness the tests and ``bench.py`` use to drive that library with torch-owned device memory and
streams: every call below forwards to one exported C symbol.  There is no CPU implementation
here and no fallback: if the library is missing or a CUDA call fails, an exception is raised.
"""
from . import binding  # noqa: F401
from .binding import (  # noqa: F401
    ScgpuError, lib, lib_path, NttPlan, NttPlanSet, GaussPlan, rand_matrix, make_params, launch_count, int_peak_gops,
    REFERENCE, BARRETT, FP, AVX, SOLINAS_7681, SOLINAS_8380417,
    OP_FWD, OP_INV, OP_FWD_LARGE, OP_INV_LARGE, OP_FFT, OP_FFT_LARGE, OP_PW, OP_PW16, OP_NORMALIZE,
    OP_CENTER, OP_POLYMUL, OP_TRIPLE16, OP_MODN, OP_MULN, OP_SQRN, OP_FLIP, OP_INVERT, OP_DIV,
    OP_PWR, OP_SCALAR, OP_SPARSE32, OP_SPARSE16,
    PRNG_AES_CTR_DRBG, PRNG_CHACHA, PLAN_INPUTS_IN_RANGE, SAMPLER_CDF, SAMPLER_KNUTH_YAO, SAMPLER_BERNOULLI, SAMPLER_KNUTH_YAO_FAST,
    NORMAL_SAMPLES, BLINDING_SAMPLES, SHUFFLE_SAMPLES,
)

__all__ = [n for n in dir() if not n.startswith("_")]
