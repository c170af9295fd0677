"""thaler_study_b200 -- B200-native sum-check prover engine behind the reference's SumCheckPolynomial trait.

The compute lives in ``libsumcheck_b200.so`` (hand-written sm_100a CUDA + a C++ host protocol layer, C ABI in
``include/sumcheck_b200.h``).  This package is a thin ctypes mirror of the reference's types for tests, benches
and the multi-GPU driver.  There is no CPU fallback: importing fails loudly if the library is not built, and every
compute call fails with SCB_ECUDA without a CUDA device.
"""
from ._lib import (  # noqa: F401
    SO_PATH,
    NoPolySet,
    ProverClaimMismatch,
    ScbError,
    get_option,
    lib,
    option_names,
    options_from_env,
    reset_options,
    set_option,
)
from .api import (  # noqa: F401
    KIND_GKR_W,
    KIND_MATMUL_G,
    KIND_PRODUCT,
    KIND_TRIANGLE_G,
    DenseMultilinearExtension,
    Field,
    GkrW,
    MatMulG,
    ProductMLE,
    Prover,
    SparsePolynomial,
    SumCheckPolynomial,
    Transcript,
    TriangleG,
    Verifier,
    cti_multilinear_from_evaluations,
    device_count,
    evals_to_univariate,
    generate_transcript,
    launch_count,
    set_stream,
    synchronize,
    verify_transcript,
    vsbw_multilinear_from_evaluations,
)
