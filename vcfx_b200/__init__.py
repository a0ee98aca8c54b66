"""vcfx_b200 — B200-native implementation of VCFX's shared hot path.

Raw VCF text -> record/field scan -> per-sample FORMAT/GT parse -> per-variant reduction ->
order-preserving text, for VCFX_allele_freq_calc, VCFX_allele_counter, VCFX_missing_detector,
VCFX_variant_counter and VCFX_hwe_tester.  The product is ``libvcfx_cuda.so`` (hand-written
sm_100a CUDA behind the C ABI in ``include/vcfx_cuda.h``) plus five drop-in C++ tools in
``vcfx_b200/bin``.  The Python modules here are thin ctypes mirrors used by tests and bench.

There is no CPU fallback: importing :mod:`vcfx_b200.api` without the built CUDA library,
or running an op without a GPU, raises.
"""
__version__ = "0.1.0"
