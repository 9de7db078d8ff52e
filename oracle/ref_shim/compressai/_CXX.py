"""compressai._CXX stand-in -> oracle/lbic_oracle.c (test infrastructure only)."""
from oracle import native as _native


def pmf_to_quantized_cdf(pmf, precision=16):
    return [int(v) for v in _native.pmf_to_quantized_cdf(pmf, precision)]
