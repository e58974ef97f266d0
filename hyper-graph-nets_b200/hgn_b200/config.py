"""Process-wide switches that need no change to the reference's config schema."""
import os

_precision = os.environ.get('HGN_B200_PRECISION', 'fp32').lower()


def precision() -> str:
    """'fp32' (parity mode) or 'bf16' (tcgen05 throughput mode)."""
    return _precision


def set_precision(value: str) -> None:
    global _precision
    value = value.lower()
    if value not in ('fp32', 'bf16'):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _precision = value
