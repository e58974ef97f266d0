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


_edge_storage = os.environ.get('HGN_EDGE_STORAGE', 'reference').lower()


def edge_storage() -> str:
    """Order in which the bf16 processor keeps its edge rows between entry and exit: 'reference' (as given) or
    'receiver_sorted' (stably sorted by receiver once on entry and restored on exit -- experimental, see plan.EdgeStorageOrder)."""
    return _edge_storage


def set_edge_storage(value: str) -> None:
    global _edge_storage
    value = value.lower()
    if value not in ('reference', 'receiver_sorted'):
        raise ValueError("edge storage must be 'reference' or 'receiver_sorted'")
    _edge_storage = value
