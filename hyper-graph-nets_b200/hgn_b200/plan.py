"""Topology plans: int32 index copies and receiver-/sender-sorted CSR views of an index vector.

The reference re-expands ``receivers`` to an ``[E,128]`` int64 index on every aggregation
(src/util.py:105-110) and moves the index tensors to the device on every use (graphnet.py:25-26).
Mesh topology is constant along a trajectory, so here each distinct index tensor gets one cached
``SegmentPlan``: the int32 copy used as gather index by the MLP tile kernels, and -- built on first
need by ``hgn_csr_build`` (stable sort on the device) -- the CSR view used by the aggregation
forward and by the deterministic gather backward.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import _cabi

_CACHE_CAPACITY = 256


class SegmentPlan:
    """``ids`` with values in ``[0, num_segments)``.

    ``ids32``  int32 copy (gather index).
    ``perm``   element ids grouped by segment, ascending id inside a segment (stable).
    ``rowptr`` ``rowptr[s]:rowptr[s+1]`` delimits segment ``s`` inside ``perm``."""

    __slots__ = ("ids", "key_tensor", "num_segments", "num_elements", "_perm", "_rowptr", "ids32", "version")

    def __init__(self, ids: torch.Tensor, num_segments: int):
        _cabi.require_cuda(ids)
        if ids.dtype != torch.int64:
            ids = ids.to(torch.int64)
        self.ids = ids.contiguous()
        self.key_tensor = None
        self.version = ids._version
        self.num_segments = int(num_segments)
        self.num_elements = self.ids.numel()
        self.ids32 = self.ids.to(torch.int32) if self.num_elements else torch.zeros(1, dtype=torch.int32, device=ids.device)
        self._perm = None
        self._rowptr = None

    def _build_csr(self) -> None:
        lib = _cabi.load()
        E, dev = self.num_elements, self.ids.device
        self._perm = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        self._rowptr = torch.empty(self.num_segments + 1, dtype=torch.int32, device=dev)
        ws_bytes = lib.hgn_csr_workspace_bytes(E, self.num_segments)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(lib.hgn_csr_build(self.ids.data_ptr(), E, self.num_segments, self._perm.data_ptr(),
                                          self._rowptr.data_ptr(), None, ws.data_ptr(), ws_bytes, _cabi.stream_ptr()),
                        "hgn_csr_build")

    @property
    def perm(self) -> torch.Tensor:
        if self._perm is None:
            self._build_csr()
        return self._perm

    @property
    def rowptr(self) -> torch.Tensor:
        if self._rowptr is None:
            self._build_csr()
        return self._rowptr


_plans: "OrderedDict[tuple, object]" = OrderedDict()


def _remember(key, value) -> None:
    _plans[key] = value
    if len(_plans) > _CACHE_CAPACITY:
        _plans.popitem(last=False)


def segment_plan(ids: torch.Tensor, num_segments: int) -> SegmentPlan:
    """Cached plan for a device index tensor.  The key is the tensor's storage identity (+ in-place
    version counter); the cache entry keeps the tensor alive so its address cannot be recycled."""
    key = ("plan", ids.data_ptr(), ids.numel(), int(num_segments), ids.device.index, ids.dtype)
    hit = _plans.get(key)
    if hit is not None and hit.version == ids._version:
        _plans.move_to_end(key)
        return hit
    plan = SegmentPlan(ids, num_segments)
    plan.key_tensor = ids
    plan.version = ids._version
    _remember(key, plan)
    return plan


def to_device_index(ids: torch.Tensor, device: torch.device) -> torch.Tensor:
    """Device copy of an index tensor, made once per source tensor."""
    if ids.device == device:
        return ids
    key = ("h2d", ids.data_ptr(), ids.numel(), str(device), ids.dtype)
    hit = _plans.get(key)
    if hit is not None and hit[2] == ids._version:
        _plans.move_to_end(key)
        return hit[1]
    dev_ids = ids.to(device)
    _remember(key, (ids, dev_ids, ids._version))
    return dev_ids


class EdgeStorageOrder:
    """Receiver-sorted STORAGE order of one edge set (experimental, ``HGN_EDGE_STORAGE=receiver_sorted``; DESIGN.md s8 round-2 plan).

    The reference's edge order (util.py:60-69: unique (max, min) pairs, then the reversed copies) is half sender-sorted and half
    receiver-sorted.  Keeping the rows stably sorted by receiver INSIDE the processor makes every receiver's edges contiguous: the
    ``Pr[r]`` / ``grad_agg[r]`` row gathers of the edge kernels then touch ~6 distinct rows per warp instruction instead of ~17
    (scripts/analysis_gather_locality.py), and the aggregation becomes a reduction over consecutive rows (the precondition for
    fusing it into the edge kernels' epilogues).  The sort is stable, so within a segment the rows keep their ascending reference
    order: the receiver-side CSR kernels sum them in the same order, so the forward results are bitwise those of the reference
    order (the sender-side sums of the backward see another order: gradients agree to fp32 / bf16 rounding).

    ``perm[k]`` = reference id of the k-th stored row; ``inverse[i]`` = storage position of reference row i.  One instance per
    receivers tensor (cached like the segment plans); ``senders`` / ``receivers`` are the permuted int64 index tensors, created once so
    that their own segment plans stay cached across steps."""

    def __init__(self, senders: torch.Tensor, receivers: torch.Tensor):
        self.perm = torch.argsort(receivers, stable=True)
        self.inverse = torch.empty_like(self.perm)
        self.inverse[self.perm] = torch.arange(self.perm.numel(), device=self.perm.device, dtype=self.perm.dtype)
        self.senders = senders.index_select(0, self.perm).contiguous()
        self.receivers = receivers.index_select(0, self.perm).contiguous()
        self.version = (senders._version, receivers._version)

    def store(self, features: torch.Tensor) -> torch.Tensor:
        """Rows in storage order (differentiable: the backward scatters every gradient row to exactly one place)."""
        return features.index_select(0, self.perm)

    def restore(self, features: torch.Tensor) -> torch.Tensor:
        """Rows back in the reference order."""
        return features.index_select(0, self.inverse)


def edge_storage_order(senders: torch.Tensor, receivers: torch.Tensor) -> EdgeStorageOrder:
    """Cached per (senders, receivers) tensor pair, like ``segment_plan``."""
    key = ("order", senders.data_ptr(), receivers.data_ptr(), receivers.numel(), str(receivers.device))
    hit = _plans.get(key)
    if hit is not None and hit[2].version == (senders._version, receivers._version):
        _plans.move_to_end(key)
        return hit[2]
    order = EdgeStorageOrder(senders, receivers)
    _remember(key, (senders, receivers, order))
    return order


def clear_plan_cache() -> None:
    _plans.clear()
