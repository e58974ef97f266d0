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


def clear_plan_cache() -> None:
    _plans.clear()
