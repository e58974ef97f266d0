"""World edges of the deforming-plate model on the device (``PlateModel.build_graph``, src/model/plate.py:86-110).

The reference builds the dense ``torch.cdist`` matrix of all node pairs (N^2 fp32: 6 MB at the plate's 1.3 k nodes, 4 TB at
1 M), thresholds it at ``radius``, clears the diagonal, the existing mesh edges, every row that is not an OBSTACLE node and every
column that is not a NORMAL node, and takes ``torch.nonzero``.  ``world_edges`` returns the identical ``(senders, receivers)``
int64 pair -- same set, same row-major order -- from the cell-list kernels of ``libhgn_b200.so`` (csrc/world_edges.cu) in O(N)
memory.  No CPU path: raises ``HgnError`` without a CUDA device.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch

from . import _cabi
from .util import NodeType

RADIUS = 0.03          # src/model/plate.py:87


def world_edges(world_pos: torch.Tensor, node_type: torch.Tensor, mesh_senders: torch.Tensor, mesh_receivers: torch.Tensor,
                radius: float = RADIUS, sender_type: int = int(NodeType.OBSTACLE),
                receiver_type: int = int(NodeType.NORMAL)) -> Tuple[torch.Tensor, torch.Tensor]:
    """``world_senders, world_receivers`` of plate.py:86-110 for ``world_pos [N,3]``, ``node_type [N]`` or ``[N,1]`` and the
    two-way mesh edge list (``util.triangles_to_edges(cells, deform=True)['two_way_connectivity']``)."""
    _cabi.require_cuda(world_pos)
    lib = _cabi.load()
    dev = world_pos.device
    assert world_pos.dim() == 2 and world_pos.shape[1] == 3, "world_pos must be [N, 3]"
    pos = world_pos.detach().to(torch.float32).contiguous()
    types = node_type.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    assert types.numel() == pos.shape[0], "one node type per node"
    ms = mesh_senders.to(device=dev, dtype=torch.int64).contiguous()
    mr = mesh_receivers.to(device=dev, dtype=torch.int64).contiguous()
    assert ms.shape == mr.shape and ms.dim() == 1, "mesh senders / receivers must be 1-D and equally long"
    n, e = pos.shape[0], ms.numel()
    ws_bytes = lib.hgn_world_edges_workspace_bytes(n, e)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    count = ctypes.c_int64(0)
    stream = _cabi.stream_ptr()
    _cabi.check(lib.hgn_world_edges_count(_cabi.ptr(pos), _cabi.ptr(types), n, _cabi.ptr(ms), _cabi.ptr(mr), e, float(radius),
                                          int(sender_type), int(receiver_type), _cabi.ptr(ws), ws_bytes, ctypes.byref(count), stream),
                "hgn_world_edges_count")
    senders = torch.empty(count.value, dtype=torch.int64, device=dev)
    receivers = torch.empty(count.value, dtype=torch.int64, device=dev)
    _cabi.check(lib.hgn_world_edges_emit(_cabi.ptr(pos), _cabi.ptr(types), n, e, float(radius), int(sender_type), _cabi.ptr(ws), ws_bytes,
                                         _cabi.ptr(senders), _cabi.ptr(receivers), count.value, stream), "hgn_world_edges_emit")
    return senders, receivers
