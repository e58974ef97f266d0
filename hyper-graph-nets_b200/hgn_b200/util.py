"""Mirror of the reference's ``src/util.py`` for the hot path: graph containers, the segment
operation and the mesh-to-edge construction (same names, argument meaning and error behaviour).

``unsorted_segment_operation`` (src/util.py:92-134) runs on the deterministic CSR kernels of
``libhgn_b200.so`` instead of ``torch_scatter``; there is no CPU path.
"""
from __future__ import annotations

import collections
import enum

import torch

from . import _cabi, ops
from .plan import segment_plan, to_device_index

# src/util.py:10-16 -- same field order; `MultiGraphWithPos` keeps the reference's typename 'MultiGraph'
device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
EdgeSet = collections.namedtuple('EdgeSet', ['name', 'features', 'senders', 'receivers'])
MultiGraph = collections.namedtuple('MultiGraph', ['node_features', 'edge_sets'])
MultiGraphWithPos = collections.namedtuple('MultiGraph', ['node_features', 'edge_sets', 'target_feature',
                                                          'mesh_features', 'model_type', 'node_dynamic',
                                                          'unnormalized_edges', 'obstacle_nodes'])


def detach(tensor: torch.Tensor):
    """src/util.py:19-24: a tensor as a numpy array, from either device."""
    return tensor.detach().cpu().numpy()


def read_yaml(config_name: str):
    """src/util.py:38-47: the 'DEFAULT' document of ``configs/<config_name>.yaml`` (path relative to the working directory,
    like the reference); ``None`` after printing the parser's message for a malformed file."""
    import yaml
    with open(f'configs/{config_name}.yaml', 'r') as stream:
        try:
            for document in yaml.safe_load_all(stream):
                if document['name'] == 'DEFAULT':
                    return document
        except yaml.YAMLError as err:
            print(err)
            return None


class NodeType(enum.IntEnum):  # src/util.py:27-35
    NORMAL = 0
    OBSTACLE = 1
    AIRFOIL = 2
    HANDLE = 3
    INFLOW = 4
    OUTFLOW = 5
    WALL_BOUNDARY = 6
    SIZE = 9


_TOPOLOGY_CACHE = collections.OrderedDict()     # id(faces) -> (faces, version, deform, result); a handful of meshes at most


def triangles_to_edges(faces: torch.Tensor, deform: bool = False):
    """Mesh edges from triangles (or, with ``deform``, tetrahedra: edges (0,1),(1,2),(2,3),(3,0) only).

    Bit-exact contract of src/util.py:50-89: unique undirected (max, min) pairs in lexicographic order
    (``torch.unique(dim=0)``), int64, followed by the reversed copies in 'two_way_connectivity'.

    The reference redoes the sort-based ``unique`` on every step of a rollout although the topology never changes along a
    trajectory.  Calling this again with the SAME, unmodified tensor object returns the first result (the
    cache holds the tensor, so its identity cannot be recycled; an in-place change bumps ``_version`` and misses)."""
    hit = _TOPOLOGY_CACHE.get(id(faces))
    if hit is not None and hit[0] is faces and hit[1] == faces._version and hit[2] == deform:
        _TOPOLOGY_CACHE.move_to_end(id(faces))
        return dict(hit[3])
    result = _triangles_to_edges(faces, deform)
    _TOPOLOGY_CACHE[id(faces)] = (faces, faces._version, deform, result)
    while len(_TOPOLOGY_CACHE) > 8:
        _TOPOLOGY_CACHE.popitem(last=False)
    return dict(result)


def _triangles_to_edges(faces: torch.Tensor, deform: bool):
    k = 4 if deform else 3
    pairs = torch.cat([torch.stack((faces[:, i], faces[:, (i + 1) % k]), dim=1) for i in range(k)], dim=0)
    receivers, _ = torch.min(pairs, dim=1)
    senders, _ = torch.max(pairs, dim=1)
    unique_edges = torch.unique(torch.stack((senders, receivers), dim=1), return_inverse=False, return_counts=False, dim=0)
    senders, receivers = torch.unbind(unique_edges, dim=1)
    senders = senders.to(torch.int64)
    receivers = receivers.to(torch.int64)
    two_way = (torch.cat((senders, receivers), dim=0), torch.cat((receivers, senders), dim=0))
    return {'two_way_connectivity': two_way, 'senders': senders, 'receivers': receivers}


_KERNEL_OPS = ('sum', 'mean', 'max', 'min')


def unsorted_segment_operation(data, segment_ids, num_segments, operation):
    """``out[segment_ids[i], ...] (+)= data[i, ...]`` for sum / max / mean / min / std.

    Same contract as src/util.py:92-134 (asserts included): computes in fp32, returns ``data.dtype``,
    empty segments give 0, ``mean`` divides by ``max(count, 1)``, max/min route the gradient to one
    winning element.  ``segment_ids`` may be 1-D (per row) or already expanded to ``data.shape``."""
    assert all([i in data.shape for i in segment_ids.shape]), "segment_ids.shape should be a prefix of data.shape"
    if not torch.cuda.is_available():
        raise _cabi.HgnError("hgn_b200.util.unsorted_segment_operation needs a CUDA device (no CPU fallback)")
    dev = data.device if data.is_cuda else torch.device('cuda')
    data = data.to(dev)
    ids = segment_ids
    if len(ids.shape) != 1:
        assert data.shape == ids.shape, "data.shape and segment_ids.shape should be equal"
        ids = ids.reshape(ids.shape[0], -1)[:, 0]   # the expanded form repeats the row id across features
    if operation not in _KERNEL_OPS and operation != 'std':
        raise Exception('Invalid operation type!')
    ids = to_device_index(ids, dev)
    plan = segment_plan(ids, int(num_segments))
    x = data.float()
    if operation in _KERNEL_OPS:
        (result,) = ops.segment_aggregate(x, plan, (operation,))
    else:  # 'std' (torch_scatter.scatter_std, unbiased): not reachable from any model path
        (mean,) = ops.segment_aggregate(x, plan, ('mean',))
        centered = x - mean.index_select(0, plan.ids)
        (var_sum,) = ops.segment_aggregate(centered * centered, plan, ('sum',))
        count = (plan.rowptr[1:] - plan.rowptr[:-1]).to(x.dtype)
        shape = (-1,) + (1,) * (x.dim() - 1)
        result = (var_sum / ((count - 1).clamp(min=1).view(shape) + 1e-6)).sqrt()
    return result.type(data.dtype)
