"""``build_graph`` of the reference's three system models on the device, with everything that is constant along a trajectory
hoisted out of the step.

The reference rebuilds the whole input graph every step of a rollout and every frame of ``fetch_data``
(src/model/flag.py:65-128, plate.py:69-200, cylinder.py:65-106): a sort-based ``torch.unique`` over the cells, ``F.one_hot`` (which
reads ``max()`` back to the host), relative mesh positions -- all functions of ``cells`` / ``mesh_pos`` / ``node_type`` only -- and, for
the plate, a dense N x N ``torch.cdist``.  A builder below is created once per trajectory from the static fields and called once per
step with the dynamic ones; it uses the model's OWN normalisers (statistics keep accumulating exactly like the reference's) and returns
the same ``MultiGraphWithPos`` -- same index lists bit for bit, features bit for bit (same torch ops in the same order on the same
inputs), which ``tests/test_dropin_gpu.py`` checks against the reference's ``build_graph`` of the same model object.

With ``is_training=False`` a call is a fixed sequence of device kernels with static shapes and no host read-back (flag, cylinder), so a
whole rollout step can be captured in one CUDA graph (``hgn_b200.graphed.rollout_graph``); the plate's world-edge set changes size
from step to step and stays eager (``hgn_b200.world_edges``: cell list, O(N) memory, one count read-back).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import util
from .util import EdgeSet, MultiGraphWithPos, NodeType
from .world_edges import world_edges


def _edge_norm(rel: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(rel.pow(2).sum(-1, keepdim=True))


class FlagGraphBuilder:
    """src/model/flag.py:65-128.  Static: ``cells``, ``mesh_pos``, ``node_type``; dynamic: ``world_pos``, ``prev|world_pos``."""

    def __init__(self, model, static: Dict[str, torch.Tensor]):
        self.model = model
        node_type = static['node_type']
        self.mesh_pos = static['mesh_pos']
        self.one_hot = F.one_hot(torch.flatten(torch.ne(node_type[:, 0], 0)).long())            # flag.py:72-73
        edges = util.triangles_to_edges(static['cells'])
        self.senders, self.receivers = edges['two_way_connectivity']
        self.num_nodes = node_type.shape[0]
        rel_mesh = torch.index_select(self.mesh_pos, 0, self.senders) - torch.index_select(self.mesh_pos, 0, self.receivers)
        self.rel_mesh = torch.cat((rel_mesh, _edge_norm(rel_mesh)), dim=-1)                      # flag.py:82-83, 89-90

    def __call__(self, inputs: Dict[str, torch.Tensor], is_training: bool, node_dynamic: bool = True) -> MultiGraphWithPos:
        """``node_dynamic=False`` skips flag.py:102-115 (max - min incident edge length per node through ``_node_dynamic_normalizer``,
        whose forward always accumulates and therefore compares a device counter on the host): nothing in the learned model reads it."""
        m = self.model
        world_pos, prev = inputs['world_pos'], inputs['prev|world_pos']
        node_features = torch.cat((world_pos - prev, self.one_hot), dim=-1)
        rel_world = torch.index_select(world_pos, 0, self.senders) - torch.index_select(world_pos, 0, self.receivers)
        length = _edge_norm(rel_world)
        edge_features = torch.cat((rel_world, length, self.rel_mesh), dim=-1)
        mesh_edges = EdgeSet(name='mesh_edges', features=m._mesh_edge_normalizer(edge_features, is_training),
                             receivers=self.receivers, senders=self.senders)
        dynamic = None
        if node_dynamic:
            dist = torch.sqrt(rel_world.pow(2).sum(-1))                                          # flag.py:102-113
            spread = (util.unsorted_segment_operation(dist, self.receivers, self.num_nodes, operation='max')
                      - util.unsorted_segment_operation(dist, self.receivers, self.num_nodes, operation='min'))
            dynamic = m._node_dynamic_normalizer(spread)
        return MultiGraphWithPos(
            node_features=[m._node_normalizer(node_features, is_training)], edge_sets=[mesh_edges], target_feature=world_pos,
            mesh_features=self.mesh_pos, model_type=m._model_type, node_dynamic=dynamic,
            unnormalized_edges=EdgeSet(name='mesh_edges', features=edge_features, receivers=self.receivers, senders=self.senders),
            obstacle_nodes=None)


class CylinderGraphBuilder:
    """src/model/cylinder.py:65-106.  Static: ``cells``, ``mesh_pos``, ``node_type``; dynamic: ``velocity``."""

    def __init__(self, model, static: Dict[str, torch.Tensor]):
        self.model = model
        node_types = torch.flatten(static['node_type'][:, 0]).long().clone()
        node_types[node_types == 4] = 1                                                          # cylinder.py:72-75
        node_types[node_types == 5] = 2
        node_types[node_types == 6] = 3
        self.one_hot = F.one_hot(node_types)
        self.mesh_pos = static['mesh_pos']
        edges = util.triangles_to_edges(static['cells'])
        self.senders, self.receivers = edges['two_way_connectivity']
        rel_mesh = torch.index_select(self.mesh_pos, 0, self.senders) - torch.index_select(self.mesh_pos, 0, self.receivers)
        self.edge_features = torch.cat([rel_mesh, _edge_norm(rel_mesh)], dim=-1)

    def __call__(self, inputs: Dict[str, torch.Tensor], is_training: bool) -> MultiGraphWithPos:
        m = self.model
        velocity = inputs['velocity']
        node_features = torch.cat((velocity, self.one_hot), dim=-1)
        mesh_edges = EdgeSet(name='mesh_edges', features=m._mesh_edge_normalizer(self.edge_features, is_training),
                             receivers=self.receivers, senders=self.senders)
        return MultiGraphWithPos(
            node_features=[m._node_normalizer(node_features, is_training)], edge_sets=[mesh_edges], mesh_features=self.mesh_pos,
            target_feature=velocity, model_type=m._model_type,
            unnormalized_edges=EdgeSet(name='mesh_edges', features=self.edge_features, receivers=self.receivers, senders=self.senders),
            node_dynamic=[], obstacle_nodes=None)


class PlateGraphBuilder:
    """src/model/plate.py:69-200.  Static: ``cells``, ``mesh_pos``, ``node_type``; dynamic: ``world_pos``, ``target|world_pos``.
    The dense ``cdist`` / mask / ``nonzero`` block (plate.py:86-110) is the cell-list search of ``hgn_b200.world_edges`` (identical
    index lists, O(N) memory)."""

    def __init__(self, model, static: Dict[str, torch.Tensor]):
        self.model = model
        node_type = static['node_type']
        self.node_type = node_type
        node_types = torch.flatten(node_type[:, 0]).long().clone()
        node_types[node_types == 3] = 2                                                          # plate.py:77-78
        self.one_hot = F.one_hot(node_types)
        self.mesh_pos = static['mesh_pos']
        edges = util.triangles_to_edges(static['cells'], deform=True)
        self.senders, self.receivers = edges['two_way_connectivity']
        rel_mesh = torch.index_select(self.mesh_pos, 0, self.senders) - torch.index_select(self.mesh_pos, 0, self.receivers)
        self.rel_mesh = torch.cat((rel_mesh, _edge_norm(rel_mesh)), dim=-1)
        self.obstacle_nodes = torch.eq(node_type[:, 0], int(NodeType.OBSTACLE))
        self.obstacle_indices = self.obstacle_nodes.nonzero().squeeze()
        self.num_nodes = node_type.shape[0]

    def __call__(self, inputs: Dict[str, torch.Tensor], is_training: bool) -> MultiGraphWithPos:
        m = self.model
        world_pos, target = inputs['world_pos'], inputs['target|world_pos']
        world_senders, world_receivers = world_edges(world_pos, self.node_type, self.senders, self.receivers)
        rel = torch.index_select(world_pos, 0, world_senders) - torch.index_select(world_pos, 0, world_receivers)
        world_edge_features = torch.cat((rel, _edge_norm(rel)), dim=-1)
        world_set = EdgeSet(name='world_edges', features=m._world_edge_normalizer(world_edge_features, is_training),
                            receivers=world_receivers, senders=world_senders)
        rel_world = torch.index_select(world_pos, 0, self.senders) - torch.index_select(world_pos, 0, self.receivers)
        mesh_edge_features = torch.cat((rel_world, _edge_norm(rel_world), self.rel_mesh), dim=-1)
        mesh_set = EdgeSet(name='mesh_edges', features=m._mesh_edge_normalizer(mesh_edge_features, is_training),
                           receivers=self.receivers, senders=self.senders)
        velocities = torch.zeros(self.num_nodes, 3, device=world_pos.device)
        velocities[self.obstacle_nodes] = (torch.index_select(target, 0, self.obstacle_indices)
                                           - torch.index_select(world_pos, 0, self.obstacle_indices))
        node_features = torch.cat((self.one_hot, velocities), dim=-1)
        return MultiGraphWithPos(
            node_features=[m._node_normalizer(node_features, is_training)], edge_sets=[mesh_set, world_set], mesh_features=self.mesh_pos,
            target_feature=world_pos, model_type=m._model_type,
            unnormalized_edges=EdgeSet(name='mesh_edges', features=mesh_edge_features, receivers=self.receivers, senders=self.senders),
            node_dynamic=None, obstacle_nodes=self.obstacle_nodes)


_BUILDERS = {'FlagModel': FlagGraphBuilder, 'PlateModel': PlateGraphBuilder, 'CylinderModel': CylinderGraphBuilder}


def graph_builder(model, static: Dict[str, torch.Tensor]):
    """The builder for a reference system model (by class name: ``FlagModel`` / ``PlateModel`` / ``CylinderModel``)."""
    for cls in type(model).__mro__:
        if cls.__name__ in _BUILDERS:
            return _BUILDERS[cls.__name__](model, static)
    raise TypeError(f'no device graph builder for {type(model).__name__}')
