"""Mirror of ``src/rmp/hierarchical_connector.py`` (+ ``AbstractConnector._get_subgraph``, src/rmp/abstract_connector.py:86-106)
with the per-cluster Python loops replaced by segment kernels and gathers on the device.

The reference runs this on the CPU for every step of an HGN trajectory (``device_0 = 'cpu'``, one ``torch.mean`` /
``torch.max`` / ``_get_subgraph`` call per cluster, lists of Python floats in between) and copies the pieces back to the GPU.
Here the cluster membership is turned once into a ``ClusterPlan`` (member list, cluster id per member, the CSR plan of
``libhgn_b200.so``'s segment kernels) and every step is a handful of launches:

* cluster means of the clustering features and of the node features  -> segment ``mean`` (deterministic CSR kernel)
* hyper-node feature augmentation (cluster size, max distance of a member from the cluster mean in mesh and in world space,
  hierarchical_connector.py:53-69)                                    -> gather, norm, segment ``max``
* ``intra_cluster_to_cluster`` / ``intra_cluster_to_mesh`` / ``inter_cluster`` index lists and relative-position edge features
  (one vectorised ``_get_subgraph`` for all clusters, in the reference's cluster-by-cluster order)

Index lists are bit-exact (``connector_indices`` is pure index arithmetic and also runs without a GPU); features agree with the
reference to fp32 summation-order differences (it averages per cluster with ``torch.mean`` on the CPU).  Same constructor,
``initialize`` and ``run`` signatures, same in-place ``graph.edge_sets.extend`` (hierarchical_connector.py:140-141).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import Tensor

from .. import util
from ..util import EdgeSet, MultiGraph


class ClusterPlan:
    """Device-side form of a ``List[Tensor]`` clustering: members concatenated cluster by cluster, their cluster ids, sizes."""

    def __init__(self, clusters: List[Tensor], device: torch.device):
        self.num_clusters = len(clusters)
        sizes = torch.tensor([int(c.numel()) for c in clusters], dtype=torch.int64)
        self.members = torch.cat([c.reshape(-1).to(torch.int64) for c in clusters]).to(device) if clusters else torch.zeros(0, dtype=torch.int64, device=device)
        self.cluster_of = torch.repeat_interleave(torch.arange(self.num_clusters, dtype=torch.int64), sizes).to(device)
        self.sizes = sizes.to(device)


_PLANS: Dict[int, Tuple[List[Tensor], ClusterPlan]] = {}


def cluster_plan(clusters: List[Tensor], device: torch.device) -> ClusterPlan:
    """One plan per clustering list object (the reference keeps ``self._clusters`` for a whole trajectory,
    remote_message_passing.py:72-77)."""
    hit = _PLANS.get(id(clusters))
    if hit is not None and hit[0] is clusters and hit[1].members.device == device and hit[1].num_clusters == len(clusters):
        return hit[1]
    plan = ClusterPlan(clusters, device)
    if len(_PLANS) > 64:
        _PLANS.clear()
    _PLANS[id(clusters)] = (clusters, plan)
    return plan


def connector_indices(clusters: List[Tensor], neighbors: List[Tensor], num_nodes: int, fully_connect: bool):
    """The three remote edge sets' (senders, receivers), int64, exactly as the reference builds them cluster by cluster
    (hierarchical_connector.py:83-128, 196-211).  Hyper node j has index ``num_nodes + j``.  Pure index arithmetic."""
    sizes = torch.tensor([int(c.numel()) for c in clusters], dtype=torch.int64)
    members = torch.cat([c.reshape(-1).to(torch.int64).cpu() for c in clusters])
    hyper = torch.repeat_interleave(torch.arange(num_nodes, num_nodes + len(clusters), dtype=torch.int64), sizes)
    # _get_subgraph(senders_list = hyper node repeated, receivers_list = cluster): first half hyper -> member ("to mesh"),
    # second half member -> hyper ("to cluster")
    to_cluster = (members, hyper)
    to_mesh = (hyper, members)
    hyper_nodes = torch.arange(num_nodes, num_nodes + len(clusters), dtype=torch.int64)
    if fully_connect or len(clusters) < 4:
        edges = torch.combinations(hyper_nodes, with_replacement=True)
        edges = edges[torch.not_equal(edges[:, 0], edges[:, 1])]
        s, r = edges[:, 0], edges[:, 1]
    else:
        edges = torch.stack([n.reshape(-1).to(torch.int64).cpu() for n in neighbors]) + num_nodes
        s, r = edges[:, 0], edges[:, 1]
    inter = (torch.cat((s, r)), torch.cat((r, s)))
    return {"intra_cluster_to_cluster": to_cluster, "intra_cluster_to_mesh": to_mesh, "inter_cluster": inter}


def _relative_features(target: Tensor, senders: Tensor, receivers: Tensor) -> Tensor:
    """abstract_connector.py:95-102: [world (3), |world|, mesh (rest), |mesh|] of target[senders] - target[receivers]."""
    rel = target.index_select(0, senders) - target.index_select(0, receivers)
    world, mesh = rel[:, :3], rel[:, 3:]
    return torch.cat((world, torch.sqrt(world.pow(2).sum(-1, keepdim=True)), mesh, torch.sqrt(mesh.pow(2).sum(-1, keepdim=True))), dim=-1)


class HierarchicalConnector:
    """Hierarchical remote message passing: hyper nodes, up / down and inter-cluster edges (same API as the reference class)."""

    def __init__(self, fully_connect, noise_scale, hyper_node_features):
        self._intra_normalizer = None
        self._inter_normalizer = None
        self._hyper_normalizer = None
        self._fully_connect = fully_connect
        self._noise_scale = noise_scale
        self._hyper_node_features = hyper_node_features

    def initialize(self, intra, inter, hyper) -> List[str]:
        self._intra_normalizer = intra
        self._inter_normalizer = inter
        self._hyper_normalizer = hyper
        return ['intra_cluster_to_mesh', 'intra_cluster_to_cluster', 'inter_cluster']

    def run(self, graph, clusters: List[Tensor], neighbors: List[Tensor], is_training: bool) -> MultiGraph:
        dev = util.device
        clustering_features = torch.cat((graph.target_feature, graph.mesh_features), dim=1).to(dev).float()
        node_feature = graph.node_features.to(dev)
        model_type = graph.model_type
        if model_type not in ('flag', 'plate'):
            raise Exception("Model type is not specified in RippleNodeConnector.")
        num_nodes = len(graph.node_features)
        plan = cluster_plan(clusters, dev)
        c = plan.num_clusters

        member_features = clustering_features.index_select(0, plan.members)
        clustering_means = util.unsorted_segment_operation(member_features, plan.cluster_of, c, 'mean')
        if is_training and self._noise_scale is not None:
            clustering_means = clustering_means + torch.normal(torch.zeros_like(clustering_means), std=self._noise_scale)
        node_feature_means = util.unsorted_segment_operation(node_feature.index_select(0, plan.members).float(), plan.cluster_of, c, 'mean')
        if self._hyper_node_features:
            away = member_features - clustering_means.index_select(0, plan.cluster_of)
            mesh_d = torch.sqrt(away[:, -3:].pow(2).sum(1))              # hierarchical_connector.py:57-66: last / first three columns
            world_d = torch.sqrt(away[:, :3].pow(2).sum(1))
            spread_mesh = util.unsorted_segment_operation(mesh_d, plan.cluster_of, c, 'max')
            spread_world = util.unsorted_segment_operation(world_d, plan.cluster_of, c, 'max')
            augmentation = torch.stack([plan.sizes.to(spread_mesh.dtype), spread_mesh, spread_world], dim=-1)
            augmentation = self._hyper_normalizer(augmentation, is_training)
            node_feature_means = torch.cat([node_feature_means, augmentation], dim=-1)

        target = torch.cat((clustering_features, clustering_means), dim=0)       # rows [mesh | hyper], like the node list
        idx = connector_indices(clusters, neighbors, num_nodes, bool(self._fully_connect))
        hyper_edges = []
        for name, normalizer in (('intra_cluster_to_cluster', self._intra_normalizer), ('intra_cluster_to_mesh', self._intra_normalizer),
                                 ('inter_cluster', self._inter_normalizer)):
            s, r = (t.to(dev) for t in idx[name])
            hyper_edges.append(EdgeSet(name=name, features=normalizer(_relative_features(target, s, r), is_training), senders=s, receivers=r))

        edge_sets = graph.edge_sets
        edge_sets.extend(hyper_edges)
        return MultiGraph(node_features=[node_feature, node_feature_means], edge_sets=edge_sets)
