"""In-trajectory minibatching: mirror of ``MeshSimulator._get_batched`` (src/algorithms/MeshSimulator.py:158-234).

The reference rebuilds every index of every edge set with a Python list comprehension per graph (``.tolist()`` -> ``torch.tensor``:
O(E) interpreter work and a device round trip per batch).  Here the same remap is one ``torch.where`` per edge set on whatever
device the indices live on.  Bit-exact by construction, including the reference's quirk for graphs WITH hyper nodes: an index
``x`` is treated as a hyper node only if ``x >= batch_size * num_nodes`` (not ``x >= num_nodes``), so for ``batch_size > 1`` the
hyper indices of graph i land at ``x + i * num_nodes`` -- the kernels accept arbitrary indices, and
parity with the reference needs the remap unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import Tensor

from .util import EdgeSet, MultiGraph


def _remap(x: Tensor, i: int, batch_size: int, num_nodes: int, num_hyper_nodes: int) -> Tensor:
    hyper_node_offset = batch_size * num_nodes
    x = x.to(torch.int64)
    return torch.where(x < hyper_node_offset, x + i * num_nodes, x + (batch_size - 1) * num_nodes + i * num_hyper_nodes)


def get_batched(data: List[Tuple[MultiGraph, Dict[str, Tensor]]], batch_size: int) -> List[Tuple[MultiGraph, Dict[str, Tensor]]]:
    """Combine the graphs of ``batch_size`` consecutive instances of a trajectory into one graph each (same return structure,
    edge-set order and index dtype -- int64 -- as the reference)."""
    batches = [data[i: i + batch_size] for i in range(0, len(data), batch_size)]
    graph = batches[0][0][0]
    trajectory_attributes = batches[0][0][1].keys()
    edge_names = [e.name for e in graph.edge_sets]

    batched_data = []
    for batch in batches:
        edge_dict = {name: {'snd': [], 'rcv': [], 'features': []} for name in edge_names}
        trajectory_dict = {key: [] for key in trajectory_attributes}
        node_features = []
        for i, (graph, traj) in enumerate(batch):
            num_nodes = tuple(x.shape[0] for x in graph.node_features)
            num_nodes, num_hyper_nodes = num_nodes if len(num_nodes) > 1 else (num_nodes[0], 0)
            node_features.append(graph.node_features)
            for key, value in traj.items():
                trajectory_dict[key].append(value)
            for e in graph.edge_sets:
                edge_dict[e.name]['features'].append(e.features)
                edge_dict[e.name]['snd'].append(_remap(e.senders, i, batch_size, num_nodes, num_hyper_nodes))
                edge_dict[e.name]['rcv'].append(_remap(e.receivers, i, batch_size, num_nodes, num_hyper_nodes))
        new_traj = {key: torch.cat(value, dim=0) for key, value in trajectory_dict.items()}
        all_nodes = [torch.cat(x, dim=0) for x in zip(*node_features)]
        new_graph = MultiGraph(
            node_features=all_nodes,
            edge_sets=[EdgeSet(name=n, features=torch.cat(edge_dict[n]['features'], dim=0), senders=torch.cat(edge_dict[n]['snd'], dim=0),
                               receivers=torch.cat(edge_dict[n]['rcv'], dim=0)) for n in edge_dict.keys()])
        batched_data.append((new_graph, new_traj))
    return batched_data
