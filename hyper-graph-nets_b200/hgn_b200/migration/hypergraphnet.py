"""Mirror of src/migration/hypergraphnet.py: the four-phase HyperGraphNet layer
(mesh -> up -> across -> down), each phase reading the node latents the previous phase wrote."""
from typing import Callable, List

from .graphnet import GraphNet
from ..util import MultiGraph


class HyperGraphNet(GraphNet):
    """Multi-Edge and Multi-Node Interaction Network with residual connections."""

    def __init__(self, model_fn: Callable, output_size: int, message_passing_aggregator: str, edge_sets: List[str]):
        super().__init__(model_fn, output_size, message_passing_aggregator, edge_sets)
        self.hyper_node_model_up = model_fn(output_size)
        self.hyper_node_model_cross = model_fn(output_size)
        self.node_model_down = model_fn(output_size)

    def forward(self, graph: MultiGraph, mask=None) -> MultiGraph:
        done = dict()
        # mesh phase (hypergraphnet.py:29-34)
        for name in ('mesh_edges', 'world_edges'):
            self.perform_edge_updates(graph, name, done)
        self._update_node_features(graph, [done[name] for name in self._present('mesh_edges', 'world_edges')])
        # node -> hyper-node pooling (:37-39)
        self.perform_edge_updates(graph, 'intra_cluster_to_cluster', done)
        self._update_hyper_node_features(graph, [done['intra_cluster_to_cluster']], self.hyper_node_model_up)
        # hyper-node <-> hyper-node (:42-47); 'inter_cluster_world' is never produced by any connector
        for name in ('inter_cluster', 'inter_cluster_world'):
            self.perform_edge_updates(graph, name, done)
        self._update_hyper_node_features(graph, [done[name] for name in self._present('inter_cluster', 'inter_cluster_world')],
                                         self.hyper_node_model_cross)
        # hyper-node -> node broadcast (:50-52)
        self.perform_edge_updates(graph, 'intra_cluster_to_mesh', done)
        self._update_down(graph, [done['intra_cluster_to_mesh']])
        return MultiGraph(graph.node_features, done.values())
