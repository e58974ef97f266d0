"""Mirror of src/migration/multiscalegraphnet.py: mesh, up, 3x across (own MLP each), down, mesh."""
from typing import Callable, List

from torch import nn

from .graphnet import GraphNet
from ..util import MultiGraph


class MultiScaleGraphNet(GraphNet):
    """Multi-Edge and Multi-Node Interaction Network with residual connections."""

    def __init__(self, model_fn: Callable, output_size: int, message_passing_aggregator: str, edge_sets: List[str]):
        super().__init__(model_fn, output_size, message_passing_aggregator, edge_sets)
        self.hyper_node_model_up = model_fn(output_size)
        self.hyper_node_models_cross = nn.ModuleList([model_fn(output_size) for _ in range(3)])
        self.node_model_down = model_fn(output_size)

    def _mesh_phase(self, graph, done):
        for name in ('mesh_edges', 'world_edges'):
            self.perform_edge_updates(graph, name, done)
        self._update_node_features(graph, [done[name] for name in self._present('mesh_edges', 'world_edges')])

    def forward(self, graph: MultiGraph, mask=None) -> MultiGraph:
        done = dict()
        self._mesh_phase(graph, done)                                              # 1 (:27-32)
        self.perform_edge_updates(graph, 'intra_cluster_to_cluster', done)         # up (:35-37)
        self._update_hyper_node_features(graph, [done['intra_cluster_to_cluster']], self.hyper_node_model_up)
        for level in range(3):                                                     # 2, 3, 4 (:40-46)
            for name in ('inter_cluster', 'inter_cluster_world'):
                self.perform_edge_updates(graph, name, done)
            self._update_hyper_node_features(graph, [done[name] for name in self._present('inter_cluster', 'inter_cluster_world')],
                                             self.hyper_node_models_cross[level])
        self.perform_edge_updates(graph, 'intra_cluster_to_mesh', done)            # down (:50-52)
        self._update_down(graph, [done['intra_cluster_to_mesh']])
        self._mesh_phase(graph, done)                                              # 5 (:56-61), original edge latents again
        return MultiGraph(graph.node_features, done.values())
