"""Mirror of src/migration/heterographnet.py: base schedule, ONE aggregation over every edge set,
separate MLPs for the mesh rows and the hyper rows."""
from typing import Callable, List

from .graphnet import GraphNet
from ..util import EdgeSet, MultiGraph


class HeteroGraphNet(GraphNet):
    """Multi-Edge and Multi-Node Interaction Network with residual connections."""

    def __init__(self, model_fn: Callable, output_size: int, message_passing_aggregator: str, edge_sets: List[str]):
        super().__init__(model_fn, output_size, message_passing_aggregator, edge_sets)
        self.hyper_node_model_cross = model_fn(output_size)

    def _update_node_features(self, graph: MultiGraph, edge_sets: List[EdgeSet]):
        # both updates read the latents from BEFORE this layer's node update and the SAME aggregation over all N + C rows
        # (heterographnet.py:17-33: one `aggregation` call, features[:N] / features[N:]): aggregate once, two fused MLP calls on
        # the two row ranges of the same aggregate tensors
        v_rows = sum(t.shape[0] for t in graph.node_features)
        aggregates = self._aggregates(edge_sets, v_rows, graph.node_features[0].device)
        new_mesh = self._fused_node_update(graph, edge_sets, self.node_model_cross, 0, aggregates)
        new_hyper = self._fused_node_update(graph, edge_sets, self.hyper_node_model_cross, 1, aggregates)
        graph.node_features[0] = new_mesh
        graph.node_features[1] = new_hyper
