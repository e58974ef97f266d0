"""Mirror of src/migration/repeatedgraphnet.py: the base block applied ``repetitions`` times with
shared weights."""
from typing import Callable, List

from .graphnet import GraphNet
from ..util import MultiGraph


class RepeatedGraphNet(GraphNet):
    """Multi-Edge and Multi-Node Interaction Network with residual connections."""

    def __init__(self, model_fn: Callable, output_size: int, message_passing_aggregator: str, edge_sets: List[str], repetitions=2):
        super().__init__(model_fn, output_size, message_passing_aggregator, edge_sets)
        self.repetitions = repetitions

    def forward(self, graph: MultiGraph, mask=None) -> MultiGraph:
        for _ in range(self.repetitions):
            graph = GraphNet.forward(self, graph)
        return graph
