"""``repetitions`` passes of the base block's schedule over ONE set of weights (API mirror of
src/migration/repeatedgraphnet.py:11-24).  Every pass runs the fused kernels of ``GraphNet``; the packed bf16 weight
blobs are built once and reused by all passes (the Parameters do not change between them)."""
from typing import Callable, List

from ..util import MultiGraph
from .graphnet import GraphNet


class RepeatedGraphNet(GraphNet):
    """Weight-shared repetition of the multi-edge interaction block."""

    def __init__(self, model_fn: Callable, output_size: int, message_passing_aggregator: str, edge_sets: List[str], repetitions=2):
        GraphNet.__init__(self, model_fn=model_fn, output_size=output_size,
                          message_passing_aggregator=message_passing_aggregator, edge_sets=edge_sets)
        self.repetitions = repetitions

    def forward(self, graph: MultiGraph, mask=None) -> MultiGraph:
        remaining = int(self.repetitions)
        while remaining > 0:
            graph = GraphNet.forward(self, graph, mask)
            remaining -= 1
        return graph
