"""Mirror of src/migration/encoder.py.  Runs once per step on stock torch ``Linear``/``LayerNorm``
(out of the per-layer hot path); same attribute names and ``state_dict`` keys."""
from typing import Callable, List

from torch import nn

from ..util import MultiGraph


class Encoder(nn.Module):
    """Encodes node and edge features into latent features."""

    def __init__(self, make_mlp: Callable, latent_size: int, edge_sets: List[str], hierarchical=True):
        super().__init__()
        self._make_mlp = make_mlp
        self._latent_size = latent_size
        self.node_model = self._make_mlp(latent_size)
        self.edge_models = nn.ModuleDict({name: self._make_mlp(latent_size) for name in edge_sets})
        self.hierarchical = hierarchical
        if hierarchical:
            self.hyper_node_model = self._make_mlp(latent_size)

    def forward(self, graph: MultiGraph) -> MultiGraph:
        nodes = [self.node_model(graph.node_features[0])]
        # a second node list (hyper nodes) goes through its own MLP only for hierarchical
        # architectures (encoder.py:26-35); a missing second entry is tolerated silently
        if isinstance(graph.node_features, (list, tuple)) and len(graph.node_features) > 1:
            hyper_model = self.hyper_node_model if self.hierarchical else self.node_model
            nodes.append(hyper_model(graph.node_features[1]))
        # edge sets without an encoder model are dropped (encoder.py:41-45)
        edge_sets = [es._replace(features=self.edge_models[es.name](es.features))
                     for es in graph.edge_sets if es.name in self.edge_models]
        return MultiGraph(nodes, edge_sets)
