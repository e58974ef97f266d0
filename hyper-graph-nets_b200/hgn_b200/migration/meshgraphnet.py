"""Mirror of src/migration/meshgraphnet.py: the Encode-Process-Decode shell and the MLP definition.

Constructor signature, attribute names and ``state_dict`` keys are the reference's
(``encoder.node_model.0.layers.linear_k.*``, ``processor.graphnet_blocks.<i>.edge_models.<set>...``,
``decoder.model.layers.linear_k.*``), so a reference ``state_dict`` loads here and vice versa, and
``nn.LazyLinear`` is kept so that ``optim.Adam(model.parameters())`` built before the first forward
(MeshSimulator.py:109-110) holds the same Parameter objects.
"""
import functools
from collections import OrderedDict
from typing import List, Tuple, Type

from torch import nn, Tensor

from .decoder import Decoder
from .encoder import Encoder
from .graphnet import GraphNet
from .heterographnet import HeteroGraphNet
from .hypergraphnet import HyperGraphNet
from .multigraphnet import MultiGraphNet
from .multiscalegraphnet import MultiScaleGraphNet
from .processor import Processor
from .repeatedgraphnet import RepeatedGraphNet
from .. import util
from ..util import MultiGraph

_ARCHITECTURES = {
    # name -> (block, hierarchical encoder?)   meshgraphnet.py:62-89
    'hyper': (HyperGraphNet, True),
    'multiscale': (MultiScaleGraphNet, True),
    'hetero': (HeteroGraphNet, True),
    'multi': (MultiGraphNet, False),
    'repeated': (RepeatedGraphNet, False),
}


class MeshGraphNet(nn.Module):
    """Encode-Process-Decode GraphNet model."""

    def __init__(self, output_size: int, latent_size: int, num_layers: int, message_passing_aggregator: str,
                 message_passing_steps: int, architecture: str, edge_sets: List[str]):
        super().__init__()
        self._latent_size = latent_size
        self._output_size = output_size
        self._num_layers = num_layers
        self._message_passing_steps = message_passing_steps
        self._message_passing_aggregator = message_passing_aggregator
        graphnet_block, hierarchical = self.get_architecture(architecture)
        self.encoder = Encoder(make_mlp=self._make_mlp, latent_size=self._latent_size,
                               hierarchical=hierarchical, edge_sets=edge_sets)
        self.processor = Processor(make_mlp=self._make_mlp, output_size=self._latent_size,
                                   message_passing_steps=self._message_passing_steps,
                                   message_passing_aggregator=self._message_passing_aggregator,
                                   edge_sets=edge_sets, graphnet_block=graphnet_block)
        self.decoder = Decoder(make_mlp=functools.partial(self._make_mlp, layer_norm=False),
                               output_size=self._output_size)

    def forward(self, graph: MultiGraph) -> Tensor:
        """Encodes and processes a multigraph, and returns node features."""
        latent = self.processor(self.encoder(graph))
        return self.decoder(latent._replace(node_features=latent.node_features[0]))

    def _make_mlp(self, output_size: int, layer_norm=True) -> nn.Module:
        """``[latent] * num_layers + [output_size]`` lazy linears, ReLU between, optional LayerNorm."""
        network = LazyMLP([self._latent_size] * self._num_layers + [output_size])
        if layer_norm:
            network = nn.Sequential(network, nn.LayerNorm(normalized_shape=output_size))
        return network

    @staticmethod
    def get_architecture(architecture: str) -> Tuple[Type[GraphNet], bool]:
        """The GraphNet block class for ``architecture`` and whether its encoder is hierarchical;
        anything unknown selects the base ``GraphNet``."""
        return _ARCHITECTURES.get(architecture, (GraphNet, False))


class LazyMLP(nn.Module):
    def __init__(self, output_sizes: List[int]):
        super().__init__()
        layers = OrderedDict()
        for index, width in enumerate(output_sizes):
            layers[f'linear_{index}'] = nn.LazyLinear(width)
            if index + 1 < len(output_sizes):
                layers[f'relu_{index}'] = nn.ReLU()
        self._layers_ordered_dict = layers
        self.layers = nn.Sequential(layers)

    def forward(self, input: Tensor) -> Tensor:
        return self.layers(input.to(util.device))
