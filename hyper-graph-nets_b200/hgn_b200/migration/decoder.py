"""Last stage of encode-process-decode (API mirror of src/migration/decoder.py:9-16).

The decoder MLP has no LayerNorm (meshgraphnet.py:41-43 builds it with ``layer_norm=False``), so it is outside the fused
LayerNorm kernels' contract and runs as stock torch: it touches ``[N, 128] -> [N, output_size]`` once per step, not per layer.
"""
from typing import Callable

from torch import Tensor, nn

from ..util import MultiGraph


class Decoder(nn.Module):
    """Maps the mesh-node latents to the model output.

    ``MeshGraphNet.forward`` hands over a graph whose ``node_features`` field is already the mesh-node latent TENSOR
    (``node_features[0]`` of the processed graph); a one-element list is accepted as well."""

    def __init__(self, make_mlp: Callable, output_size: int):
        super().__init__()
        # state_dict contract: decoder.model.layers.linear_{0,1,2}.{weight,bias} -- no '.0.' level, there is no LayerNorm wrapper
        self.model = make_mlp(output_size)

    def forward(self, graph: MultiGraph) -> Tensor:
        latents = graph.node_features
        if isinstance(latents, (list, tuple)):
            latents = latents[0]
        return self.model(latents)
