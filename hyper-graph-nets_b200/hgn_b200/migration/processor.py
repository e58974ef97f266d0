"""Mirror of src/migration/processor.py: ``message_passing_steps`` blocks with unshared weights.

Precision: the reference computes in fp32.  ``precision='fp32'`` keeps fp32 storage and fp32 FFMA
arithmetic (parity mode, 1e-5 relative).  ``precision='bf16'`` converts the latents to bf16 once on
entry, runs every block on the bf16 tcgen05 kernels (fp32 accumulation, fp32 LayerNorm/residual
arithmetic) and converts the NODE latents back on exit (throughput mode, 2e-2 relative); the edge latents
are returned in bf16 -- nothing downstream in the reference reads them (decoder.py:15-16 decodes
``node_features[0]`` only) and an [E,128] fp32 round trip per step would cost two full HBM passes.  Set
``processor.restore_edge_dtype = True`` to get them back in the caller's dtype.  Select the mode per module
(``processor.precision = 'bf16'``), globally (``hgn_b200.set_precision``) or with the environment
variable ``HGN_B200_PRECISION``; no config-schema change is needed.
"""
from typing import Callable, List, Type

import torch
from torch import nn

from .. import config
from .graphnet import GraphNet
from ..util import MultiGraph


class Processor(nn.Module):
    """The Graph Neural Network that transforms the input graph."""

    def __init__(self, make_mlp: Callable, output_size: int, message_passing_steps: int,
                 message_passing_aggregator: str, edge_sets: List[str], graphnet_block: Type[GraphNet]):
        super().__init__()
        self.graphnet_blocks = nn.Sequential(*[
            graphnet_block(model_fn=make_mlp, output_size=output_size,
                           message_passing_aggregator=message_passing_aggregator, edge_sets=edge_sets)
            for _ in range(message_passing_steps)])
        self.precision = None       # None -> hgn_b200.config.precision()
        self.restore_edge_dtype = False

    def forward(self, latent_graph: MultiGraph) -> MultiGraph:
        precision = self.precision or config.precision()
        if precision == 'fp32':
            return self.graphnet_blocks(latent_graph)
        if precision != 'bf16':
            raise ValueError(f"unknown precision {precision!r} (expected 'fp32' or 'bf16')")
        in_dtype = latent_graph.node_features[0].dtype
        nodes = latent_graph.node_features
        for i in range(len(nodes)):      # keep the caller's list object, like the blocks do
            nodes[i] = nodes[i].to(torch.bfloat16)
        graph = latent_graph._replace(edge_sets=[es._replace(features=es.features.to(torch.bfloat16))
                                                 for es in latent_graph.edge_sets])
        orders = None
        if config.edge_storage() == 'receiver_sorted':       # experimental: see plan.EdgeStorageOrder
            from ..plan import edge_storage_order, to_device_index
            dev = graph.node_features[0].device
            orders = {}
            sets = []
            for es in graph.edge_sets:
                order = edge_storage_order(to_device_index(es.senders, dev), to_device_index(es.receivers, dev))
                orders[es.name] = (order, es.senders, es.receivers)
                sets.append(es._replace(features=order.store(es.features), senders=order.senders, receivers=order.receivers))
            graph = graph._replace(edge_sets=sets)
        graph = self.graphnet_blocks(graph)
        if orders is not None:                                # rows, and the caller's own index tensors, back in the reference order
            graph = graph._replace(edge_sets=[es._replace(features=orders[es.name][0].restore(es.features), senders=orders[es.name][1],
                                                          receivers=orders[es.name][2]) for es in graph.edge_sets])
        for i in range(len(graph.node_features)):
            graph.node_features[i] = graph.node_features[i].to(in_dtype)
        if not getattr(self, 'restore_edge_dtype', False):
            return graph
        return graph._replace(edge_sets=[es._replace(features=es.features.to(in_dtype)) for es in graph.edge_sets])
