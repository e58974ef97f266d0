"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the message-passing processor hot path.

This file is a *restatement* (functional, weight-dict driven, torch CPU ops) of the algorithm in the
reference's ``src/migration`` + ``src/util.py``.  It is the checker: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.
The product path (``hgn_b200``) never does, and fails loudly if its CUDA extension is missing.

Pinning: the reference ships no golden vectors or known-answer tests for this path (SURVEY.md s4,
s8c: "parity unpinned by the reference's own tests").  The oracle is therefore pinned against outputs
of the reference itself, imported in the build container through ``oracle/reference_shim.py``:
``tests/golden/make_golden.py`` writes the committed fixtures in ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks this file against them (and live against the reference when
``/root/reference`` is mounted).  Third-party arithmetic absent from the reference tree:
``torch-scatter==2.0.9`` (``requirements.txt:7``); its published semantics are restated in
``segment_reduce`` below.

Every function names the reference lines it follows (paths relative to ``/root/reference``).
Weights are addressed by the reference's ``state_dict`` keys, e.g.
``processor.graphnet_blocks.0.edge_models.mesh_edges.0.layers.linear_0.weight``.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# Same field order as src/util.py:11-12.
EdgeSet = namedtuple("EdgeSet", ["name", "features", "senders", "receivers"])
MultiGraph = namedtuple("MultiGraph", ["node_features", "edge_sets"])

LN_EPS = 1e-5  # nn.LayerNorm default, src/migration/meshgraphnet.py:58-59


# ----------------------------------------------------------------------------------------------
# a13  LazyMLP / _make_mlp  (src/migration/meshgraphnet.py:53-60, 93-108)
# ----------------------------------------------------------------------------------------------
def mlp(weights: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, layer_norm: bool = True) -> torch.Tensor:
    """``Linear -> ReLU -> Linear -> ReLU -> Linear [-> LayerNorm]``.

    ``prefix`` addresses an ``nn.Sequential(LazyMLP, LayerNorm)`` (keys ``<prefix>.0.layers.linear_k.*``
    and ``<prefix>.1.*``) when ``layer_norm`` else a bare ``LazyMLP`` (``<prefix>.layers.linear_k.*``).
    ReLU follows every linear except the last (meshgraphnet.py:101-102).
    """
    base = f"{prefix}.0.layers" if layer_norm else f"{prefix}.layers"
    n_lin = 0
    while f"{base}.linear_{n_lin}.weight" in weights:
        n_lin += 1
    h = x
    for k in range(n_lin):
        h = F.linear(h, weights[f"{base}.linear_{k}.weight"], weights[f"{base}.linear_{k}.bias"])
        if k < n_lin - 1:
            h = F.relu(h)
    if layer_norm:
        h = F.layer_norm(h, (h.shape[-1],), weights[f"{prefix}.1.weight"], weights[f"{prefix}.1.bias"], LN_EPS)
    return h


# ----------------------------------------------------------------------------------------------
# a4  unsorted_segment_operation  (src/util.py:92-134)  + torch_scatter 2.0.9 semantics
# ----------------------------------------------------------------------------------------------
class _SegmentArgReduce(torch.autograd.Function):
    """max / min over segments; gradient goes to ONE winner (first edge in input order)."""

    @staticmethod
    def forward(ctx, data, seg, num_segments, is_max):
        E = data.shape[0]
        tail = tuple(data.shape[1:])
        idx = seg.view((E,) + (1,) * len(tail)).expand_as(data)
        out = torch.zeros((num_segments,) + tail, dtype=data.dtype, device=data.device)
        out.scatter_reduce_(0, idx, data, "amax" if is_max else "amin", include_self=False)
        pos = torch.arange(E, dtype=torch.int64, device=data.device).view((E,) + (1,) * len(tail)).expand_as(data)
        cand = torch.where(data == out.gather(0, idx), pos, torch.full_like(pos, E))
        arg = torch.full(out.shape, E, dtype=torch.int64, device=data.device)
        arg.scatter_reduce_(0, idx, cand, "amin", include_self=True)
        ctx.save_for_backward(arg)
        ctx.E = E
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        buf = torch.zeros((ctx.E + 1,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        buf.scatter_(0, arg, g)
        return buf[: ctx.E], None, None, None


def segment_reduce(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int, operation: str) -> torch.Tensor:
    """``out[segment_ids[i], ...] (+)= data[i, ...]`` -- util.py:92-134.

    Empty segments give 0 for every op; ``mean`` = ``sum / max(count, 1)`` (torch_scatter 2.0.9
    ``scatter_mean``); ``max``/``min`` send the gradient to a single arg element; ``std`` is the
    unbiased torch_scatter form.  Computation is in fp32 (``data.float()``, util.py:118) and the
    result is cast back to ``data.dtype`` (util.py:133).  Shape asserts follow util.py:101-102,112.
    """
    assert all(i in data.shape for i in segment_ids.shape), "segment_ids.shape should be a prefix of data.shape"
    if segment_ids.dim() != 1:
        assert data.shape == segment_ids.shape, "data.shape and segment_ids.shape should be equal"
        # the reference accepts a pre-expanded index; only the per-row form reaches the hot path
        seg = segment_ids.reshape(segment_ids.shape[0], -1)[:, 0]
    else:
        seg = segment_ids
    seg = seg.to(torch.int64)
    x = data.float()
    E = x.shape[0]
    tail = tuple(x.shape[1:])
    idx = seg.view((E,) + (1,) * len(tail)).expand_as(x)
    if operation == "sum":
        out = torch.zeros((num_segments,) + tail, dtype=x.dtype, device=x.device).scatter_add(0, idx, x)
    elif operation == "mean":
        total = torch.zeros((num_segments,) + tail, dtype=x.dtype, device=x.device).scatter_add(0, idx, x)
        count = torch.bincount(seg, minlength=num_segments).clamp(min=1).to(x.dtype)
        out = total / count.view((-1,) + (1,) * len(tail))
    elif operation == "max":
        out = _SegmentArgReduce.apply(x, seg, num_segments, True)
    elif operation == "min":
        out = _SegmentArgReduce.apply(x, seg, num_segments, False)
    elif operation == "std":
        count = torch.bincount(seg, minlength=num_segments).to(x.dtype)
        shape1 = (-1,) + (1,) * len(tail)
        mean = torch.zeros((num_segments,) + tail, dtype=x.dtype, device=x.device).scatter_add(0, idx, x) / count.clamp(min=1).view(shape1)
        var = torch.zeros((num_segments,) + tail, dtype=x.dtype, device=x.device).scatter_add(0, idx, (x - mean.gather(0, idx)) ** 2)
        out = (var / ((count - 1).clamp(min=1).view(shape1) + 1e-6)).sqrt()
    else:
        raise Exception("Invalid operation type!")
    return out.type(data.dtype)


PNA_OPS = ("sum", "mean", "max", "min")  # order fixed by graphnet.py:53-64


# ----------------------------------------------------------------------------------------------
# a2  edge update  (src/migration/graphnet.py:22-32)
# ----------------------------------------------------------------------------------------------
def edge_update(weights, prefix: str, node_features: Sequence[torch.Tensor], edge_set: EdgeSet) -> torch.Tensor:
    """``e' = e + LN(MLP([v[s] | v[r] | e]))`` with ``v`` the row-concatenation of the node list."""
    v = torch.cat(tuple(node_features), dim=0)
    x = torch.cat([v.index_select(0, edge_set.senders), v.index_select(0, edge_set.receivers), edge_set.features], dim=-1)
    return edge_set.features + mlp(weights, f"{prefix}.edge_models.{edge_set.name}", x)


# ----------------------------------------------------------------------------------------------
# a3  aggregation  (src/migration/graphnet.py:50-70)
# ----------------------------------------------------------------------------------------------
def aggregate(aggregator: str, v: torch.Tensor, edge_sets: Sequence[EdgeSet]) -> torch.Tensor:
    """``[v | agg(set_0) | agg(set_1) ...]``; 'pna' expands to sum, mean, max, min per set."""
    n = v.shape[0]
    cols = [v]
    for es in edge_sets:
        ops = PNA_OPS if aggregator == "pna" else (aggregator,)
        for op in ops:
            cols.append(segment_reduce(es.features, es.receivers, n, op))
    return torch.cat(cols, dim=-1)


def _with_model(edge_model_names, edge_sets):
    # graphnet.py:43 -- only sets that have an edge model take part in the aggregation
    return [es for es in edge_sets if es.name in edge_model_names]


def _edge_model_names(weights, prefix: str) -> List[str]:
    marker = f"{prefix}.edge_models."
    names: List[str] = []
    for k in weights:
        if k.startswith(marker):
            nm = k[len(marker):].split(".")[0]
            if nm not in names:
                names.append(nm)
    return names


# ----------------------------------------------------------------------------------------------
# a5 / a6 / a7  node updates  (graphnet.py:34-48, 94-108, 110-124; heterographnet.py:17-33)
# ----------------------------------------------------------------------------------------------
def node_update(weights, prefix, aggregator, node_features: List[torch.Tensor], edge_sets, model: str, target: int):
    """Aggregate over ALL rows, run ``model`` on the mesh rows (target 0) or hyper rows (target 1),
    add the residual and replace that list slot in place (like the reference mutates the list)."""
    n_mesh = node_features[0].shape[0]
    v = torch.cat(tuple(node_features), dim=0)
    feats = aggregate(aggregator, v, _with_model(_edge_model_names(weights, prefix), edge_sets))
    rows = feats[:n_mesh] if target == 0 else feats[n_mesh:]
    node_features[target] = mlp(weights, f"{prefix}.{model}", rows) + node_features[target]


# ----------------------------------------------------------------------------------------------
# a8 - a10  block schedules
# ----------------------------------------------------------------------------------------------
def block_graphnet(weights, prefix, aggregator, graph: MultiGraph) -> MultiGraph:
    """graphnet.py:72-84: all edge sets from the OLD node latents, then one mesh-node update.
    An edge set with no model raises KeyError (graphnet.py:32)."""
    names = _edge_model_names(weights, prefix)
    new_sets = []
    for es in graph.edge_sets:
        if es.name not in names:
            raise KeyError(es.name)
        new_sets.append(es._replace(features=edge_update(weights, prefix, graph.node_features, es)))
    nf = list(graph.node_features)
    node_update(weights, prefix, aggregator, nf, new_sets, "node_model_cross", 0)
    return MultiGraph(nf, new_sets)


def block_hetero(weights, prefix, aggregator, graph: MultiGraph) -> MultiGraph:
    """heterographnet.py:17-33 on top of graphnet.py:72-84: one aggregation, two MLPs."""
    names = _edge_model_names(weights, prefix)
    new_sets = []
    for es in graph.edge_sets:
        if es.name not in names:
            raise KeyError(es.name)
        new_sets.append(es._replace(features=edge_update(weights, prefix, graph.node_features, es)))
    nf = list(graph.node_features)
    n_mesh = nf[0].shape[0]
    feats = aggregate(aggregator, torch.cat(tuple(nf), 0), _with_model(names, new_sets))
    new_mesh = mlp(weights, f"{prefix}.node_model_cross", feats[:n_mesh]) + nf[0]
    new_hyper = mlp(weights, f"{prefix}.hyper_node_model_cross", feats[n_mesh:]) + nf[1]
    return MultiGraph([new_mesh, new_hyper], new_sets)


def _phase_edges(weights, prefix, names, graph_sets, nf, name, new_sets):
    # graphnet.py:86-92 perform_edge_updates: silently a no-op for names without a model
    if name not in names:
        return
    es = [e for e in graph_sets if e.name == name][0]
    new_sets[name] = es._replace(features=edge_update(weights, prefix, nf, es))


def block_hyper(weights, prefix, aggregator, graph: MultiGraph, set_order=None) -> MultiGraph:
    """hypergraphnet.py:21-54 four-phase schedule.  ``set_order`` pins the iteration order of the
    reference's ``set.intersection`` (hypergraphnet.py:31,44: hash-seed dependent);
    default is the list order ('mesh_edges', 'world_edges')."""
    names = _edge_model_names(weights, prefix)
    nf = list(graph.node_features)
    new_sets: Dict[str, EdgeSet] = {}
    order = set_order or {}
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "mesh_edges", new_sets)
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "world_edges", new_sets)
    mesh_phase = [n for n in order.get("mesh", ("mesh_edges", "world_edges")) if n in names]
    node_update(weights, prefix, aggregator, nf, [new_sets[n] for n in mesh_phase], "node_model_cross", 0)
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "intra_cluster_to_cluster", new_sets)
    node_update(weights, prefix, aggregator, nf, [new_sets["intra_cluster_to_cluster"]], "hyper_node_model_up", 1)
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "inter_cluster", new_sets)
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "inter_cluster_world", new_sets)
    inter_phase = [n for n in order.get("inter", ("inter_cluster", "inter_cluster_world")) if n in names]
    node_update(weights, prefix, aggregator, nf, [new_sets[n] for n in inter_phase], "hyper_node_model_cross", 1)
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "intra_cluster_to_mesh", new_sets)
    node_update(weights, prefix, aggregator, nf, [new_sets["intra_cluster_to_mesh"]], "node_model_down", 0)
    return MultiGraph(nf, list(new_sets.values()))


def block_multiscale(weights, prefix, aggregator, graph: MultiGraph, set_order=None) -> MultiGraph:
    """multiscalegraphnet.py:20-63: mesh, up, 3x inter (separate MLPs), down, mesh again."""
    names = _edge_model_names(weights, prefix)
    nf = list(graph.node_features)
    new_sets: Dict[str, EdgeSet] = {}
    order = set_order or {}

    def mesh_phase():
        # later phases read the ORIGINAL graph.edge_sets features again (multiscalegraphnet.py:56-57
        # passes `graph`, whose edge_sets are never replaced) but the UPDATED node list
        _phase_edges(weights, prefix, names, graph.edge_sets, nf, "mesh_edges", new_sets)
        _phase_edges(weights, prefix, names, graph.edge_sets, nf, "world_edges", new_sets)
        sel = [n for n in order.get("mesh", ("mesh_edges", "world_edges")) if n in names]
        node_update(weights, prefix, aggregator, nf, [new_sets[n] for n in sel], "node_model_cross", 0)

    mesh_phase()
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "intra_cluster_to_cluster", new_sets)
    node_update(weights, prefix, aggregator, nf, [new_sets["intra_cluster_to_cluster"]], "hyper_node_model_up", 1)
    for i in range(3):
        _phase_edges(weights, prefix, names, graph.edge_sets, nf, "inter_cluster", new_sets)
        _phase_edges(weights, prefix, names, graph.edge_sets, nf, "inter_cluster_world", new_sets)
        sel = [n for n in order.get("inter", ("inter_cluster", "inter_cluster_world")) if n in names]
        node_update(weights, prefix, aggregator, nf, [new_sets[n] for n in sel], f"hyper_node_models_cross.{i}", 1)
    _phase_edges(weights, prefix, names, graph.edge_sets, nf, "intra_cluster_to_mesh", new_sets)
    node_update(weights, prefix, aggregator, nf, [new_sets["intra_cluster_to_mesh"]], "node_model_down", 0)
    mesh_phase()
    return MultiGraph(nf, list(new_sets.values()))


def block_repeated(weights, prefix, aggregator, graph: MultiGraph, repetitions: int = 2) -> MultiGraph:
    """repeatedgraphnet.py:18-22: the base block applied ``repetitions`` times, shared weights."""
    for _ in range(repetitions):
        graph = block_graphnet(weights, prefix, aggregator, graph)
    return graph


BLOCKS = {
    "hyper": block_hyper,
    "multiscale": block_multiscale,
    "hetero": block_hetero,
    "multi": block_graphnet,       # multigraphnet.py:16-18 defers to the base block
    "repeated": block_repeated,
}


# ----------------------------------------------------------------------------------------------
# a11  Processor  (src/migration/processor.py:15-28)
# ----------------------------------------------------------------------------------------------
def processor(weights, aggregator: str, architecture: str, graph: MultiGraph, prefix: str = "processor",
              set_order=None) -> MultiGraph:
    """L sequential blocks with unshared weights ``<prefix>.graphnet_blocks.<i>``."""
    block = BLOCKS.get(architecture, block_graphnet)
    i = 0
    while any(k.startswith(f"{prefix}.graphnet_blocks.{i}.") for k in weights):
        p = f"{prefix}.graphnet_blocks.{i}"
        if block in (block_hyper, block_multiscale):
            graph = block(weights, p, aggregator, graph, set_order)
        else:
            graph = block(weights, p, aggregator, graph)
        i += 1
    return graph


# ----------------------------------------------------------------------------------------------
# a12  Encoder / Decoder / MeshGraphNet.forward  (encoder.py:24-47, decoder.py:15-16, meshgraphnet.py:46-51)
# ----------------------------------------------------------------------------------------------
HIERARCHICAL = ("hyper", "multiscale", "hetero")  # meshgraphnet.py:76-83


def encode(weights, architecture: str, graph: MultiGraph) -> MultiGraph:
    nodes = [mlp(weights, "encoder.node_model", graph.node_features[0])]
    if len(graph.node_features) > 1:
        model = "encoder.hyper_node_model" if architecture in HIERARCHICAL else "encoder.node_model"
        nodes.append(mlp(weights, model, graph.node_features[1]))
    names = _edge_model_names(weights, "encoder")
    sets = [es._replace(features=mlp(weights, f"encoder.edge_models.{es.name}", es.features))
            for es in graph.edge_sets if es.name in names]  # unknown sets are dropped, encoder.py:41-45
    return MultiGraph(nodes, sets)


def mesh_graph_net(weights, aggregator: str, architecture: str, graph: MultiGraph, set_order=None) -> torch.Tensor:
    latent = processor(weights, aggregator, architecture, encode(weights, architecture, graph), set_order=set_order)
    return mlp(weights, "decoder.model", latent.node_features[0], layer_norm=False)


# ----------------------------------------------------------------------------------------------
# graph structure (bit-exact contract)  (src/util.py:50-89)
# ----------------------------------------------------------------------------------------------
def triangles_to_edges(faces: torch.Tensor, deform: bool = False):
    """Unique undirected (max,min) pairs in lexicographic order, then the reversed copies.
    Tetrahedra use edges (0,1),(1,2),(2,3),(3,0) only (util.py:72-75)."""
    k = faces.shape[1]
    assert k == (4 if deform else 3)
    pairs = torch.cat([torch.stack((faces[:, i], faces[:, (i + 1) % k]), dim=1) for i in range(k)], dim=0)
    hi = pairs.max(dim=1).values
    lo = pairs.min(dim=1).values
    uniq = torch.unique(torch.stack((hi, lo), dim=1), dim=0)
    s = uniq[:, 0].to(torch.int64)
    r = uniq[:, 1].to(torch.int64)
    return {"two_way_connectivity": (torch.cat((s, r)), torch.cat((r, s))), "senders": s, "receivers": r}


# ----------------------------------------------------------------------------------------------
# world edges of the deforming-plate graph  (src/model/plate.py:86-110)
# ----------------------------------------------------------------------------------------------
def world_edges(world_pos: torch.Tensor, node_type: torch.Tensor, mesh_senders: torch.Tensor, mesh_receivers: torch.Tensor,
                radius: float = 0.03, obstacle: int = 1, normal: int = 0):
    """Dense restatement, statement by statement: ``torch.cdist`` matrix, ``< radius``, diagonal cleared, existing mesh edges
    cleared, rows that are not OBSTACLE nodes and columns that are not NORMAL nodes cleared, ``torch.nonzero``.
    ``node_type`` is the reference's ``[N, 1]`` tensor (or ``[N]``).  O(N^2) memory: small cases only."""
    types = node_type.reshape(-1)
    dist = torch.cdist(world_pos, world_pos, p=2)
    conn = torch.where(dist < radius, True, False)
    conn = conn.fill_diagonal_(False)
    conn[mesh_senders, mesh_receivers] = False
    conn[torch.ne(types, obstacle), :] = False
    conn[:, torch.ne(types, normal)] = False
    senders, receivers = torch.nonzero(conn, as_tuple=True)
    return senders, receivers
