"""Edge-cut graph partitioning with a one-ring halo of sender latents.

The reference is single-device (no ``torch.distributed`` anywhere), so this is new functionality whose
parity contract is "P-rank result == 1-rank result".  Scheme:

* nodes are split into P parts; every directed edge belongs to the OWNER OF ITS RECEIVER, so the
  edge->node aggregation and the node update are rank-local and edge latents never move;
* the only remote data are sender latents of cut edges: each rank keeps them as GHOST rows.  A rank's
  node list is ``[owned | ghosts]`` -- exactly the reference's ``node_features = [mesh, hyper]`` row
  concatenation (util.py:11-12, hierarchical_connector.py:37), so the unmodified ``GraphNet`` block
  updates the owned rows (list slot 0) and leaves the ghosts (slot 1) alone;
* before every block the owners send their boundary rows to the ranks that hold them as ghosts
  (``HaloExchange``: pack kernel -> grouped NCCL send/recv over NVLink -> ghosts land contiguously, ordered by
  owner, so no unpack is needed); in the backward pass ghost gradients travel back and are added at the
  owner peer by peer in rank order (deterministic);
* weight gradients are summed with one all-reduce after the backward pass.

Nothing here does arithmetic on latents besides the pack / scatter-add row kernels of ``libhgn_b200.so``
(CPU tensors use torch indexing instead, which is how the world_size-2 gloo tests exercise this logic).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _cabi


# ------------------------------------------------------------------------------------------------------
# partition plan (pure index bookkeeping, bit-exact, CPU)
# ------------------------------------------------------------------------------------------------------
@dataclass
class LocalGraph:
    rank: int
    world: int
    owned: torch.Tensor            # [n_own]   global ids of the owned nodes (ascending)
    ghosts: torch.Tensor           # [n_ghost] global ids of the ghost nodes, sorted by (owner, id)
    ghost_splits: List[int]        # ghosts received from each rank (sums to n_ghost)
    send_index: torch.Tensor       # [n_send]  LOCAL owned indices to send, grouped by destination rank
    send_splits: List[int]         # rows sent to each rank
    edge_ids: torch.Tensor         # [e_loc]   global ids of the local edges (ascending -> original order kept)
    senders: torch.Tensor          # [e_loc]   local sender index into [owned | ghosts]
    receivers: torch.Tensor        # [e_loc]   local receiver index (always < n_own)
    n_interior: Optional[int] = None   # with interior_first: edges [0, n_interior) have an owned sender, the rest a ghost sender

    @property
    def n_own(self) -> int:
        return int(self.owned.numel())

    @property
    def n_ghost(self) -> int:
        return int(self.ghosts.numel())


def block_partition(num_nodes: int, world: int) -> torch.Tensor:
    """Contiguous id ranges (row slabs of a row-major grid mesh): part id per node."""
    bounds = [(num_nodes * p) // world for p in range(world + 1)]
    part = torch.empty(num_nodes, dtype=torch.int64)
    for p in range(world):
        part[bounds[p]:bounds[p + 1]] = p
    return part


def coordinate_bisection(pos: torch.Tensor, world: int) -> torch.Tensor:
    """Recursive coordinate bisection of node positions into ``world`` (power of two) parts."""
    n = pos.shape[0]
    part = torch.zeros(n, dtype=torch.int64)
    groups = [torch.arange(n)]
    while len(groups) < world:
        nxt = []
        for ids in groups:
            p = pos[ids]
            axis = int(torch.argmax(p.max(0).values - p.min(0).values))
            order = torch.argsort(p[:, axis], stable=True)
            half = ids.numel() // 2
            nxt += [ids[order[:half]], ids[order[half:]]]
        groups = nxt
    for k, ids in enumerate(groups):
        part[ids] = k
    return part


def build_local_graph(senders: torch.Tensor, receivers: torch.Tensor, part: torch.Tensor, rank: int, world: int,
                      interior_first: bool = False) -> LocalGraph:
    """Rank ``rank``'s share of a directed edge list under the receiver-owner rule.  ``interior_first`` lists the edges whose
    sender is owned before the cut edges (each group in original order), which is what lets the halo exchange overlap the
    interior edge tiles (``PartitionedProcessor``)."""
    senders, receivers, part = senders.cpu(), receivers.cpu(), part.cpu()
    owned = torch.nonzero(part == rank).flatten()
    edge_ids = torch.nonzero(part[receivers] == rank).flatten()
    n_interior = None
    if interior_first:
        cut = part[senders[edge_ids]] != rank
        edge_ids = torch.cat([edge_ids[~cut], edge_ids[cut]])
        n_interior = int((~cut).sum())
    s_glob, r_glob = senders[edge_ids], receivers[edge_ids]
    # ghosts: senders of local edges owned elsewhere, ordered by (owner, global id)
    remote = torch.unique(s_glob[part[s_glob] != rank])
    order = torch.argsort(part[remote] * (part.numel() + 1) + remote)
    ghosts = remote[order]
    ghost_splits = torch.bincount(part[ghosts], minlength=world).tolist()
    # global -> local id
    local_of = torch.full((part.numel(),), -1, dtype=torch.int64)
    local_of[owned] = torch.arange(owned.numel())
    local_of[ghosts] = owned.numel() + torch.arange(ghosts.numel())
    # what do the other ranks need from me?  (their ghost lists restricted to my nodes, same ordering rule)
    send_lists = []
    for q in range(world):
        if q == rank:
            send_lists.append(torch.empty(0, dtype=torch.int64))
            continue
        q_edges = part[receivers] == q
        need = torch.unique(senders[q_edges & (part[senders] == rank)])   # ascending global id == q's ghost order
        send_lists.append(local_of[need])
    return LocalGraph(rank=rank, world=world, owned=owned, ghosts=ghosts, ghost_splits=ghost_splits,
                      send_index=torch.cat(send_lists), send_splits=[int(t.numel()) for t in send_lists],
                      edge_ids=edge_ids, senders=local_of[s_glob], receivers=local_of[r_glob], n_interior=n_interior)


# ------------------------------------------------------------------------------------------------------
# halo exchange
# ------------------------------------------------------------------------------------------------------
def _gather_rows(src: torch.Tensor, index: torch.Tensor, index32: Optional[torch.Tensor]) -> torch.Tensor:
    if not src.is_cuda:
        return src.index_select(0, index)
    lib = _cabi.load()
    out = torch.empty((index.numel(), src.shape[1]), dtype=src.dtype, device=src.device)
    if index.numel():
        _cabi.check(lib.hgn_rows_gather(_cabi.dtype_code(src.dtype), src.data_ptr(), index32.data_ptr(), index.numel(), src.shape[1],
                                        out.data_ptr(), _cabi.stream_ptr()), "hgn_rows_gather")
    return out


def _scatter_add_rows(dst: torch.Tensor, rows: torch.Tensor, index: torch.Tensor, index32: Optional[torch.Tensor]) -> None:
    """dst[index[i]] += rows[i]; ``index`` has no duplicates (one peer at a time)."""
    if not rows.numel():
        return
    if not dst.is_cuda:
        dst.index_add_(0, index, rows)
        return
    lib = _cabi.load()
    _cabi.check(lib.hgn_rows_scatter(_cabi.dtype_code(dst.dtype), rows.data_ptr(), index32.data_ptr(), index.numel(), dst.shape[1],
                                     dst.data_ptr(), 1, _cabi.stream_ptr()), "hgn_rows_scatter")


def _exchange_async(send: torch.Tensor, send_splits: Sequence[int], recv_splits: Sequence[int], group):
    """Starts the all-to-all-v of ``_exchange`` and returns ``(recv, works)`` without waiting: kernels launched on the current
    stream afterwards run while the rows are in flight; ``work.wait()`` makes the current stream wait for them."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    recv = torch.empty((sum(recv_splits), send.shape[1]), dtype=send.dtype, device=send.device)
    ops_, so, ro = [], 0, 0
    for q in range(world):
        if q != rank and recv_splits[q]:
            ops_.append(dist.P2POp(dist.irecv, recv[ro:ro + recv_splits[q]], q, group))
        ro += recv_splits[q]
    for q in range(world):
        if q != rank and send_splits[q]:
            ops_.append(dist.P2POp(dist.isend, send[so:so + send_splits[q]].contiguous(), q, group))
        so += send_splits[q]
    return recv, (dist.batch_isend_irecv(ops_) if ops_ else [])


def _exchange(send: torch.Tensor, send_splits: Sequence[int], recv_splits: Sequence[int], group) -> torch.Tensor:
    """All-to-all-v of row blocks with grouped point-to-point operations (NCCL: one fused launch over NVLink)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    recv = torch.empty((sum(recv_splits), send.shape[1]), dtype=send.dtype, device=send.device)
    ops, so, ro = [], 0, 0
    for q in range(world):
        if q != rank and recv_splits[q]:
            ops.append(dist.P2POp(dist.irecv, recv[ro:ro + recv_splits[q]], q, group))
        ro += recv_splits[q]
    for q in range(world):
        if q != rank and send_splits[q]:
            ops.append(dist.P2POp(dist.isend, send[so:so + send_splits[q]].contiguous(), q, group))
        so += send_splits[q]
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    return recv


class HaloPlan:
    """Device-side copy of a ``LocalGraph``'s exchange lists."""

    def __init__(self, lg: LocalGraph, device, group=None):
        self.lg = lg
        self.group = group
        self.send_index = lg.send_index.to(device)
        self.send_index32 = self.send_index.to(torch.int32) if torch.device(device).type == "cuda" else None
        self.send_splits = list(lg.send_splits)
        self.ghost_splits = list(lg.ghost_splits)
        self.peer_offsets = [0]
        for n in self.send_splits:
            self.peer_offsets.append(self.peer_offsets[-1] + n)


class _HaloExchange(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owned: torch.Tensor, plan: HaloPlan):
        ctx.plan = plan
        ctx.n_own = owned.shape[0]
        packed = _gather_rows(owned.contiguous(), plan.send_index, plan.send_index32)
        return _exchange(packed, plan.send_splits, plan.ghost_splits, plan.group)

    @staticmethod
    def backward(ctx, grad_ghosts: torch.Tensor):
        plan: HaloPlan = ctx.plan
        back = _exchange(grad_ghosts.contiguous(), plan.ghost_splits, plan.send_splits, plan.group)
        grad_owned = torch.zeros((ctx.n_own, grad_ghosts.shape[1]), dtype=grad_ghosts.dtype, device=grad_ghosts.device)
        for q in range(len(plan.send_splits)):          # peers in rank order: fixed summation order
            lo, hi = plan.peer_offsets[q], plan.peer_offsets[q + 1]
            if hi > lo:
                idx32 = plan.send_index32[lo:hi] if plan.send_index32 is not None else None
                _scatter_add_rows(grad_owned, back[lo:hi], plan.send_index[lo:hi], idx32)
        return grad_owned, None


def halo_exchange(owned: torch.Tensor, plan: HaloPlan) -> torch.Tensor:
    """Ghost rows ``[n_ghost, D]`` (ordered by owner rank) holding the owners' current latents."""
    return _HaloExchange.apply(owned, plan)


class OverlapPlan:
    """Device-side plans of one rank's share for the overlapped edge update: int32 gather indices, CSR plans of the senders
    (all local edges, and the cut edges alone) and receivers, and the exchange lists."""

    def __init__(self, lg: LocalGraph, halo: "HaloPlan", device):
        from .plan import segment_plan
        if lg.n_interior is None:
            raise ValueError("OverlapPlan needs a LocalGraph built with interior_first=True")
        self.halo = halo
        self.n_own, self.n_ghost, self.n_interior = lg.n_own, lg.n_ghost, int(lg.n_interior)
        self.senders = lg.senders.to(device)
        self.receivers = lg.receivers.to(device)
        self.cut_senders = self.senders[self.n_interior:]
        self.s_plan = segment_plan(self.senders, self.n_own + self.n_ghost)
        self.r_plan = segment_plan(self.receivers, self.n_own)
        self.cut_plan = segment_plan(self.cut_senders, self.n_own + self.n_ghost) if self.cut_senders.numel() else None


class _PartitionedEdgeUpdate(torch.autograd.Function):
    """Edge update + 'sum' aggregation of one partitioned ``GraphNet`` block (bf16) with the halo exchange in flight behind the
    INTERIOR edge tiles: pack boundary rows -> start the NCCL exchange -> project the owned rows and run the fused edge kernel
    over the edges with an owned sender -> wait -> project the ghost rows and run the kernel over the cut edges.  Backward runs
    the cut edges first, sends the ghost gradients home, and does the interior edges while those travel."""

    @staticmethod
    def forward(ctx, owned, e, W0, b0, W1, b1, W2, b2, gamma, beta, packed, op: OverlapPlan):
        lib, BF, halo = _cabi.load(), _cabi.HGN_BF16, op.halo
        No, G, E, Ei = op.n_own, op.n_ghost, e.shape[0], op.n_interior
        owned, e = owned.contiguous(), e.contiguous()
        s32, r32 = op.s_plan.ids32, op.r_plan.ids32
        with torch.cuda.device(owned.device):
            st = _cabi.stream_ptr()
            send = _gather_rows(owned, halo.send_index, halo.send_index32)
            ghosts, works = _exchange_async(send, halo.send_splits, halo.ghost_splits, halo.group)
            ps = torch.empty((No + G, owned.shape[1]), dtype=owned.dtype, device=owned.device)
            pr = torch.empty_like(ps)
            out = torch.empty_like(e)
            _cabi.check(lib.hgn_edge_project_forward(BF, No, owned.data_ptr(), packed.data_ptr(), ps.data_ptr(), pr.data_ptr(), st),
                        "hgn_edge_project_forward")
            if Ei:
                _cabi.check(lib.hgn_edge_update_forward(BF, Ei, e.data_ptr(), ps.data_ptr(), pr.data_ptr(), s32.data_ptr(), r32.data_ptr(),
                                                        packed.data_ptr(), out.data_ptr(), None, None, st), "hgn_edge_update_forward")
            for w in works:
                w.wait()
            if G:
                _cabi.check(lib.hgn_edge_project_forward(BF, G, ghosts.data_ptr(), packed.data_ptr(), ps[No:].data_ptr(), pr[No:].data_ptr(), st),
                            "hgn_edge_project_forward")
            if E - Ei:
                _cabi.check(lib.hgn_edge_update_forward(BF, E - Ei, e[Ei:].data_ptr(), ps.data_ptr(), pr.data_ptr(), s32[Ei:].data_ptr(),
                                                        r32[Ei:].data_ptr(), packed.data_ptr(), out[Ei:].data_ptr(), None, None, st),
                            "hgn_edge_update_forward")
            agg = torch.empty((No, owned.shape[1]), dtype=owned.dtype, device=owned.device)
            _cabi.check(lib.hgn_segment_reduce(BF, out.data_ptr(), E, owned.shape[1], op.r_plan.perm.data_ptr(), op.r_plan.rowptr.data_ptr(), No,
                                               agg.data_ptr(), None, None, None, None, None, 0, st), "hgn_segment_reduce")
        from . import ops as _ops
        _ops._count(6)
        ctx.save_for_backward(owned, ghosts, e, ps, pr)
        ctx.op, ctx.packed = op, packed
        ctx.param_shapes = [tuple(p.shape) for p in (W0, b0, W1, b1, W2, b2, gamma, beta)]
        return out, agg

    @staticmethod
    def backward(ctx, grad_out, grad_agg):
        lib, BF, op = _cabi.load(), _cabi.HGN_BF16, ctx.op
        halo = op.halo
        owned, ghosts, e, ps, pr = ctx.saved_tensors
        No, G, E, Ei = op.n_own, op.n_ghost, e.shape[0], op.n_interior
        D, dev = owned.shape[1], owned.device
        s32, r32 = op.s_plan.ids32, op.r_plan.ids32
        grad_out = grad_out.contiguous().to(e.dtype) if grad_out is not None else None
        grad_agg = grad_agg.contiguous().to(e.dtype) if grad_agg is not None else None
        grad_e, g0 = torch.empty_like(e), torch.empty_like(e)

        def new_gparams():
            return [torch.zeros(shape, dtype=torch.float32, device=dev) for shape in ctx.param_shapes]

        with torch.cuda.device(dev):
            st = _cabi.stream_ptr()

            def edge_bwd(lo, n, gp):
                ws_bytes = lib.hgn_edge_update_backward_workspace_bytes(BF, n)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _cabi.check(lib.hgn_edge_update_backward(
                    BF, n, e[lo:].data_ptr(), ps.data_ptr(), pr.data_ptr(), s32[lo:].data_ptr(), r32[lo:].data_ptr(), None, None,
                    ctx.packed.data_ptr(), grad_out[lo:].data_ptr() if grad_out is not None else None, _cabi.ptr(grad_agg),
                    grad_e[lo:].data_ptr(), g0[lo:].data_ptr(), *[g.data_ptr() for g in gp], ws.data_ptr(), ws_bytes, st),
                    "hgn_edge_update_backward")

            def project_bwd(n, v, gs, gr, grad_v, gw0):
                ws_bytes = lib.hgn_edge_project_backward_workspace_bytes(BF, n)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _cabi.check(lib.hgn_edge_project_backward(BF, n, v.data_ptr(), ctx.packed.data_ptr(), gs.data_ptr(), gr.data_ptr(),
                                                          grad_v.data_ptr(), gw0.data_ptr(), ws.data_ptr(), ws_bytes, st),
                            "hgn_edge_project_backward")

            def seg_sum(data, n_rows, plan, n_seg, dst):
                _cabi.check(lib.hgn_segment_reduce(BF, data.data_ptr(), n_rows, D, plan.perm.data_ptr(), plan.rowptr.data_ptr(), n_seg,
                                                   dst.data_ptr(), None, None, None, None, None, 0, st), "hgn_segment_reduce")

            gp = new_gparams()
            back, works = None, []
            if E - Ei:                                      # 1. cut edges first: their G0 rows hold everything the ghosts' gradient needs
                edge_bwd(Ei, E - Ei, gp)
                gs_cut = torch.empty((No + G, D), dtype=e.dtype, device=dev)
                seg_sum(g0[Ei:], E - Ei, op.cut_plan, No + G, gs_cut)
                grad_ghosts = torch.empty((G, D), dtype=e.dtype, device=dev)
                gw0_ghost = torch.zeros(ctx.param_shapes[0], dtype=torch.float32, device=dev)
                project_bwd(G, ghosts, gs_cut[No:], torch.zeros((G, D), dtype=e.dtype, device=dev), grad_ghosts, gw0_ghost)
                back, works = _exchange_async(grad_ghosts, halo.ghost_splits, halo.send_splits, halo.group)   # 2. ghost gradients go home
                gp[0] += gw0_ghost
            if Ei:                                          # 3. interior edges while they travel
                gp_int = new_gparams()
                edge_bwd(0, Ei, gp_int)
                for a, b in zip(gp, gp_int):
                    a += b
            gs = torch.empty((No + G, D), dtype=e.dtype, device=dev)           # 4. owned rows: sender- and receiver-keyed sums of ALL local G0 rows
            gr = torch.empty((No, D), dtype=e.dtype, device=dev)
            seg_sum(g0, E, op.s_plan, No + G, gs)
            seg_sum(g0, E, op.r_plan, No, gr)
            grad_owned = torch.empty((No, D), dtype=e.dtype, device=dev)
            gw0_owned = torch.zeros(ctx.param_shapes[0], dtype=torch.float32, device=dev)
            project_bwd(No, owned, gs, gr, grad_owned, gw0_owned)
            gp[0] += gw0_owned
            for w in works:                                 # 5. received ghost gradients are added at their owners, peers in rank order
                w.wait()
            if back is not None:
                for q in range(len(halo.send_splits)):
                    lo, hi = halo.peer_offsets[q], halo.peer_offsets[q + 1]
                    if hi > lo:
                        _scatter_add_rows(grad_owned, back[lo:hi], halo.send_index[lo:hi], halo.send_index32[lo:hi])
        from . import ops as _ops
        _ops._count(16)
        return (grad_owned, grad_e, *gp, None, None)


class PartitionedProcessor(torch.nn.Module):
    """Runs a (reference-API) ``Processor`` on one rank's share of the graph: ghosts are refreshed before
    every block; the block itself is unchanged (it sees ``[owned | ghosts]`` as ``[mesh | hyper]`` rows).

    With an ``OverlapPlan`` (edges listed interior-first), bf16 latents, plain ``GraphNet`` blocks, the 'sum' aggregator and one
    edge set, ``HGN_HALO_OVERLAP=1`` makes every block run ``_PartitionedEdgeUpdate`` -- the halo exchange travels behind the
    interior edge tiles -- followed by the projected node update on the owned rows.  Opt-in: it is validated against the generic
    path (scripts/check_overlap.py, 2 GPUs) but measured SLOWER at N = 8 (28.4 vs 21.2 ms per step): the persistent edge kernels
    occupy every SM, so the NCCL send/recv kernels get no SM until a compute CTA retires and each side ends up waiting for the
    other; it needs compute grids that leave SMs to the communication kernels."""

    def __init__(self, processor: torch.nn.Module, plan: HaloPlan, overlap: Optional[OverlapPlan] = None):
        super().__init__()
        self.processor = processor
        self.plan = plan
        self.overlap = overlap

    def _can_overlap(self, owned, edge_sets) -> bool:
        import os
        from .migration.graphnet import GraphNet
        blocks = self.processor.graphnet_blocks
        return (self.overlap is not None and os.environ.get("HGN_HALO_OVERLAP", "0") == "1" and owned.is_cuda
                and len(edge_sets) == 1 and all(type(b) is GraphNet and b.message_passing_aggregator == "sum"
                                                and list(b.edge_models.keys()) == [edge_sets[0].name] for b in blocks))

    def forward(self, owned: torch.Tensor, edge_sets):
        from .util import MultiGraph
        precision = getattr(self.processor, "precision", None)
        in_dtype = owned.dtype
        if precision == "bf16":
            owned = owned.to(torch.bfloat16)
            edge_sets = [es._replace(features=es.features.to(torch.bfloat16)) for es in edge_sets]
        if precision == "bf16" and self._can_overlap(owned, edge_sets):
            from . import ops as _ops
            from .migration.graphnet import _mlp_parameters, _packed_cache
            es = edge_sets[0]
            e = es.features
            for block in self.processor.graphnet_blocks:
                edge_model = block.edge_models[es.name]
                ep = _mlp_parameters(edge_model, 3 * owned.shape[1], owned)
                with torch.cuda.device(owned.device):
                    packed = _ops._pack_weights(_packed_cache(edge_model), torch.bfloat16, 3, ep)
                e, agg = _PartitionedEdgeUpdate.apply(owned, e, *ep, packed, self.overlap)
                np_ = _mlp_parameters(block.node_model_cross, 2 * owned.shape[1], owned)
                owned = _ops.node_update(np_, _packed_cache(block.node_model_cross), owned, agg)
            return owned.to(in_dtype), [es._replace(features=e)]
        graph = MultiGraph([owned, None], list(edge_sets))
        for block in self.processor.graphnet_blocks:
            graph.node_features[1] = halo_exchange(graph.node_features[0], self.plan)
            graph = block(graph)
        out_nodes = graph.node_features[0].to(in_dtype)
        return out_nodes, list(graph.edge_sets)      # edge latents stay in the compute dtype (see Processor.forward)


def allreduce_gradients(module: torch.nn.Module, group=None) -> None:
    """Sum the per-rank partial weight gradients (one flat fp32 all-reduce)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


_NORMALIZER_FIELDS = ("_acc_sum", "_acc_sum_squared", "_acc_count", "_num_accumulations")


def allreduce_normalizers(normalizers, group=None) -> None:
    """Keep the replicas' online feature statistics identical: every ``Normalizer``
    (src/migration/normalizer.py:53-63) accumulates its own rank's batches; this sums, over the ranks, what each accumulated since
    the previous call and adds it to the common state -- the totals a single process that had seen every rank's batches would hold
    (``_num_accumulations`` advances by the world size per step, like that process's counter).  One flat all-reduce for all of them."""
    normalizers = list(normalizers)
    if not normalizers:
        return
    bases, deltas = [], []
    for nz in normalizers:
        base = getattr(nz, "_synced_state", None)
        if base is None:
            base = {f: torch.zeros_like(getattr(nz, f)) for f in _NORMALIZER_FIELDS}
        bases.append(base)
        deltas.extend((getattr(nz, f) - base[f]).reshape(-1).float() for f in _NORMALIZER_FIELDS)
    flat = torch.cat(deltas)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for nz, base in zip(normalizers, bases):
        state = {}
        for f in _NORMALIZER_FIELDS:
            cur = getattr(nz, f)
            state[f] = base[f] + flat[off:off + cur.numel()].view_as(cur).to(cur.dtype)
            off += cur.numel()
            setattr(nz, f, state[f])
        nz._synced_state = {f: t.clone() for f, t in state.items()}


# ------------------------------------------------------------------------------------------------------
# multi-GPU benchmark (bench.py --gpus N under torchrun)
# ------------------------------------------------------------------------------------------------------
def bench_partitioned(args, world, rank, dev, width, height, layers, metric, unit, peaks, clock_sampler_cls, roofline_fn=None):
    import json
    import os
    from . import ops, synthetic
    from .migration.meshgraphnet import MeshGraphNet
    from .util import EdgeSet

    n = width * height
    senders, receivers = synthetic.grid_edges_two_way(width, height)
    e_total = senders.numel()
    part = block_partition(n, world)
    lg = build_local_graph(senders, receivers, part, rank, world, interior_first=True)
    gen = torch.Generator().manual_seed(0)
    v0 = torch.randn(n, 128, generator=gen)[lg.owned]
    e0 = torch.randn(e_total, 128, generator=gen)[lg.edge_ids]
    coef = torch.randn(n, 128, generator=gen)[lg.owned].to(dev)
    weights = synthetic.seeded_state_dict(synthetic.processor_shapes(layers, ["mesh_edges"], "sum"), seed=17)
    shell = MeshGraphNet(3, 128, 2, "sum", layers, "none", ["mesh_edges"])
    proc = shell.processor
    proc.load_state_dict({k[len("processor."):]: t for k, t in weights.items()})
    proc = proc.to(dev)
    proc.precision = "bf16"
    plan = HaloPlan(lg, dev)
    model = PartitionedProcessor(proc, plan, OverlapPlan(lg, plan, dev))
    params = list(proc.parameters())
    v_dev, e_dev = v0.to(dev), e0.to(dev)
    v_host, e_host = v0.pin_memory(), e0.pin_memory()
    s_loc, r_loc = lg.senders.to(dev), lg.receivers.to(dev)
    loss_host = torch.empty(1).pin_memory()

    def step(v_in, e_in):
        for p in params:
            p.grad = None
        v = v_in.requires_grad_(True)
        ed = e_in.requires_grad_(True)
        out_v, out_sets = model(v, [EdgeSet("mesh_edges", ed, s_loc, r_loc)])
        loss = (out_v * coef).sum() + out_sets[0].features.sum() * 1e-3
        loss.backward()
        allreduce_gradients(proc)
        return loss

    def timed(fn, steps):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step(v_dev.detach(), e_dev.detach())
    from . import _cabi
    launches0 = ops.launch_count
    with clock_sampler_cls(dev.index) as clocks:
        ms_per_step = timed(lambda: step(v_dev.detach(), e_dev.detach()), args.steps)
    launches = ops.launch_count - launches0
    # per-kernel times from a second pass of the same steps: at several ranks the kernels are short and the two CUDA events the
    # library records around every launch would make the timed region host-bound
    _cabi.profile(True)
    timed(lambda: step(v_dev.detach(), e_dev.detach()), args.steps)
    kernels = _cabi.profile_report()
    _cabi.profile(False)
    roofline = roofline_fn(kernels, args.steps, int(lg.edge_ids.numel()), int(lg.owned.numel()), peaks) if roofline_fn else None

    copy_stream = torch.cuda.Stream(device=dev)

    def upload():
        with torch.cuda.stream(copy_stream):
            v, ed = v_host.to(dev, non_blocking=True), e_host.to(dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return v, ed, ready

    e2e_steps = max(2, min(args.steps, 8))

    def run_e2e(steps=e2e_steps):
        """Every step's inputs cross PCIe inside the region (step i+1's copy overlaps step i) and every loss is read back."""
        nxt = upload()
        for i in range(steps):
            v, ed, ready = nxt
            if i + 1 < steps:
                nxt = upload()
            torch.cuda.current_stream().wait_event(ready)
            v.record_stream(torch.cuda.current_stream()); ed.record_stream(torch.cuda.current_stream())
            loss = step(v, ed)
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    run_e2e(3)                                  # warm-up: two input sets are alive at a time, let the allocator cache both
    e2e_ms = timed(run_e2e, 1) / e2e_steps
    halo_rows = torch.tensor([float(lg.n_ghost)], device=dev)
    dist.all_reduce(halo_rows, op=dist.ReduceOp.MAX)
    if rank == 0:
        value = e_total * layers / (ms_per_step * 1e-3)
        print(json.dumps({
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (seeded 1000x1000 triangulated grid, seeded weights)",
            "config": {"workload": "cfg5: 1M-node / 5 992 002-edge triangulated mesh, 15 GraphNet layers, sum aggregator, processor fwd+bwd",
                       "nodes": n, "edges": e_total, "layers": layers, "latent": 128,
                       "partitioning": f"edge-cut, {world} row slabs, receiver-owner rule, halo of sender latents per layer "
                                       f"(max {int(halo_rows)} ghost rows per rank), NCCL grouped send/recv "
                                       f"{'overlapped with the interior edge tiles' if model._can_overlap(v_dev.to(torch.bfloat16), [EdgeSet('mesh_edges', e_dev, s_loc, r_loc)]) else 'before each block'}, "
                                       "weight-gradient all-reduce",
                       "l2_policy": "inputs larger than L2"},
            "e2e": {"value": e_total * layers / (e2e_ms * 1e-3), "unit": unit, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int((v_host.numel() + e_host.numel()) * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks.summary(),
            "roofline": roofline, "cpu_baseline": None,
            "kernels": [{"name": k["name"], "launches": k["launches"], "ms_per_step": k["ms"] / args.steps} for k in kernels],
            "kernel_times": "second pass of the same steps with the library's per-launch CUDA events enabled",
        }))
    dist.destroy_process_group()
