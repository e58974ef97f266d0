"""Whole-step CUDA-graph capture for the launch-bound small-mesh shapes.

A flag_simple-sized training step (1.6 k nodes x batch 21, 15 layers) is ~1 100 kernel launches of a few microseconds each;
eager PyTorch spends more time launching than the GPU spends computing.  ``GraphedStep`` captures one call of a step function
(forward, backward and -- if it is part of the function -- a capturable optimizer step) and replays it.  The library side is
capture-safe by construction: kernels run on the current (capturing) stream, workspaces come from torch's allocator (the graph's
private pool), segment plans are built during warm-up, and the weight-pack kernels are always recorded while capturing
(``ops._pack_weights``), so a replay re-reads the master weights an optimizer step inside the graph has changed.

    step = GraphedStep(lambda: loss_of(static_inputs), params)     # warm-up + capture
    static_inputs.copy_(batch); step()                              # replay

Inputs must live in static tensors that the caller overwrites in place between replays; gradients are (re)allocated inside the
graph's pool and stay valid until the next replay.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from . import _cabi


class GraphedStep:
    def __init__(self, step_fn: Callable[[], Optional[torch.Tensor]], params: Iterable[torch.nn.Parameter], warmup: int = 3):
        _cabi.require_cuda()
        self.params = list(params)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # lazy parameters, plans, kernel attributes, allocator warm-up
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        for p in self.params:
            p.grad = None                            # the captured backward allocates the gradients in the graph's pool
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = step_fn()

    def __call__(self):
        self.graph.replay()
        return self.result


class FlagRolloutGraph:
    """One closed-loop step of the reference's ``FlagModel.rollout`` (src/model/flag.py:194-246: build_graph with the model's
    normalisers, predict, ``update`` = second-order integration, pin every node that is not NORMAL) captured in ONE CUDA graph and
    replayed per step.  Same model object, same arithmetic as the eager reference loop; what is left out are its per-step host
    round trips (``torch.unique`` over the cells, ``F.one_hot``'s ``max()`` read-back, the ``.cpu().numpy()`` copy of the positions at
    flag.py:234, the normalisers' ``_num_accumulations`` comparison) -- the topology-only work is hoisted into
    ``graph_building.FlagGraphBuilder``.  ``node_dynamic`` (flag.py:102-115; read only by the intra-cluster sampling of the rmp
    clusterings) is not computed inside the graph, so ``_node_dynamic_normalizer`` does not advance during a graphed rollout."""

    def __init__(self, model, initial_state, warmup: int = 3):
        from .graph_building import FlagGraphBuilder
        from .util import NodeType
        _cabi.require_cuda(initial_state['world_pos'])
        self.model = model
        self.builder = FlagGraphBuilder(model, initial_state)
        node_type = initial_state['node_type']
        mask = torch.eq(node_type[:, 0], int(NodeType.NORMAL))
        self.mask = torch.stack((mask, mask, mask), dim=1)
        self.cur = initial_state['world_pos'].clone()
        self.prev = initial_state['prev|world_pos'].clone()
        cur0, prev0 = self.cur.clone(), self.prev.clone()

        def step():
            inputs = {'world_pos': self.cur, 'prev|world_pos': self.prev}
            graph = self.builder(inputs, False, node_dynamic=False)
            prediction = model.update(inputs, model(graph))
            nxt = torch.where(self.mask, torch.squeeze(prediction), torch.squeeze(self.cur))
            self.prev.copy_(self.cur)
            self.cur.copy_(nxt)

        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            self.cur.copy_(cur0)
            self.prev.copy_(prev0)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                step()
            self.cur.copy_(cur0)
            self.prev.copy_(prev0)

    def reset(self, world_pos: torch.Tensor, prev_world_pos: torch.Tensor) -> None:
        self.cur.copy_(world_pos)
        self.prev.copy_(prev_world_pos)

    @torch.no_grad()
    def rollout(self, num_steps: int) -> torch.Tensor:
        """``pred_pos [num_steps, N, 3]`` exactly as ``FlagModel.rollout`` stacks it (the position BEFORE each step, flag.py:244)."""
        out = torch.empty((num_steps,) + tuple(self.cur.shape), dtype=self.cur.dtype, device=self.cur.device)
        for i in range(num_steps):
            out[i].copy_(self.cur)
            self.graph.replay()
        return out


def flag_rollout_steps_per_second(model, traj, steps: int) -> dict:
    """bench.py helper: the graphed rollout's rate, and its agreement with ``model.rollout`` on the same trajectory."""
    import time
    initial_state = {k: torch.squeeze(v, 0)[0] for k, v in traj.items()}
    runner = FlagRolloutGraph(model, initial_state)
    check_steps = min(steps, 20)
    eager, _ = model.rollout({k: v[:check_steps] for k, v in traj.items()}, check_steps)
    got = runner.rollout(check_steps)
    ref = eager['pred_pos']
    travelled = (ref - ref[:1]).abs().max().clamp_min(1e-12)
    err = float((got - ref).abs().max() / travelled)
    runner.reset(initial_state['world_pos'], initial_state['prev|world_pos'])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    runner.rollout(steps)
    torch.cuda.synchronize()
    return {"value": steps / (time.perf_counter() - t0), "unit": "rollout steps/s", "steps": steps,
            "max_error_vs_FlagModel_rollout_relative_to_distance_travelled": err, "checked_steps": check_steps,
            "what": "FlagModel.rollout's step (same model object and normalisers) captured in one CUDA graph per step (hgn_b200.graphed.FlagRolloutGraph)"}


class FlagTrainingGraph:
    """One training iteration of the reference's loop for a flag-style model -- ``FlagModel.build_graph`` with its accumulating
    normalisers, ``training_step`` (encoder, processor, decoder, target normalisation, MSE over the NORMAL nodes; flag.py:143-154),
    ``loss.backward()`` and ``optimizer.step()`` (MeshSimulator.py:131-139) -- captured in ONE CUDA graph and replayed per frame.

    Same model object, normalisers and optimizer as the eager loop; two things are restated so that nothing returns to the host:
    the per-trajectory constants of ``build_graph`` are hoisted (``graph_building.FlagGraphBuilder``) and the masked loss
    ``mse(target[mask], out[mask])`` is computed as ``sum(mask * (target - out)^2) / (3 * count(mask))`` (the same number; boolean
    indexing has a data-dependent shape).  The optimizer must be capture-safe (``torch.optim.Adam(..., capturable=True)`` or SGD).
    The warm-up iterations torch needs before a capture are undone: weights, optimizer state and normaliser statistics are restored
    in place, so the first ``step`` is the first training iteration."""

    def __init__(self, model, frame, optimizer, warmup: int = 3):
        from .graph_building import FlagGraphBuilder
        from .util import NodeType
        _cabi.require_cuda(frame['world_pos'])
        self.model, self.optimizer = model, optimizer
        self.builder = FlagGraphBuilder(model, frame)
        self.params = [p for p in model.parameters() if p.requires_grad]
        mask = torch.eq(frame['node_type'][:, 0], int(NodeType.NORMAL))
        self.maskf = mask.to(torch.float32).unsqueeze(1)
        self.inv_count = 1.0 / (3.0 * float(mask.sum()))
        self.inputs = {k: frame[k].clone() for k in ('world_pos', 'prev|world_pos', 'target|world_pos')}
        self.loss = None

        def step():
            optimizer.zero_grad(set_to_none=True)
            graph = self.builder(self.inputs, True, node_dynamic=False)
            out = model(graph)
            target = model.get_target(self.inputs, True)
            diff = (target - out) * self.maskf
            loss = (diff * diff).sum() * self.inv_count
            loss.backward()
            optimizer.step()
            return loss.detach()

        # everything the warm-up iterations change, restored IN PLACE after the capture (the graph holds these tensors' addresses)
        with torch.no_grad():
            model(self.builder(self.inputs, False, node_dynamic=False))          # lazy linears, plans, packed weights
        saved_params = [p.detach().clone() for p in self.params]
        normalizers = [m for m in model.modules() if hasattr(m, '_acc_sum') and hasattr(m, '_num_accumulations')]
        fields = ('_acc_sum', '_acc_sum_squared', '_acc_count', '_num_accumulations')
        saved_norm = [[getattr(nz, f).clone() for f in fields] for nz in normalizers]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step()
        torch.cuda.current_stream().wait_stream(side)
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step()
        with torch.no_grad():
            for p, w in zip(self.params, saved_params):
                p.copy_(w)
            for nz, vals in zip(normalizers, saved_norm):
                for f, val in zip(fields, vals):
                    getattr(nz, f).copy_(val)
            for state in optimizer.state.values():
                for key, val in state.items():
                    if torch.is_tensor(val):
                        val.zero_()

    def step(self, frame) -> torch.Tensor:
        """One training iteration on ``frame`` (same mesh as the one given at construction); returns the loss (a device scalar that
        the next ``step`` overwrites)."""
        for k, buf in self.inputs.items():
            buf.copy_(frame[k])
        self.graph.replay()
        return self.loss
