"""torchrun --nproc-per-node 2 scripts/check_overlap.py : the overlapped partitioned path (interior edge tiles under the halo
exchange) against the generic partitioned path (exchange before each block) on the same rank-local data."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
from hgn_b200 import partition, synthetic
from hgn_b200.migration.meshgraphnet import MeshGraphNet
from hgn_b200.util import EdgeSet

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
W, H, L = 96, 64, 3
s, r = synthetic.grid_edges_two_way(W, H)
n, e_total = W * H, s.numel()
part = partition.block_partition(n, world)
lg = partition.build_local_graph(s, r, part, rank, world, interior_first=True)
gen = torch.Generator().manual_seed(0)
v0 = torch.randn(n, 128, generator=gen)[lg.owned].to(dev)
e0 = torch.randn(e_total, 128, generator=gen)[lg.edge_ids].to(dev)
coef = torch.randn(n, 128, generator=gen)[lg.owned].to(dev)
weights = synthetic.seeded_state_dict(synthetic.processor_shapes(L, ["mesh_edges"], "sum"), seed=17)
proc = MeshGraphNet(3, 128, 2, "sum", L, "none", ["mesh_edges"]).processor
proc.load_state_dict({k[len("processor."):]: t for k, t in weights.items()})
proc = proc.to(dev)
proc.precision = "bf16"
plan = partition.HaloPlan(lg, dev)
model = partition.PartitionedProcessor(proc, plan, partition.OverlapPlan(lg, plan, dev))
s_loc, r_loc = lg.senders.to(dev), lg.receivers.to(dev)
params = list(proc.parameters())

def run(overlap):
    os.environ["HGN_HALO_OVERLAP"] = "1" if overlap else "0"
    for p in params: p.grad = None
    v = v0.clone().requires_grad_(True); ed = e0.clone().requires_grad_(True)
    out_v, out_sets = model(v, [EdgeSet("mesh_edges", ed, s_loc, r_loc)])
    loss = (out_v * coef).sum() + (out_sets[0].features.float() ** 2).sum() * 1e-3
    loss.backward()
    partition.allreduce_gradients(proc)
    torch.cuda.synchronize()
    return [out_v.detach(), out_sets[0].features.detach().float(), v.grad, ed.grad] + [p.grad.clone() for p in params]

assert model._can_overlap(v0.to(torch.bfloat16), [EdgeSet("mesh_edges", e0, s_loc, r_loc)])
a, b = run(True), run(False)
worst = 0.0
for i, (x, y) in enumerate(zip(a, b)):
    err = float((x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30))
    worst = max(worst, err)
    assert err < 2e-3, (i, err)
print(f"rank {rank}: interior {lg.n_interior} / cut {lg.senders.numel() - lg.n_interior} edges, {lg.n_ghost} ghosts; "
      f"max rel l2 difference overlap vs generic = {worst:.2e}", flush=True)
dist.destroy_process_group()
