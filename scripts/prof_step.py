"""torch.profiler view of one cfg5 bench step: which kernels (ours and torch's) fill the step, and the gaps."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
import bench
from hgn_b200.util import EdgeSet, MultiGraph
dev = torch.device("cuda", 0)
L = int(os.environ.get("LAYERS", 15))
data = bench.build_inputs(bench.GRID_W, bench.GRID_H, L)
proc = bench.make_processor(data["weights"], L, "bf16", dev)
params = list(proc.parameters())
s, r = data["senders"].to(dev), data["receivers"].to(dev)
v_dev, e_dev, coef = data["v0"].to(dev), data["e0"].to(dev), data["coef_v"].to(dev)
def step():
    for p in params: p.grad = None
    v = v_dev.detach().requires_grad_(True); ed = e_dev.detach().requires_grad_(True)
    out = proc(MultiGraph([v], [EdgeSet("mesh_edges", ed, s, r)]))
    loss = (out.node_features[0] * coef).sum() + out.edge_sets[0].features.sum() * 1e-3
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
