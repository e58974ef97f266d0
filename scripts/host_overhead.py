"""Host-side enqueue time of one bench step (no synchronisation inside the timed region) and a cProfile of it."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
import bench
from hgn_b200.util import EdgeSet, MultiGraph
dev = torch.device("cuda", 0)
W = int(os.environ.get("GRID_W", 354)); H = int(os.environ.get("GRID_H", 354))      # ~1/8 of cfg5: the per-rank size at N = 8
data = bench.build_inputs(W, H, bench.LAYERS)
proc = bench.make_processor(data["weights"], bench.LAYERS, "bf16", dev)
params = list(proc.parameters())
s, r = data["senders"].to(dev), data["receivers"].to(dev)
v_dev, e_dev, coef = data["v0"].to(dev), data["e0"].to(dev), data["coef_v"].to(dev)
def step():
    for p in params: p.grad = None
    v = v_dev.detach().requires_grad_(True); ed = e_dev.detach().requires_grad_(True)
    out = proc(MultiGraph([v], [EdgeSet("mesh_edges", ed, s, r)]))
    loss = (out.node_features[0] * coef).sum() + out.edge_sets[0].features.sum() * 1e-3
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
host, total = [], []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    host.append((t1 - t0) * 1e3); total.append((t2 - t0) * 1e3)
print(f"grid {W}x{H}: host enqueue {min(host):.2f} ms, step (to sync) {min(total):.2f} ms")
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
