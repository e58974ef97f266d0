"""The fp32 parity mode against an fp32 denominator measured the same way as MEASURED_PEAKS.json: torch.matmul fp32
8192^3 with TF32 off (best of 10 and a 4 s sustained loop), then the processor in `precision = "fp32"` (FFMA kernels, the 1e-5 mode) on a
1000 x 250 slab of the cfg5 mesh, 3 layers, fwd+bwd.  Prints one JSON line.  No oracle, no reference: hgn_b200 only."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from hgn_b200.util import EdgeSet, MultiGraph

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); a @ b; t1.record(); torch.cuda.synchronize()
    best = max(best, 2 * 8192 ** 3 / (t0.elapsed_time(t1) * 1e-3) / 1e12)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps, wall = 0, time.perf_counter()
t0.record()
while time.perf_counter() - wall < 4.0:
    for _ in range(10): a @ b
    reps += 10
    torch.cuda.synchronize()
t1.record(); torch.cuda.synchronize()
sustained = reps * 2 * 8192 ** 3 / (t0.elapsed_time(t1) * 1e-3) / 1e12
del a, b

W, H, L = 1000, 250, 3
data = bench.build_inputs(W, H, L)
proc = bench.make_processor(data["weights"], L, "fp32", dev)
params = list(proc.parameters())
s, r = data["senders"].to(dev), data["receivers"].to(dev)
v0, e0, coef = data["v0"].to(dev), data["e0"].to(dev), data["coef_v"].to(dev)
def step():
    for p in params: p.grad = None
    v, ed = v0.detach().requires_grad_(True), e0.detach().requires_grad_(True)
    out = proc(MultiGraph([v], [EdgeSet("mesh_edges", ed, s, r)]))
    ((out.node_features[0] * coef).sum() + out.edge_sets[0].features.sum() * 1e-3).backward()
for _ in range(3): step()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5): step()
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 5
n, e = data["n"], data["e"]
flops = 3 * (e * bench.F_EDGE + n * bench.F_NODE_SUM) * L
print(json.dumps({"fp32_matmul_tflops_best": best, "fp32_matmul_tflops_sustained": sustained,
                  "fp32_mode": {"mesh": f"{W}x{H} ({n} nodes, {e} directed edges), {L} layers, sum, fwd+bwd", "ms_per_step": ms,
                                "edge_updates_per_s": e * L / (ms * 1e-3), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                                "frac_of_fp32_sustained": flops / (ms * 1e-3) / 1e12 / sustained}}))
