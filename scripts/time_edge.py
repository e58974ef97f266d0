"""Times the projected edge update (forward / backward kernels) on the cfg5 mesh with the library's kernel timers."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
from hgn_b200 import ops, synthetic, _cabi
from hgn_b200.plan import segment_plan
dev = "cuda"
W, H = int(os.environ.get("GRID_W", 1000)), int(os.environ.get("GRID_H", 1000))
s, r = synthetic.grid_edges_two_way(W, H)
n, E = W * H, s.numel()
s, r = s.to(dev), r.to(dev)
sd = synthetic.seeded_state_dict(synthetic.mlp_shapes("m", 384), 3)
w = [sd[f"m.0.layers.linear_{k}.{p}"].to(dev).requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
w += [sd["m.1.weight"].to(dev).requires_grad_(True), sd["m.1.bias"].to(dev).requires_grad_(True)]
v = torch.randn(n, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
e = torch.randn(E, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
gup = torch.randn(E, 128, device=dev).to(torch.bfloat16)
gagg = torch.randn(n, 128, device=dev).to(torch.bfloat16)
sp, rp = segment_plan(s, n), segment_plan(r, n)
cache = {}
def it():
    out, agg = ops.edge_update(w, cache, v, e, sp, rp, True)
    torch.autograd.backward([out, agg], [gup, gagg])
for _ in range(2): it()
_cabi.profile(True)
for _ in range(3): it()
rep = _cabi.profile_report()
print(os.environ.get("HGN_TC_ABLATE", "0"), {k["name"]: round(k["ms"] / k["launches"], 3) for k in rep})
