"""Times forward / backward of one fused edge update on the cfg5 mesh with the library's kernel timers."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from hgn_b200 import ops, synthetic, _cabi
from hgn_b200.plan import segment_plan
from dev_tc import weights
dev = "cuda"
s, r = synthetic.grid_edges_two_way(1000, 1000)
n, E = 1000000, s.numel()
s, r = s.to(dev), r.to(dev)
w = [p.requires_grad_(True) for p in weights(3)]
v = torch.randn(n, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
e = torch.randn(E, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
gup = torch.randn(E, 128, device=dev).to(torch.bfloat16)
sp, rp = segment_plan(s, n), segment_plan(r, n)
cache = {}
def it():
    out = ops.fused_mlp(w, cache, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], E, resid_source=1)
    out.backward(gup)
for _ in range(2): it()
_cabi.profile(True)
for _ in range(3): it()
rep = _cabi.profile_report()
print(os.environ.get("HGN_TC_ABLATE", "0"), {k["name"]: round(k["ms"] / 3, 3) for k in rep})
