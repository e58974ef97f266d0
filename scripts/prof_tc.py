"""One fused edge update (forward + backward) on the cfg5 mesh, for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from hgn_b200 import ops, synthetic
from hgn_b200.plan import segment_plan
from dev_tc import weights
dev = "cuda"
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
s, r = synthetic.grid_edges_two_way(W, H)
n, E = W * H, s.numel()
s, r = s.to(dev), r.to(dev)
w = [p.requires_grad_(True) for p in weights(3)]
v = torch.randn(n, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
e = torch.randn(E, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
gup = torch.randn(E, 128, device=dev).to(torch.bfloat16)
sp, rp = segment_plan(s, n), segment_plan(r, n)
cache = {}
for it in range(2):
    out = ops.fused_mlp(w, cache, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], E, resid_source=1)
    out.backward(gup)
torch.cuda.synchronize()
print("done", E)
