import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from hgn_b200 import ops
from hgn_b200.plan import segment_plan
from dev_tc import weights
dev = "cuda"
rows, n_nodes = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
w = [p.requires_grad_(True) for p in weights(3)]
s = torch.randint(0, n_nodes, (rows,), device=dev); r = torch.randint(0, n_nodes, (rows,), device=dev)
v = torch.randn(n_nodes, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
e = torch.randn(rows, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
sp, rp = segment_plan(s, n_nodes), segment_plan(r, n_nodes)
out = ops.fused_mlp(w, {}, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], rows, resid_source=1)
torch.cuda.synchronize(); print("fwd ok", flush=True)
out.backward(torch.randn(rows, 128, device=dev).to(torch.bfloat16))
torch.cuda.synchronize(); print("bwd ok", float(e.grad.float().abs().mean()), flush=True)
