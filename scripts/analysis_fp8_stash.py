"""Round-2 planning, CPU only: what does stashing the hidden activations H1 / H2 of the edge MLP as FP8 (e4m3) instead of bf16 do to
the weight gradients?  H1 / H2 (post-ReLU) feed only dW1 = dH2'^T H1, dW2 = dY^T H2 and the ReLU masks (exact: zero stays zero), so
their rounding is zero-mean noise averaged over all edges.  One edge-MLP layer on a seeded grid mesh, fp32 reference."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
from hgn_b200 import synthetic

torch.manual_seed(0)
W = int(os.environ.get("GRID_W", 300)); H = int(os.environ.get("GRID_H", 300))
s, r = synthetic.grid_edges_two_way(W, H)
n, E = W * H, s.numel()
sd = synthetic.seeded_state_dict(synthetic.mlp_shapes("m", 384), 3)
W0, b0, W1, b1, W2, b2 = (sd[f"m.0.layers.linear_{k}.{p}"] for k in range(3) for p in ("weight", "bias"))
v, e = torch.randn(n, 128), torch.randn(E, 128)
x = torch.cat([v[s], v[r], e], -1)
h1 = torch.relu(x @ W0.T + b0)
h2 = torch.relu(h1 @ W1.T + b1)
y = h2 @ W2.T + b2
dy = torch.randn(E, 128) / 128 ** 0.5                      # stand-in for the LayerNorm backward output
dh2 = (dy @ W2) * (h2 > 0)
def rel(a, b): return float((a - b).norm() / b.norm())
ref_dW2, ref_dW1 = dy.T @ h2, dh2.T @ h1
for name, q in (("bf16", lambda t: t.to(torch.bfloat16).float()), ("fp8 e4m3", lambda t: t.to(torch.float8_e4m3fn).float()),
                ("fp8 e5m2", lambda t: t.to(torch.float8_e5m2).float())):
    q1, q2 = q(h1), q(h2)
    print(f"{name:9s} E={E}: element error H1 {rel(q1, h1):.2e}  H2 {rel(q2, h2):.2e} | dW2 {rel(dy.T @ q2, ref_dW2):.2e}  dW1 {rel(dh2.T @ q1, ref_dW1):.2e}"
          f" | mask flips {int(((q1 > 0) != (h1 > 0)).sum() + ((q2 > 0) != (h2 > 0)).sum())}")
