"""Development check of the bf16 tcgen05 MLP-tile kernels against a torch fp32 computation on the
same bf16-rounded operands (GPU box only; not part of the test-suite)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
from hgn_b200 import ops, synthetic, _cabi
from hgn_b200.plan import segment_plan

torch.manual_seed(0)
dev = "cuda"

def ref_mlp(x, w):
    W0, b0, W1, b1, W2, b2, g, b = w
    r = lambda t: t.to(torch.bfloat16).float()
    h = torch.relu(x @ r(W0).t() + b0)
    h = torch.relu(r(h) @ r(W1).t() + b1)
    y = r(h) @ r(W2).t() + b2
    return torch.nn.functional.layer_norm(y, (128,), g, b, 1e-5)

def weights(nch, seed=3):
    sd = synthetic.seeded_state_dict(synthetic.mlp_shapes("m", 128 * nch), seed)
    return [sd[f"m.0.layers.linear_{k}.{p}"].to(dev) for k in range(3) for p in ("weight", "bias")] + [sd["m.1.weight"].to(dev), sd["m.1.bias"].to(dev)]

def check_edge(rows, n_nodes, backward=False):
    w = [p.requires_grad_(backward) for p in weights(3)]
    s = torch.randint(0, n_nodes, (rows,), device=dev); r = torch.randint(0, n_nodes, (rows,), device=dev)
    v = torch.randn(n_nodes, 128, device=dev).to(torch.bfloat16).requires_grad_(backward)
    e = torch.randn(rows, 128, device=dev).to(torch.bfloat16).requires_grad_(backward)
    sp, rp = segment_plan(s, n_nodes), segment_plan(r, n_nodes)
    out = ops.fused_mlp(w, {}, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], rows, resid_source=1)
    torch.cuda.synchronize()
    vf, ef = v.detach().float().requires_grad_(backward), e.detach().float().requires_grad_(backward)
    wr = [p.detach().clone().requires_grad_(backward) for p in w]
    x = torch.cat([vf[s], vf[r], ef], -1)
    ref = ef + ref_mlp(x, wr)
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    msg = f"edge rows={rows:>8} fwd rel_err={err:.3e}"
    if backward:
        gup = torch.randn(rows, 128, device=dev).to(torch.bfloat16)
        (out.float() * gup.float()).sum().backward()
        (ref * gup.float()).sum().backward()
        ge = float((e.grad.float() - ef.grad).abs().max() / ef.grad.abs().max())
        gv = float((v.grad.float() - vf.grad).abs().max() / vf.grad.abs().max())
        gw = max(float((a.grad - b.grad).abs().max() / b.grad.abs().max()) for a, b in zip(w, wr))
        msg += f" grad_e={ge:.3e} grad_v={gv:.3e} grad_w(max)={gw:.3e}"
    print(msg, flush=True)

def check_node(rows, nch, backward=False):
    w = [p.requires_grad_(backward) for p in weights(nch, 4)]
    srcs = [torch.randn(rows + 7, 128, device=dev).to(torch.bfloat16).requires_grad_(backward) for _ in range(nch)]
    out = ops.fused_mlp(w, {}, srcs, [ops.ChunkSpec(i, None, 5) for i in range(nch)], rows, resid_source=0, resid_offset=5)
    torch.cuda.synchronize()
    x = torch.cat([t.detach().float()[5:5 + rows] for t in srcs], -1)
    ref = srcs[0].detach().float()[5:5 + rows] + ref_mlp(x, [p.detach() for p in w])
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    print(f"node rows={rows:>8} nch={nch} fwd rel_err={err:.3e}", flush=True)

def time_edge(width=1000, height=1000, iters=5):
    s, r = synthetic.grid_edges_two_way(width, height)
    n, E = width * height, s.numel()
    s, r = s.to(dev), r.to(dev)
    w = weights(3)
    v = torch.randn(n, 128, device=dev).to(torch.bfloat16)
    e = torch.randn(E, 128, device=dev).to(torch.bfloat16)
    sp, rp = segment_plan(s, n), segment_plan(r, n)
    cache = {}
    f = lambda: ops.fused_mlp(w, cache, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], E, resid_source=1)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print(f"edge fwd E={E}: {ms:.3f} ms  -> {E*163840/ms/1e9:.1f} TFLOP/s, {E/ms/1e3:.1f} M edges/s", flush=True)

if __name__ == "__main__":
    bw = "--bwd" in sys.argv
    for rows, n in ((1, 5), (127, 40), (128, 64), (129, 33), (1000, 300), (9282, 1600), (300000, 50000)):
        check_edge(rows, n, bw)
    for rows, nch in ((100, 2), (1600, 2), (1600, 5), (31, 21), (20000, 9)):
        check_node(rows, nch, False)
    time_edge()
