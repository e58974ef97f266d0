import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from hgn_b200 import ops
from hgn_b200.plan import segment_plan
from dev_tc import weights, ref_mlp
dev = "cuda"

def run(rows, n_nodes, seed):
    torch.manual_seed(seed)
    s = torch.randint(0, n_nodes, (rows,), device=dev); r = torch.randint(0, n_nodes, (rows,), device=dev)
    v0 = torch.randn(n_nodes, 128, device=dev).to(torch.bfloat16)
    e0 = torch.randn(rows, 128, device=dev).to(torch.bfloat16)
    gup = torch.randn(rows, 128, device=dev).to(torch.bfloat16)
    sp, rp = segment_plan(s, n_nodes), segment_plan(r, n_nodes)
    res = []
    for rep in range(2):
        w = [p.clone().requires_grad_(True) for p in weights(3)]
        v = v0.clone().requires_grad_(True); e = e0.clone().requires_grad_(True)
        out = ops.fused_mlp(w, {}, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], rows, resid_source=1)
        out.backward(gup)
        torch.cuda.synchronize()
        res.append((e.grad.clone(), v.grad.clone(), [p.grad.clone() for p in w]))
    same_e = torch.equal(res[0][0], res[1][0]); same_v = torch.equal(res[0][1], res[1][1])
    same_w = all(torch.equal(a, b) for a, b in zip(res[0][2], res[1][2]))
    wr = [p.detach().clone().requires_grad_(True) for p in weights(3)]
    vf, ef = v0.float().requires_grad_(True), e0.float().requires_grad_(True)
    ref = ef + ref_mlp(torch.cat([vf[s], vf[r], ef], -1), wr)
    ref.backward(gup.float())
    d = (res[0][0].float() - ef.grad).abs()
    rowerr = d.max(1).values / ef.grad.abs().max()
    bad = (rowerr > 2e-2).nonzero().flatten()
    gw = [float((a - b.grad).abs().max() / b.grad.abs().max()) for a, b in zip(res[0][2], wr)]
    print(f"rows={rows} seed={seed} repeatable e/v/w={same_e}/{same_v}/{same_w} grad_e max={float(rowerr.max()):.3e} bad_rows={bad.numel()} first={bad[:8].tolist()} "
          f"cols_of_worst={d[int(rowerr.argmax())].topk(4).indices.tolist()} gw={['%.1e' % x for x in gw]}", flush=True)

for rows, n in ((128, 64), (128, 64), (127, 40), (129, 33), (256, 50), (1000, 300), (9282, 1600)):
    for seed in (0, 1, 2):
        run(rows, n, seed)
