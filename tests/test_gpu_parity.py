"""GPU parity tests (run on the B200 with ``-m gpu``): the CUDA path, called through the Python mirror
of the reference API and through the raw C ABI, against the committed golden vectors (produced by the
live reference) and against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): fp32 mode 1e-5 relative, bf16 mode 2e-2 relative, with
relative error = max|a-b| / max|b| per tensor; index structure bit-exact.
"""
import os

import numpy as np
import pytest
import torch

import hgn_oracle as orc
from conftest import GOLDEN_DIR, GoldenCase, MODEL_CASES, rel_err, rel_l2
from hgn_b200 import _cabi, ops, synthetic
from hgn_b200 import util as hutil
from hgn_b200.migration.meshgraphnet import MeshGraphNet
from hgn_b200.plan import segment_plan

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "bf16": 2e-2}
# fp32: max metric.  bf16: L2 metric against the FP32 reference; the bound is loose on purpose -- with ~0.3% of
# the ReLU pre-activations inside bf16 rounding noise of zero, about half of all rows have one of their ~64 active
# hidden units switched relative to fp32, which moves that row's data gradient by ~1/sqrt(64).  Measured on the B200
# (printed per case with -s; profiles/r2_parity_measured.txt): 0.099 (mgn_pna_L1) .. 0.179 (hgn_hyper_sum_L2) on the one-tile
# goldens, 0.18 .. 0.235 on the 300-node ones, 0.279 on hgn_multiscale_sum_L1 -- the bound sits 8 % above the worst case.
# Kernel correctness of the bf16 backward is pinned separately, at 1e-2, by test_bf16_kernels_vs_bf16_emulation (same
# rounding points => same ReLU masks; 3-4e-3 measured on the full cfg5 mesh).
GRAD_TOL = {"fp32": 5e-5, "bf16": 3e-1}


def grad_err(a, b, precision):
    return rel_err(a, b) if precision == "fp32" else rel_l2(a, b)


def fp32_grad_check(a, b, aggregation, what):
    """fp32 gradients against the reference: max metric -- except that with a max / min aggregate ('pna', 'max', 'min') a handful of
    elements may take another route: the forward values agree to ~1e-6, so where two candidates of a segment lie closer than that, the
    GPU and the CPU pick different winners and the gradient of that (segment, column) goes to another edge (both are valid
    subgradients; torch_scatter's own CPU and CUDA reducers differ the same way).  One rerouted entry moves the whole rows of the dense
    node-level gradients that the encoder / node MLPs spread it over, so the share of elements beyond the tolerance depends on the
    row count of the tensor; the checks are therefore 5e-3 in the max metric (100 x the tolerance: a wrong gradient is O(1)), 5e-3 in
    relative L2, and at most 15 % of the elements beyond the fp32 tolerance."""
    err = rel_err(a, b)
    if err < GRAD_TOL["fp32"]:
        return False
    assert aggregation in ("pna", "max", "min"), f"{what}: {err:.3e}"
    a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
    beyond = (a64 - b64).abs() > GRAD_TOL["fp32"] * b64.abs().max()
    outliers = float(beyond.double().mean())
    bad_rows = int(beyond.reshape(beyond.shape[0], -1).any(dim=1).sum())
    # measured on the B200 (hgn_multiscale_pna_L1_300): mesh-node gradient 2.1 % of the elements (6-7 of 300 rows) beyond the 5e-5
    # tolerance, max metric 1.3e-3, L2 3.4e-4; hyper-node gradient (16 rows) 7.8 % of the elements, max metric 4.0e-4, L2 1.6e-4
    assert err < 5e-3 and rel_l2(a, b) < 5e-3 and outliers < 0.15, \
        f"{what}: max metric {err:.3e}, {outliers:.2e} of the elements / {bad_rows} rows beyond tolerance, L2 {rel_l2(a, b):.3e}"
    print(f"\n{what}: a max / min winner was rerouted: max metric {err:.3e}, {outliers:.2e} of the elements beyond tolerance, L2 {rel_l2(a, b):.3e}")
    return True


def _build(case, precision):
    m = MeshGraphNet(3, 128, 2, case.meta["aggregation"], case.meta["steps"], case.meta["architecture"], case.meta["edge_sets"])
    m.load_state_dict(case.weights())
    m = m.cuda()
    m.processor.precision = precision
    return m


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_matches_reference_golden(name, precision):
    case = GoldenCase(name)
    m = _build(case, precision)
    g = case.graph(hutil.MultiGraph, hutil.EdgeSet, device="cuda", requires_grad=True)
    latent_in = m.encoder(g)
    latent_out = m.processor(latent_in)
    out = m.decoder(latent_out._replace(node_features=latent_out.node_features[0]))
    tol = TOL[precision]
    assert rel_err(out, case.arr("output")) < tol
    for i, t in enumerate(latent_out.node_features):
        assert rel_err(t, case.arr(f"proc_node_{i}")) < tol, f"node latents {i}"
    assert [es.name for es in latent_out.edge_sets] == case.meta["proc_edge_sets"]
    for es in latent_out.edge_sets:
        if f"proc_edge_{es.name}" in case.z:       # the 300-node fixtures leave the mesh-edge latents out (size)
            assert rel_err(es.features, case.arr(f"proc_edge_{es.name}")) < tol, es.name
    coef = synthetic.seeded_tensor("loss_coef", out.shape, 3).cuda()
    (out * coef).sum().backward()
    gtol = GRAD_TOL[precision]
    agg = case.meta["aggregation"]
    rerouted = False
    worst_bf16 = 0.0
    for i, nf in enumerate(g.node_features):
        if precision == "fp32":
            rerouted |= fp32_grad_check(nf.grad, case.arr(f"grad_node_features_{i}"), agg, f"grad node {i}")
        else:
            err = grad_err(nf.grad, case.arr(f"grad_node_features_{i}"), precision)
            worst_bf16 = max(worst_bf16, err)
            assert err < gtol, f"grad node {i}: {err:.3e}"
    for es in g.edge_sets:
        key = f"grad_edge_{es.name}_features"
        if key in case.z:
            if precision == "fp32":
                rerouted |= fp32_grad_check(es.features.grad, case.arr(key), agg, key)
            else:
                err = grad_err(es.features.grad, case.arr(key), precision)
                worst_bf16 = max(worst_bf16, err)
                assert err < gtol, f"{key}: {err:.3e}"
    if precision == "bf16":
        print(f"\nbf16 input-gradient error vs the fp32 reference golden [{name}]: relative L2 {worst_bf16:.3e} (bound {gtol:.2e})")
    if rerouted:
        gtol = 5e-3                                # parameter gradients: sums over all rows, a rerouted element moves them by O(1/rows)
    params = dict(m.named_parameters())
    for key, ref in case.meta["grad_proj"].items():
        gflat = params[key].grad.double().reshape(-1).cpu()
        norm = ref[-1]
        assert abs(float(gflat.norm()) - norm) <= gtol * max(norm, 1e-6), key
        for i, val in enumerate(ref[:-1]):
            proj = float(torch.dot(gflat, synthetic.seeded_tensor(f"proj{i}:{key}", gflat.shape, 11).double()))
            assert abs(proj - val) <= gtol * max(norm * np.sqrt(gflat.numel()), 1e-6), (key, i)


def test_segment_ops_match_reference_golden():
    z = np.load(f"{GOLDEN_DIR}/segment_ops.npz")
    ids = torch.from_numpy(z["ids"])           # CPU ids, like the reference's callers pass them
    S = int(z["num_segments"])
    for op in ("sum", "mean", "max", "min", "std"):
        x = torch.from_numpy(z["data"]).cuda().requires_grad_(True)
        out = hutil.unsorted_segment_operation(x, ids, S, op)
        tol = 1e-6 if op != "std" else 1e-5
        assert torch.allclose(out.cpu(), torch.from_numpy(z[f"out_{op}"]), rtol=tol, atol=tol), op
        if op in ("max", "min"):
            assert torch.equal(out.detach().cpu(), torch.from_numpy(z[f"out_{op}"]))
        if op != "std":
            (out * torch.from_numpy(z["grad_up"]).cuda()).sum().backward()
            assert torch.allclose(x.grad.cpu(), torch.from_numpy(z[f"grad_{op}"]), rtol=1e-6, atol=1e-6), op
        out1 = hutil.unsorted_segment_operation(torch.from_numpy(z["data1"]).cuda(), ids.cuda(), S, op)
        assert torch.allclose(out1.cpu(), torch.from_numpy(z[f"out1_{op}"]), rtol=tol, atol=tol), op
    with pytest.raises(Exception, match="Invalid operation type"):
        hutil.unsorted_segment_operation(torch.zeros(3, 2).cuda(), torch.zeros(3, dtype=torch.int64), 2, "median")
    with pytest.raises(AssertionError):
        hutil.unsorted_segment_operation(torch.zeros(3, 2).cuda(), torch.zeros(5, dtype=torch.int64), 2, "sum")
    with pytest.raises(_cabi.HgnError):      # id outside [0, S)
        hutil.unsorted_segment_operation(torch.zeros(3, 4).cuda(), torch.tensor([0, 5, 1]).cuda(), 2, "sum")


def test_csr_plan_is_stable_and_bit_exact():
    rng = np.random.default_rng(3)
    for E, S in ((0, 5), (1, 1), (1000, 37), (9282, 1600), (200000, 3)):
        ids = torch.from_numpy(rng.integers(0, S, size=E).astype(np.int64))
        plan = segment_plan(ids.cuda(), S)
        perm_ref = torch.sort(ids, stable=True).indices.to(torch.int32)
        rowptr_ref = torch.cat([torch.zeros(1, dtype=torch.int64), torch.bincount(ids, minlength=S).cumsum(0)]).to(torch.int32)
        assert torch.equal(plan.rowptr.cpu(), rowptr_ref)
        if E:
            assert torch.equal(plan.perm.cpu()[:E], perm_ref)
            assert torch.equal(plan.ids32.cpu()[:E], ids.to(torch.int32))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_segment_kernels_vs_oracle_ragged(dtype):
    """Ragged / empty / extreme fan-in segments (hyper-node pooling: N edges into a handful of rows)."""
    rng = np.random.default_rng(11)
    for E, S, D in ((0, 4, 128), (5000, 1600, 128), (3000, 7, 128), (64, 64, 128), (500, 50, 4), (300, 20, 3)):
        ids = torch.from_numpy(rng.integers(0, max(S - 1, 1), size=E).astype(np.int64))
        x_cpu = torch.from_numpy(rng.standard_normal((E, D)).astype(np.float32)).to(dtype).float()
        x = x_cpu.to(dtype).cuda().requires_grad_(True)
        plan = segment_plan(ids.cuda(), S)
        outs = ops.segment_aggregate(x, plan, ("sum", "mean", "max", "min"))
        xo = x_cpu.clone().requires_grad_(True)
        refs = [orc.segment_reduce(xo, ids, S, op) for op in ("sum", "mean", "max", "min")]
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        for o, r, op in zip(outs, refs, ("sum", "mean", "max", "min")):
            if E == 0:
                assert float(o.abs().max()) == 0.0 if o.numel() else True
                continue
            assert rel_err(o.float(), r) < tol, (E, S, D, op)
            if op in ("max", "min"):
                assert torch.equal(o.float().cpu(), r.detach())       # selections are exact in any precision
        if E == 0:
            continue
        gs = [torch.from_numpy(rng.standard_normal(tuple(r.shape)).astype(np.float32)) for r in refs]
        sum(((o.float() * g.cuda()).sum() for o, g in zip(outs, gs))).backward()
        sum(((r * g).sum() for r, g in zip(refs, gs))).backward()
        assert rel_err(x.grad.float(), xo.grad) < (1e-5 if dtype == torch.float32 else 2e-2), (E, S, D)


def _random_mlp_weights(n_chunks, seed):
    shapes = synthetic.mlp_shapes("m", 128 * n_chunks)
    return synthetic.seeded_state_dict(shapes, seed)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,n_nodes", [(1, 5), (63, 40), (64, 64), (129, 33), (1000, 300), (9282, 1600)])
def test_fused_edge_update_vs_oracle(rows, n_nodes, precision):
    """Edge update through the C ABI on ragged tile counts (1 row ... 145 tiles), forward and backward."""
    dtype = torch.float32 if precision == "fp32" else torch.bfloat16
    rng = np.random.default_rng(rows)
    w = _random_mlp_weights(3, 5)
    s = torch.from_numpy(rng.integers(0, n_nodes, size=rows).astype(np.int64))
    r = torch.from_numpy(rng.integers(0, n_nodes, size=rows).astype(np.int64))
    v_cpu = torch.from_numpy(rng.standard_normal((n_nodes, 128)).astype(np.float32)).to(dtype).float()
    e_cpu = torch.from_numpy(rng.standard_normal((rows, 128)).astype(np.float32)).to(dtype).float()
    gup = torch.from_numpy(rng.standard_normal((rows, 128)).astype(np.float32)).to(dtype).float()
    # oracle
    wo = {f"blk.edge_models.mesh_edges.{k[2:]}": t.clone().requires_grad_(True) for k, t in w.items()}
    vo, eo = v_cpu.clone().requires_grad_(True), e_cpu.clone().requires_grad_(True)
    ref = orc.edge_update(wo, "blk", [vo], orc.EdgeSet("mesh_edges", eo, s, r))
    (ref * gup).sum().backward()
    # CUDA
    params = [w[f"m.0.layers.linear_{k}.{p}"].cuda().requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
    params += [w["m.1.weight"].cuda().requires_grad_(True), w["m.1.bias"].cuda().requires_grad_(True)]
    v = v_cpu.to(dtype).cuda().requires_grad_(True)
    e = e_cpu.to(dtype).cuda().requires_grad_(True)
    sp, rp = segment_plan(s.cuda(), n_nodes), segment_plan(r.cuda(), n_nodes)
    out = ops.fused_mlp(params, {}, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], rows, resid_source=1)
    (out.float() * gup.cuda()).sum().backward()
    tol, gtol = TOL[precision], GRAD_TOL[precision]
    assert rel_err(out.float(), ref) < tol
    assert grad_err(e.grad.float(), eo.grad, precision) < gtol
    assert grad_err(v.grad.float(), vo.grad, precision) < gtol
    names = [f"blk.edge_models.mesh_edges.0.layers.linear_{k}.{p}" for k in range(3) for p in ("weight", "bias")]
    names += ["blk.edge_models.mesh_edges.1.weight", "blk.edge_models.mesh_edges.1.bias"]
    for p, nm in zip(params, names):
        assert grad_err(p.grad, wo[nm].grad, precision) < gtol, nm


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_processor_vs_oracle_flag_shape(precision):
    """40x40 cloth (N=1600, E=9282), 3 layers, pna: whole-processor parity against the CPU oracle."""
    s, r = synthetic.grid_edges_two_way(40, 40)
    n, E = 1600, s.numel()
    shapes = synthetic.processor_shapes(3, ["mesh_edges"], "pna")
    w = synthetic.seeded_state_dict(shapes, 21)
    v0 = synthetic.seeded_tensor("v0", (n, 128), 1)
    e0 = synthetic.seeded_tensor("e0", (E, 128), 1)
    ref = orc.processor(w, "pna", "none", orc.MultiGraph([v0], [orc.EdgeSet("mesh_edges", e0, s, r)]))
    from hgn_b200.migration.processor import Processor
    from hgn_b200.migration.graphnet import GraphNet
    shell = MeshGraphNet(3, 128, 2, "pna", 3, "none", ["mesh_edges"])
    proc = shell.processor
    proc.load_state_dict({k[len("processor."):]: t for k, t in w.items()})
    proc = proc.cuda()
    proc.precision = precision
    with torch.no_grad():
        out = proc(hutil.MultiGraph([v0.cuda()], [hutil.EdgeSet("mesh_edges", e0.cuda(), s, r)]))
    assert rel_err(out.node_features[0], ref.node_features[0]) < TOL[precision]
    assert rel_err(out.edge_sets[0].features, ref.edge_sets[0].features) < TOL[precision]


def test_fp32_results_are_deterministic():
    case = GoldenCase("mgn_pna_L1")
    m = _build(case, "fp32")
    outs = []
    for _ in range(2):
        m.zero_grad()
        g = case.graph(hutil.MultiGraph, hutil.EdgeSet, device="cuda", requires_grad=True)
        out = m(g)
        out.sum().backward()
        outs.append((out.detach().clone(), g.node_features[0].grad.clone(),
                     m.processor.graphnet_blocks[0].edge_models["mesh_edges"][0].layers.linear_0.weight.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)      # no float atomics anywhere: bitwise reproducible


def test_rows_gather_scatter_and_colsum():
    lib = _cabi.load()
    rng = np.random.default_rng(0)
    for dtype in (torch.float32, torch.bfloat16):
        src = torch.from_numpy(rng.standard_normal((500, 128)).astype(np.float32)).to(dtype).cuda()
        idx = torch.from_numpy(rng.permutation(500)[:77].astype(np.int32)).cuda()
        dst = torch.empty((77, 128), dtype=dtype, device="cuda")
        _cabi.check(lib.hgn_rows_gather(_cabi.dtype_code(dtype), src.data_ptr(), idx.data_ptr(), 77, 128, dst.data_ptr(), _cabi.stream_ptr()))
        assert torch.equal(dst, src[idx.long()])
        back = torch.zeros_like(src)
        _cabi.check(lib.hgn_rows_scatter(_cabi.dtype_code(dtype), dst.data_ptr(), idx.data_ptr(), 77, 128, back.data_ptr(), 0, _cabi.stream_ptr()))
        assert torch.equal(back[idx.long()], dst)
        cs = ops.colsum(src)
        assert rel_err(cs, src.float().sum(0)) < 1e-5


def _bf16_emulated_mlp(x, w):
    """torch fp32 arithmetic on bf16-rounded operands, rounding H1/H2 where the kernel does."""
    W0, b0, W1, b1, W2, b2, g, b = w
    r = lambda t: t.to(torch.bfloat16).float()
    h = torch.relu(x @ r(W0).t() + b0)
    h = torch.relu(r(h) @ r(W1).t() + b1)
    return torch.nn.functional.layer_norm(r(h) @ r(W2).t() + b2, (128,), g, b, 1e-5)


@pytest.mark.parametrize("rows,n_nodes", [(128, 64), (9282, 1600), (200000, 40000)])
def test_bf16_kernels_vs_bf16_emulation(rows, n_nodes):
    """The tcgen05 kernels against torch arithmetic with the SAME rounding points: isolates kernel bugs
    from the bf16-vs-fp32 model difference (forward max metric 1e-2 = bf16 output rounding)."""
    torch.manual_seed(rows)
    w = _random_mlp_weights(3, 9)
    params = [w[f"m.0.layers.linear_{k}.{p}"].cuda().requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
    params += [w["m.1.weight"].cuda().requires_grad_(True), w["m.1.bias"].cuda().requires_grad_(True)]
    s = torch.randint(0, n_nodes, (rows,), device="cuda")
    r = torch.randint(0, n_nodes, (rows,), device="cuda")
    v = torch.randn(n_nodes, 128, device="cuda").to(torch.bfloat16).requires_grad_(True)
    e = torch.randn(rows, 128, device="cuda").to(torch.bfloat16).requires_grad_(True)
    gup = torch.randn(rows, 128, device="cuda").to(torch.bfloat16)
    sp, rp = segment_plan(s, n_nodes), segment_plan(r, n_nodes)
    out = ops.fused_mlp(params, {}, [v, e], [ops.ChunkSpec(0, sp), ops.ChunkSpec(0, rp), ops.ChunkSpec(1)], rows, resid_source=1)
    out.backward(gup)
    wr = [p.detach().clone().requires_grad_(True) for p in params]
    vf, ef = v.detach().float().requires_grad_(True), e.detach().float().requires_grad_(True)
    ref = ef + _bf16_emulated_mlp(torch.cat([vf[s], vf[r], ef], -1), wr)
    ref.backward(gup.float())
    assert rel_err(out.float(), ref) < 1e-2
    assert rel_l2(e.grad.float(), ef.grad) < 1e-2 and rel_l2(v.grad.float(), vf.grad) < 1e-2
    for a, b in zip(params, wr):
        assert rel_l2(a.grad, b.grad) < 1e-2


def _bf16_emulated_projected_edge(v, e, s, r, w):
    """torch fp32 arithmetic with the rounding points of the projected edge kernels (csrc/edge_tc.cu): bf16 operands,
    bf16 per-node projection tables, bf16 H1/H2."""
    W0, b0, W1, b1, W2, b2, g, b = w
    rd = lambda t: t.to(torch.bfloat16).float()
    ps = rd(v @ rd(W0[:, :128]).t())
    pr = rd(v @ rd(W0[:, 128:256]).t())
    h = torch.relu(ps[s] + pr[r] + e @ rd(W0[:, 256:]).t() + b0)
    h = torch.relu(rd(h) @ rd(W1).t() + b1)
    return e + torch.nn.functional.layer_norm(rd(h) @ rd(W2).t() + b2, (128,), g, b, 1e-5)


@pytest.mark.parametrize("want_agg", [False, True])
@pytest.mark.parametrize("rows,n_nodes", [(1, 5), (63, 40), (128, 64), (129, 33), (1000, 300), (9282, 1600), (200000, 40000)])
def test_projected_edge_update_vs_bf16_emulation(rows, n_nodes, want_agg):
    """ops.edge_update (node projection + fused edge forward/backward kernels with the aggregate's gradient gathered in
    the kernel) against torch arithmetic with the same rounding points, on ragged tile counts."""
    torch.manual_seed(rows + int(want_agg))
    w = _random_mlp_weights(3, 11)
    params = [w[f"m.0.layers.linear_{k}.{p}"].cuda().requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
    params += [w["m.1.weight"].cuda().requires_grad_(True), w["m.1.bias"].cuda().requires_grad_(True)]
    s = torch.randint(0, n_nodes, (rows,), device="cuda")
    r = torch.randint(0, n_nodes, (rows,), device="cuda")
    v = torch.randn(n_nodes, 128, device="cuda").to(torch.bfloat16).requires_grad_(True)
    e = torch.randn(rows, 128, device="cuda").to(torch.bfloat16).requires_grad_(True)
    gup = torch.randn(rows, 128, device="cuda").to(torch.bfloat16)
    gagg = torch.randn(n_nodes, 128, device="cuda").to(torch.bfloat16)
    sp, rp = segment_plan(s, n_nodes), segment_plan(r, n_nodes)
    out, agg = ops.edge_update(params, {}, v, e, sp, rp, want_agg)
    loss = (out.float() * gup.float()).sum()
    if want_agg:
        loss = loss + (agg.float() * gagg.float()).sum()
    loss.backward()
    wr = [p.detach().clone().requires_grad_(True) for p in params]
    vf, ef = v.detach().float().requires_grad_(True), e.detach().float().requires_grad_(True)
    ref = _bf16_emulated_projected_edge(vf, ef, s, r, wr)
    ref_loss = (ref * gup.float()).sum()
    if want_agg:
        ref_agg = torch.zeros(n_nodes, 128, device="cuda").index_add_(0, r, ref)
        ref_loss = ref_loss + (ref_agg * gagg.float()).sum()
        assert rel_err(agg.float(), ref_agg) < 1e-2
    ref_loss.backward()
    assert rel_err(out.float(), ref) < 1e-2
    assert rel_l2(e.grad.float(), ef.grad) < 1e-2 and rel_l2(v.grad.float(), vf.grad) < 1.5e-2
    for a, b in zip(params, wr):
        assert rel_l2(a.grad, b.grad) < 1.5e-2


def _bf16_emulated_projected_node(v, aggs, w):
    """Rounding points of ops.node_update: bf16 operands, bf16 tables q1 = agg_1 Wa_1^T + agg_2 Wa_2^T, q2 = agg_3 Wa_3^T + agg_4 Wa_4^T,
    bf16 H1/H2."""
    W0, b0, W1, b1, W2, b2, g, b = w
    rd = lambda t: t.to(torch.bfloat16).float()
    pre = v @ rd(W0[:, :128]).t() + b0
    for t in range(0, len(aggs), 2):
        q = sum(aggs[j] @ rd(W0[:, 128 * (1 + j): 128 * (2 + j)]).t() for j in range(t, min(t + 2, len(aggs))))
        pre = pre + rd(q)
    h = torch.relu(pre)
    h = torch.relu(rd(h) @ rd(W1).t() + b1)
    return v + torch.nn.functional.layer_norm(rd(h) @ rd(W2).t() + b2, (128,), g, b, 1e-5)


@pytest.mark.parametrize("n_agg", [1, 2, 3, 4])
@pytest.mark.parametrize("n_nodes", [1, 63, 129, 1600, 40000])
def test_projected_node_update_vs_bf16_emulation(n_nodes, n_agg):
    """ops.node_update (aggregate projections + the fused edge kernels driven with identity indices) against torch arithmetic
    with the same rounding points, on ragged tile counts, for 1..4 aggregates ('sum' ... 'pna'); run twice for bit
    determinism."""
    torch.manual_seed(n_nodes + n_agg)
    w = _random_mlp_weights(1 + n_agg, 13)
    v0 = torch.randn(n_nodes, 128, device="cuda").to(torch.bfloat16)
    a0 = [torch.randn(n_nodes, 128, device="cuda").to(torch.bfloat16) for _ in range(n_agg)]
    gup = torch.randn(n_nodes, 128, device="cuda").to(torch.bfloat16)
    runs = []
    for _ in range(2):
        params = [w[f"m.0.layers.linear_{k}.{p}"].cuda().requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
        params += [w["m.1.weight"].cuda().requires_grad_(True), w["m.1.bias"].cuda().requires_grad_(True)]
        v = v0.clone().requires_grad_(True)
        aggs = [a.clone().requires_grad_(True) for a in a0]
        out = ops.node_update(params, {}, v, aggs)
        out.backward(gup)
        runs.append([out.detach(), v.grad] + [a.grad for a in aggs] + [p.grad for p in params])
    for a, b in zip(*runs):
        assert torch.equal(a, b)
    wr = [p.detach().clone().requires_grad_(True) for p in params]
    vf = v0.float().requires_grad_(True)
    af = [a.float().requires_grad_(True) for a in a0]
    ref = _bf16_emulated_projected_node(vf, af, wr)
    ref.backward(gup.float())
    assert rel_err(out.float(), ref) < 1e-2
    assert rel_l2(v.grad.float(), vf.grad) < 1e-2
    for a, b in zip(aggs, af):
        assert rel_l2(a.grad.float(), b.grad) < 1.5e-2
    for a, b in zip(params, wr):
        assert rel_l2(a.grad, b.grad) < 1.5e-2


def test_receiver_sorted_edge_storage_is_transparent():
    """HGN_EDGE_STORAGE=receiver_sorted (plan.EdgeStorageOrder): the bf16 processor keeps its edge rows sorted by receiver between
    entry and exit.  Forward results are bitwise those of the reference order (stable sort: every receiver's rows keep their order);
    gradients agree to rounding (the sender-side sums see another order)."""
    from hgn_b200 import config
    s, r = (t.cuda() for t in synthetic.grid_edges_two_way(40, 30))
    n, e = 1200, s.numel()
    w = synthetic.seeded_state_dict(synthetic.processor_shapes(3, ["mesh_edges"], "sum"), 21)
    v0 = synthetic.seeded_tensor("rs_v", (n, 128), 2).cuda()
    e0 = synthetic.seeded_tensor("rs_e", (e, 128), 2).cuda()
    runs = {}
    try:
        for mode in ("reference", "receiver_sorted"):
            config.set_edge_storage(mode)
            proc = MeshGraphNet(3, 128, 2, "sum", 3, "none", ["mesh_edges"]).processor
            proc.load_state_dict({k[len("processor."):]: t for k, t in w.items()})
            proc = proc.cuda()
            proc.precision = "bf16"
            v, ed = v0.clone().requires_grad_(True), e0.clone().requires_grad_(True)
            out = proc(hutil.MultiGraph([v], [hutil.EdgeSet("mesh_edges", ed, s, r)]))
            assert out.edge_sets[0].senders is s and out.edge_sets[0].receivers is r
            ((out.node_features[0] ** 2).sum() + out.edge_sets[0].features.float().sum()).backward()
            runs[mode] = (out.node_features[0].detach(), out.edge_sets[0].features.detach(), v.grad, ed.grad)
    finally:
        config.set_edge_storage("reference")
    a, b = runs["reference"], runs["receiver_sorted"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert rel_l2(b[2].float(), a[2].float()) < 2e-2 and rel_l2(b[3].float(), a[3].float()) < 2e-2


def test_projected_edge_update_is_deterministic():
    torch.manual_seed(3)
    w = _random_mlp_weights(3, 11)
    s = torch.randint(0, 3000, (20000,), device="cuda")
    r = torch.randint(0, 3000, (20000,), device="cuda")
    sp, rp = segment_plan(s, 3000), segment_plan(r, 3000)
    v0 = torch.randn(3000, 128, device="cuda").to(torch.bfloat16)
    e0 = torch.randn(20000, 128, device="cuda").to(torch.bfloat16)
    runs = []
    for _ in range(2):
        params = [w[f"m.0.layers.linear_{k}.{p}"].cuda().requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
        params += [w["m.1.weight"].cuda().requires_grad_(True), w["m.1.bias"].cuda().requires_grad_(True)]
        v, e = v0.clone().requires_grad_(True), e0.clone().requires_grad_(True)
        out, agg = ops.edge_update(params, {}, v, e, sp, rp, True)
        (out.float().sum() + (agg.float() ** 2).sum()).backward()
        runs.append([out.detach(), agg.detach(), v.grad, e.grad] + [p.grad for p in params])
    for a, b in zip(*runs):
        assert torch.equal(a, b)


def test_invalidate_packed_weights_after_a_write_through_data():
    """the staged weight copy follows Parameter._version; a write through ``p.data`` does not bump it, so the documented
    ``hgn_b200.invalidate_packed_weights()`` must make the kernels see the new weights."""
    import hgn_b200
    w = _random_mlp_weights(1, 11)
    params = [w[f"m.0.layers.linear_{k}.{p}"].cuda().requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
    params += [w["m.1.weight"].cuda().requires_grad_(True), w["m.1.bias"].cuda().requires_grad_(True)]
    x = torch.from_numpy(np.random.default_rng(5).standard_normal((300, 128)).astype(np.float32)).cuda()
    cache = {}
    with torch.no_grad():
        first = ops.fused_mlp(params, cache, [x], [ops.ChunkSpec(0)], 300, resid_source=0).clone()
        params[5].data.add_(torch.linspace(-1.0, 1.0, 128, device='cuda'))   # last linear's bias, behind the version counter's back
        hgn_b200.invalidate_packed_weights()
        second = ops.fused_mlp(params, cache, [x], [ops.ChunkSpec(0)], 300, resid_source=0)
    assert float((second - first).abs().max()) > 1e-3
    with torch.no_grad():
        fresh = ops.fused_mlp(params, {}, [x], [ops.ChunkSpec(0)], 300, resid_source=0)
    assert torch.equal(second, fresh)


@pytest.mark.parametrize("E,Sa,Sb", [(1, 1, 1), (1000, 300, 37), (9282, 1600, 1601), (200_000, 40_000, 33_333)])
def test_segment_sum_pair_equals_two_segment_sums_bitwise(E, Sa, Sb):
    """hgn_segment_sum_pair (one interleaved pass, sender- and receiver-keyed sums of the same rows) == two hgn_segment_reduce calls,
    bit for bit, for different segment counts, empty segments and ragged block counts."""
    lib = _cabi.load()
    rng = np.random.default_rng(E)
    x = torch.from_numpy(rng.standard_normal((E, 128)).astype(np.float32)).cuda().to(torch.bfloat16)
    ia = torch.from_numpy(rng.integers(0, max(Sa - 1, 1), size=E).astype(np.int64)).cuda()     # the last segment of A stays empty
    ib = torch.from_numpy(rng.integers(0, Sb, size=E).astype(np.int64)).cuda()
    pa, pb = segment_plan(ia, Sa), segment_plan(ib, Sb)
    st = _cabi.stream_ptr()
    ref_a, ref_b = torch.empty(Sa, 128, dtype=torch.bfloat16, device="cuda"), torch.empty(Sb, 128, dtype=torch.bfloat16, device="cuda")
    for plan, S, out in ((pa, Sa, ref_a), (pb, Sb, ref_b)):
        _cabi.check(lib.hgn_segment_reduce(_cabi.HGN_BF16, x.data_ptr(), E, 128, plan.perm.data_ptr(), plan.rowptr.data_ptr(), S, out.data_ptr(),
                                           None, None, None, None, None, 0, st), "hgn_segment_reduce")
    out_a, out_b = torch.full_like(ref_a, float("nan")), torch.full_like(ref_b, float("nan"))
    _cabi.check(lib.hgn_segment_sum_pair(_cabi.HGN_BF16, x.data_ptr(), E, 128, pa.perm.data_ptr(), pa.rowptr.data_ptr(), Sa, out_a.data_ptr(),
                                         pb.perm.data_ptr(), pb.rowptr.data_ptr(), Sb, out_b.data_ptr(), st), "hgn_segment_sum_pair")
    assert torch.equal(out_a.view(torch.int16), ref_a.view(torch.int16)) and torch.equal(out_b.view(torch.int16), ref_b.view(torch.int16))
    want = torch.zeros(Sa, 128, device="cuda").index_add_(0, ia, x.float())
    assert rel_err(out_a.float(), want) < 2e-2
