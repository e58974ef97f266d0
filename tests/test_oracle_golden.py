"""Pin the oracle (oracle/hgn_oracle.py) to the golden vectors produced by the live reference."""
import numpy as np
import pytest
import torch

import hgn_oracle as orc
from conftest import GoldenCase, rel_err
from hgn_b200 import synthetic

FP32_TOL = 2e-6   # same ATen kernels, different op grouping -> re-association noise only


def _run_oracle(case, dtype=torch.float32):
    w = {k: v.requires_grad_(True) for k, v in case.weights(dtype=dtype).items()}
    g = case.graph(orc.MultiGraph, orc.EdgeSet, requires_grad=True, dtype=dtype)
    arch, agg = case.meta["architecture"], case.meta["aggregation"]
    lat_in = orc.encode(w, arch, g)
    lat_out = orc.processor(w, agg, arch, lat_in)
    out = orc.mlp(w, "decoder.model", lat_out.node_features[0], layer_norm=False)
    loss = (out * synthetic.seeded_tensor("loss_coef", out.shape, 3).to(dtype)).sum()
    loss.backward()
    return w, g, lat_in, lat_out, out, loss


def test_oracle_forward_matches_reference(golden_case):
    w, g, lat_in, lat_out, out, loss = _run_oracle(golden_case)
    assert rel_err(out, golden_case.arr("output")) < FP32_TOL
    for i, t in enumerate(lat_in.node_features):
        assert rel_err(t, golden_case.arr(f"enc_node_{i}")) < FP32_TOL
    for i, t in enumerate(lat_out.node_features):
        assert rel_err(t, golden_case.arr(f"proc_node_{i}")) < FP32_TOL
    assert [es.name for es in lat_out.edge_sets] == golden_case.meta["proc_edge_sets"]
    for es in lat_out.edge_sets:
        if f"proc_edge_{es.name}" in golden_case.z:       # the 300-node fixtures leave the mesh-edge latents out (size)
            assert rel_err(es.features, golden_case.arr(f"proc_edge_{es.name}")) < FP32_TOL
    assert abs(float(loss) - golden_case.meta["loss"]) < 1e-4 * max(1.0, abs(golden_case.meta["loss"]))


def test_oracle_gradients_match_reference(golden_case):
    w, g, *_ = _run_oracle(golden_case)
    for i, nf in enumerate(g.node_features):
        assert rel_err(nf.grad, golden_case.arr(f"grad_node_features_{i}")) < 2e-5
    for es in g.edge_sets:
        key = f"grad_edge_{es.name}_features"
        if key in golden_case.z:
            assert rel_err(es.features.grad, golden_case.arr(key)) < 2e-5
    for key, ref in golden_case.meta["grad_proj"].items():
        gflat = w[key].grad.double().reshape(-1)
        norm = ref[-1]
        assert abs(float(gflat.norm()) - norm) <= 2e-5 * max(norm, 1e-6), key
        for i, val in enumerate(ref[:-1]):
            proj = float(torch.dot(gflat, synthetic.seeded_tensor(f"proj{i}:{key}", gflat.shape, 11).double()))
            assert abs(proj - val) <= 5e-5 * max(norm * np.sqrt(gflat.numel()), 1e-6), (key, i)
    for key in golden_case.meta["no_grad_params"]:
        assert w[key].grad is None or float(w[key].grad.abs().max()) == 0.0


def test_oracle_segment_ops_match_reference():
    z = np.load(f"{__import__('conftest').GOLDEN_DIR}/segment_ops.npz")
    ids = torch.from_numpy(z["ids"])
    S = int(z["num_segments"])
    for op in ("sum", "mean", "max", "min", "std"):
        x = torch.from_numpy(z["data"]).clone().requires_grad_(True)
        out = orc.segment_reduce(x, ids, S, op)
        tol = 1e-6 if op != "std" else 1e-5
        assert torch.allclose(out, torch.from_numpy(z[f"out_{op}"]), rtol=tol, atol=tol), op
        if op in ("max", "min"):
            assert torch.equal(out, torch.from_numpy(z[f"out_{op}"]))   # selection: bit-exact
        if op != "std":
            (out * torch.from_numpy(z["grad_up"])).sum().backward()
            assert torch.allclose(x.grad, torch.from_numpy(z[f"grad_{op}"]), rtol=1e-6, atol=1e-6), op
        out1 = orc.segment_reduce(torch.from_numpy(z["data1"]), ids, S, op)
        assert torch.allclose(out1, torch.from_numpy(z[f"out1_{op}"]), rtol=tol, atol=tol), op
    with pytest.raises(Exception, match="Invalid operation type"):
        orc.segment_reduce(torch.zeros(3, 2), torch.zeros(3, dtype=torch.int64), 2, "median")
    with pytest.raises(AssertionError):
        orc.segment_reduce(torch.zeros(3, 2), torch.zeros(5, dtype=torch.int64), 2, "sum")


def test_oracle_triangles_to_edges_bit_exact():
    z = np.load(f"{__import__('conftest').GOLDEN_DIR}/mesh_edges.npz")
    d = orc.triangles_to_edges(torch.from_numpy(z["tri_cells"]))
    assert torch.equal(d["two_way_connectivity"][0], torch.from_numpy(z["tri_senders"]))
    assert torch.equal(d["two_way_connectivity"][1], torch.from_numpy(z["tri_receivers"]))
    d = orc.triangles_to_edges(torch.from_numpy(z["tet_cells"]), deform=True)
    assert torch.equal(d["two_way_connectivity"][0], torch.from_numpy(z["tet_senders"]))
    assert torch.equal(d["two_way_connectivity"][1], torch.from_numpy(z["tet_receivers"]))
    assert d["two_way_connectivity"][0].dtype == torch.int64


def test_grid_edges_closed_form_matches_unique_path():
    for w, h in ((7, 5), (12, 9), (40, 40)):
        d = orc.triangles_to_edges(torch.from_numpy(synthetic.grid_triangles(w, h)))
        s, r = synthetic.grid_edges_two_way(w, h)
        assert torch.equal(s, d["two_way_connectivity"][0]) and torch.equal(r, d["two_way_connectivity"][1])
    s, r = synthetic.grid_edges_two_way(40, 40)
    assert s.numel() == 9282   # 40x40 cloth -> 9 282 directed edges
