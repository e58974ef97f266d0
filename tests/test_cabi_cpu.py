"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol the
header declares, the Python mirror keeps the reference's API/state_dict contract, and the product
path refuses to run without its CUDA kernels (no silent fallback)."""
import os
import pickle
import re

import pytest
import torch

from conftest import ROOT, GoldenCase, MODEL_CASES
from hgn_b200 import _cabi
from hgn_b200.migration.meshgraphnet import MeshGraphNet
from hgn_b200 import util as hutil


def _header_functions():
    text = open(os.path.join(ROOT, "include", "hgn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hgn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    declared = _header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hgn_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == declared, "ctypes signatures out of sync with the header"
    assert lib.hgn_abi_version() == _cabi.ABI_VERSION == 6


def test_size_queries_need_no_gpu():
    lib = _cabi.load()
    assert lib.hgn_mlp_packed_bytes(_cabi.HGN_F32, 3) >= 2 * (384 * 128 + 2 * 128 * 128) * 4
    assert lib.hgn_mlp_packed_bytes(_cabi.HGN_BF16, 3) >= (384 * 128 + 2 * 128 * 128) * 2
    assert lib.hgn_mlp_packed_bytes(_cabi.HGN_F32, 0) == 0
    assert lib.hgn_mlp_backward_workspace_bytes(_cabi.HGN_F32, 1000, 3) > 0
    assert lib.hgn_csr_workspace_bytes(1000, 100) > 3 * 4000
    assert lib.hgn_edge_update_backward_workspace_bytes(_cabi.HGN_BF16, 1000) > 3 * 128 * 128 * 4
    assert lib.hgn_node_update_backward_workspace_bytes(_cabi.HGN_BF16, 1000) > 1000 * 128 * 2
    assert lib.hgn_edge_update_backward_workspace_bytes(_cabi.HGN_F32, 1000) == 0      # bf16-only entry points say so
    small, big = lib.hgn_world_edges_workspace_bytes(1000, 6000), lib.hgn_world_edges_workspace_bytes(1000000, 6000000)
    assert 0 < small < big and big < 400 * 2 ** 20 and lib.hgn_world_edges_workspace_bytes(0, 0) > 0      # O(N): ~150 MiB at 1 M nodes, not N^2


@pytest.mark.parametrize("name", MODEL_CASES)
def test_state_dict_contract_matches_reference(name):
    case = GoldenCase(name)
    m = MeshGraphNet(3, 128, 2, case.meta["aggregation"], case.meta["steps"], case.meta["architecture"], case.meta["edge_sets"])
    assert list(m.state_dict().keys()) == list(case.meta["shapes"].keys())   # same keys, same order
    m.load_state_dict(case.weights())                                        # lazy linears take the reference shapes
    for k, v in m.state_dict().items():
        assert list(v.shape) == case.meta["shapes"][k]


def test_optimizer_built_before_first_forward_keeps_parameters():
    m = MeshGraphNet(3, 128, 2, "sum", 1, "none", ["mesh_edges"])
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)      # MeshSimulator.py:109-110
    held = {id(p) for g in opt.param_groups for p in g["params"]}
    case = GoldenCase("mgn_sum_L2")
    m2 = MeshGraphNet(3, 128, 2, "sum", 2, "none", ["mesh_edges"])
    m2.load_state_dict(case.weights())
    assert {id(p) for p in m.parameters()} == held


def test_module_pickles_like_the_reference_checkpoint():
    case = GoldenCase("mgn_sum_L2")
    m = MeshGraphNet(3, 128, 2, "sum", 2, "none", ["mesh_edges"])
    m.load_state_dict(case.weights())
    m.processor.graphnet_blocks[0].__dict__["_hgn_packed"] = {"x": object()}   # device cache must not be pickled
    m2 = pickle.loads(pickle.dumps(m))
    assert "_hgn_packed" not in m2.processor.graphnet_blocks[0].__dict__
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_namedtuple_contract():
    assert hutil.EdgeSet._fields == ("name", "features", "senders", "receivers")
    assert hutil.MultiGraph._fields == ("node_features", "edge_sets")
    assert hutil.MultiGraphWithPos._fields == ("node_features", "edge_sets", "target_feature", "mesh_features", "model_type",
                                               "node_dynamic", "unnormalized_edges", "obstacle_nodes")
    assert MeshGraphNet.get_architecture("hyper")[1] and not MeshGraphNet.get_architecture("anything-else")[1]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    case = GoldenCase("mgn_sum_L2")
    m = MeshGraphNet(3, 128, 2, "sum", 2, "none", ["mesh_edges"])
    m.load_state_dict(case.weights())
    g = case.graph(hutil.MultiGraph, hutil.EdgeSet)
    with pytest.raises(_cabi.HgnError):
        m(g)
    with pytest.raises(_cabi.HgnError):
        hutil.unsorted_segment_operation(torch.zeros(4, 2), torch.zeros(4, dtype=torch.int64), 2, "sum")
    with pytest.raises(AssertionError):
        hutil.unsorted_segment_operation(torch.zeros(3, 2), torch.zeros(5, dtype=torch.int64), 2, "sum")


def test_triangles_to_edges_bit_exact_vs_golden():
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "mesh_edges.npz"))
    d = hutil.triangles_to_edges(torch.from_numpy(z["tri_cells"]))
    assert torch.equal(d["two_way_connectivity"][0], torch.from_numpy(z["tri_senders"]))
    assert torch.equal(d["two_way_connectivity"][1], torch.from_numpy(z["tri_receivers"]))
    d = hutil.triangles_to_edges(torch.from_numpy(z["tet_cells"]), deform=True)
    assert torch.equal(d["two_way_connectivity"][0], torch.from_numpy(z["tet_senders"]))
    assert torch.equal(d["two_way_connectivity"][1], torch.from_numpy(z["tet_receivers"]))
