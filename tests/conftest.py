"""pytest configuration: markers, import paths, golden-fixture loader."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "hyper-graph-nets_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MODEL_CASES = ["mgn_sum_L2", "mgn_pna_L1", "mgn_max_L1", "repeated_sum_L1", "multi_mean_L1",
               "hgn_hyper_pna_L1", "hgn_hyper_sum_L2", "hgn_hetero_pna_L1", "hgn_multiscale_sum_L1",
               # 300 nodes / 16 clusters / 13 mesh-edge tiles (round 2): the remote architectures above one 128-row tile
               "hgn_hyper_pna_L2_300", "hgn_hetero_pna_L2_300", "hgn_multiscale_pna_L1_300"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class GoldenCase:
    """One ``tests/golden/<name>.npz`` fixture produced by ``tests/golden/make_golden.py``."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        self.name = name

    def weights(self, device="cpu", dtype=torch.float32):
        from hgn_b200 import synthetic
        sd = synthetic.seeded_state_dict(self.meta["shapes"], self.meta["weight_seed"])
        return {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}

    def graph(self, graph_cls, edge_cls, device="cpu", requires_grad=False, dtype=torch.float32):
        nfs = [torch.from_numpy(self.z[f"node_features_{i}"]).to(device=device, dtype=dtype).requires_grad_(requires_grad)
               for i in range(self.meta["n_node_lists"])]
        sets = []
        for nm in self.meta["graph_edge_sets"]:
            sets.append(edge_cls(
                name=nm,
                features=torch.from_numpy(self.z[f"edge_{nm}_features"]).to(device=device, dtype=dtype).requires_grad_(requires_grad),
                senders=torch.from_numpy(self.z[f"edge_{nm}_senders"]).to(device),
                receivers=torch.from_numpy(self.z[f"edge_{nm}_receivers"]).to(device)))
        return graph_cls(nfs, sets)

    def arr(self, key):
        return torch.from_numpy(self.z[key])


@pytest.fixture(params=MODEL_CASES)
def golden_case(request):
    return GoldenCase(request.param)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|  -- the parity metric used throughout (SURVEY.md s4 tolerance hygiene)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 -- used for GRADIENTS in bf16 mode.  A ReLU unit whose pre-activation sits within
    the bf16 rounding noise of zero switches on/off relative to the fp32 reference; that changes isolated
    gradient rows by O(10%) (a property of bf16 training, not of the kernels), so gradients are compared in
    norm, while forward quantities (continuous in the inputs) keep the max metric."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
