"""World-edge search (src/model/plate.py:86-110): oracle vs the golden vectors recorded from the live reference, and the fp32
distance arithmetic the CUDA kernel restates (csrc/world_edges.cu: we_norm2 / we_distance) vs torch.cdist itself."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
import hgn_oracle as orc  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "world_edges.npz"))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_oracle_matches_reference_golden(case):
    pos = torch.from_numpy(GOLD[f"{case}_world_pos"])
    types = torch.from_numpy(GOLD[f"{case}_node_type"])
    cells = torch.from_numpy(GOLD[f"{case}_cells"])
    ms, mr = orc.triangles_to_edges(cells, deform=True)["two_way_connectivity"]
    assert np.array_equal(ms.numpy(), GOLD[f"{case}_mesh_edges_senders"]) and np.array_equal(mr.numpy(), GOLD[f"{case}_mesh_edges_receivers"])
    s, r = orc.world_edges(pos, types, ms, mr)
    assert s.dtype == torch.int64 and r.dtype == torch.int64
    assert np.array_equal(s.numpy(), GOLD[f"{case}_world_edges_senders"])
    assert np.array_equal(r.numpy(), GOLD[f"{case}_world_edges_receivers"])
    if case != "c":
        assert s.numel() > 0
        t = types.reshape(-1)
        assert bool((t[s] == 1).all()) and bool((t[r] == 0).all())           # OBSTACLE -> NORMAL only
        key = s * pos.shape[0] + r
        assert bool((key[1:] > key[:-1]).all())                              # torch.nonzero: row-major, unique
    else:
        assert s.numel() == 0


def _fma32(a, b, c):
    # fp32 fused multiply-add: the product of two fp32 values is exact in fp64
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def kernel_distance(x):
    """numpy restatement of we_norm2 / we_distance for all pairs (i = row, j = column)."""
    sq = (x * x).astype(np.float32)
    n2 = ((sq[:, 0] + sq[:, 1]).astype(np.float32) + sq[:, 2]).astype(np.float32)
    a = (-2.0 * x).astype(np.float32)
    acc = (a[:, None, 0] * x[None, :, 0]).astype(np.float32)
    acc = _fma32(np.broadcast_to(a[:, None, 1], acc.shape), np.broadcast_to(x[None, :, 1], acc.shape), acc)
    acc = _fma32(np.broadcast_to(a[:, None, 2], acc.shape), np.broadcast_to(x[None, :, 2], acc.shape), acc)
    acc = (n2[:, None] + acc).astype(np.float32)
    acc = (acc + n2[None, :]).astype(np.float32)
    return np.maximum(acc, np.float32(0))


@pytest.mark.parametrize("offset", [0.0, 5.0])
def test_kernel_distance_arithmetic_is_torch_cdist(offset):
    # the squared distance (before the square root) must be bit-identical to the GEMM inside torch.cdist
    torch.manual_seed(3)
    x = (torch.rand(700, 3) * torch.tensor([1.0, 0.5, 0.3]) + offset).float()
    n2 = x.pow(2).sum(-1, keepdim=True)
    one = torch.ones_like(n2)
    gemm = torch.cat([x.mul(-2), n2, one], -1).matmul(torch.cat([x, one, n2], -1).mT).clamp_min(0)
    assert torch.equal(gemm.sqrt(), torch.cdist(x, x, p=2))                  # this IS what cdist computes (mm path, > 25 rows)
    assert np.array_equal(kernel_distance(x.numpy()), gemm.numpy())
    # and the textbook form is NOT equivalent at the threshold level of rounding: it differs in many entries
    direct = (x[:, None, :] - x[None, :, :]).pow(2).sum(-1)
    assert int((direct != gemm).sum()) > 1000
