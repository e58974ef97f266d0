"""hgn_b200.graph_building (device ``build_graph`` with the per-trajectory constants hoisted) against the LIVE
reference's ``FlagModel`` / ``CylinderModel`` / ``PlateModel.build_graph`` (src/model/flag.py:65-128, cylinder.py:65-106,
plate.py:69-200) on the CPU: every index list and every feature tensor bit for bit, over several frames of a trajectory with the
normalisers accumulating.  The two kernel-backed calls are replaced by their CPU restatements here (the reference's own
``unsorted_segment_operation``; the dense oracle for the world edges); the GPU run of the same comparison is part of
tests/test_dropin_gpu.py.  Runs in a subprocess (the reference is imported)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
ROOT = sys.argv[1]
for p in ("oracle", "hyper-graph-nets_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
os.environ.setdefault("WANDB_MODE", "disabled")
import torch
import reference_shim
reference_shim.load()
os.chdir(reference_shim.REFERENCE_ROOT)
import src.util as ref_util
import hgn_oracle as orc
from hgn_b200 import graph_building, util as our_util
import dropin_runner
from src.model.flag import FlagModel
from src.model.cylinder import CylinderModel
from src.model.plate import PlateModel

our_util.unsorted_segment_operation = ref_util.unsorted_segment_operation          # CPU stand-ins for the two kernel-backed calls
graph_building.world_edges = lambda pos, types, ms, mr: orc.world_edges(pos, types, ms, mr)


def same(a, b, what):
    if isinstance(a, torch.Tensor):
        assert isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), what
    elif isinstance(a, (list, tuple)) and not hasattr(a, "_fields"):
        assert len(a) == len(b), what
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, f"{what}[{i}]")
    elif hasattr(a, "_fields"):
        assert tuple(a._fields) == tuple(b._fields), what
        for f in a._fields:
            same(getattr(a, f), getattr(b, f), f"{what}.{f}")
    else:
        assert a == b, (what, a, b)


for case, cls in (("flag", FlagModel), ("cylinder", CylinderModel), ("plate", PlateModel)):
    params = dropin_runner.model_params("flag" if case == "flag" else "cylinder")       # no remote message passing: build_graph only
    frames = dropin_runner.make_frames(case, 4)
    ref_model, our_model = cls(params), cls(params)
    builder = graph_building.graph_builder(our_model, frames[0])
    n_world = []
    for t, frame in enumerate(frames):
        training = t < 3
        want = ref_model.build_graph(frame, training)
        got = builder(frame, training)
        same(got, want, f"{case} frame {t}")
        if case == "plate":
            n_world.append(int(want.edge_sets[1].senders.numel()))
    for name in ("_node_normalizer", "_mesh_edge_normalizer", "_node_dynamic_normalizer", "_world_edge_normalizer"):
        if hasattr(ref_model, name):
            a, b = getattr(ref_model, name), getattr(our_model, name)
            assert torch.equal(a._acc_sum, b._acc_sum) and torch.equal(a._acc_count, b._acc_count) and torch.equal(a._acc_sum_squared, b._acc_sum_squared), name
    if case == "plate":
        assert min(n_world) > 0 and len(set(n_world)) > 1, n_world        # the contact set is non-empty and changes along the trajectory
print("BUILDERS-OK")
'''


def test_device_graph_builders_equal_reference_build_graph():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_shim
    if not reference_shim.available():
        pytest.skip("no reference tree")
    out = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "BUILDERS-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
