"""Drop-in boundary against the LIVE reference (build container only; skipped where /root/reference is not mounted):
``hgn_b200.install_as_reference_modules()`` makes the reference's own ``FlagModel`` (src/model/flag.py:36-63) build OUR
``MeshGraphNet`` through its unchanged constructor call, swaps ``util.unsorted_segment_operation``, and -- without a CUDA device --
the first kernel-backed call fails loudly instead of falling back to a CPU path.  Runs in a subprocess: it rewires sys.modules."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
ROOT = sys.argv[1]
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
os.environ.setdefault("WANDB_MODE", "disabled")
import torch
import reference_shim
reference_shim.load()                                   # stubs for torch_scatter & co, puts /root/reference on sys.path
os.chdir(reference_shim.REFERENCE_ROOT)
import hgn_b200
from hgn_b200 import _cabi, synthetic
hgn_b200.install_as_reference_modules()
import src.migration.meshgraphnet as ref_named
import hgn_b200.migration.meshgraphnet as ours
assert ref_named.MeshGraphNet is ours.MeshGraphNet
import src.util
import hgn_b200.util
assert src.util.unsorted_segment_operation is hgn_b200.util.unsorted_segment_operation
from make_golden import flag_params
from src.model.flag import FlagModel
model = FlagModel(flag_params("pna", 3))                # the reference's own system model, unmodified
net = model.learned_model
assert type(net) is ours.MeshGraphNet, type(net)
assert len(net.processor.graphnet_blocks) == 3 and net.processor.graphnet_blocks[0].message_passing_aggregator == "pna"
keys = [k for k, _ in net.named_parameters()]
assert any(k.startswith("processor.graphnet_blocks.2.edge_models.mesh_edges.0.layers.linear_0") for k in keys), keys[:5]
opt = torch.optim.Adam(net.parameters())                # MeshSimulator.py:109-110: built before the first forward
if not torch.cuda.is_available():
    frame = synthetic.cloth_frame(6, 5, seed=1)
    try:
        model.build_graph(frame, is_training=True)      # flag.py:102-113 calls the segment op -> kernels -> needs the GPU
    except _cabi.HgnError as exc:
        assert "CUDA" in str(exc)
    else:
        raise SystemExit("expected a loud failure without a CUDA device")
print("DROPIN-OK")
'''


def test_reference_flag_model_builds_on_our_modules():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_shim
    if not reference_shim.available():
        pytest.skip("/root/reference not mounted")
    out = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
