"""TEST INFRASTRUCTURE -- worker of tests/test_training_graph_gpu.py (a subprocess: it installs hgn_b200 as the reference's
``src.migration``).  The reference's own training loop for a FlagModel (MeshSimulator.py:131-139: build_graph, training_step,
backward, Adam step, per frame) eagerly, against the same loop captured in one CUDA graph per iteration
(``hgn_b200.graphed.FlagTrainingGraph``: encoder + processor + decoder + normalisers + masked loss + backward + Adam)."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
sys.path.insert(0, HERE)
os.environ.setdefault("WANDB_MODE", "disabled")

import reference_shim  # noqa: E402
import dropin_runner  # noqa: E402

if not reference_shim.available():
    raise SystemExit("SKIP: no reference tree")
reference_shim._install_stub_modules()
sys.path.insert(0, reference_shim.REFERENCE_ROOT)
os.chdir(reference_shim.REFERENCE_ROOT)
import hgn_b200  # noqa: E402
from hgn_b200 import synthetic  # noqa: E402
from hgn_b200.graphed import FlagTrainingGraph  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
hgn_b200.set_precision(precision)
hgn_b200.install_as_reference_modules()
from src.model.flag import FlagModel  # noqa: E402
import src.util  # noqa: E402

dev = src.util.device
assert dev.type == "cuda"
params = dropin_runner.model_params("flag")
params["message_passing_steps"] = 5
frames = [{k: v.to(dev) for k, v in f.items()} for f in dropin_runner.make_frames("flag_hyper", 10)]     # 20 x 15 cloth
torch.manual_seed(0)
model = FlagModel(params)
model.train()
with torch.no_grad():
    model(model.build_graph(frames[0], False))                   # materialise the lazy linears
net = model.learned_model
shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
net.load_state_dict({k: v.to(dev) for k, v in synthetic.seeded_state_dict(shapes, 3).items()})
twin = copy.deepcopy(model)                                      # same weights, same (empty) normaliser statistics
lr = 1e-3

# eager: the reference's loop
opt = torch.optim.Adam(model.parameters(), lr=lr)
eager = []
for f in frames:
    graph = model.build_graph(f, True)
    loss = model.training_step(graph, f)
    loss.backward()
    opt.step()
    opt.zero_grad()
    eager.append(float(loss))
# frozen weights (lr = 0): what stale packed weights inside the graph would reproduce
frozen_model = copy.deepcopy(twin)
frozen = []
for f in frames:
    with torch.no_grad():
        frozen.append(float(frozen_model.training_step(frozen_model.build_graph(f, True), f)))
# graphed
opt2 = torch.optim.Adam(twin.parameters(), lr=lr, capturable=True)
runner = FlagTrainingGraph(twin, frames[0], opt2)
graphed = [float(runner.step(f)) for f in frames]
w_eager = torch.cat([p.detach().reshape(-1) for p in model.learned_model.parameters()])
w_graph = torch.cat([p.detach().reshape(-1) for p in twin.learned_model.parameters()])
w_start = torch.cat([p.detach().reshape(-1) for p in frozen_model.learned_model.parameters()])
norm_err = max(float((getattr(a, f) - getattr(b, f)).abs().max() / getattr(a, f).abs().max().clamp_min(1e-30))
               for a, b in ((model._node_normalizer, twin._node_normalizer), (model._mesh_edge_normalizer, twin._mesh_edge_normalizer),
                            (model._output_normalizer, twin._output_normalizer)) for f in ("_acc_sum", "_acc_sum_squared", "_acc_count"))
np.savez(sys.argv[2], eager=np.asarray(eager), graphed=np.asarray(graphed), frozen=np.asarray(frozen),
         weight_gap=float((w_graph - w_eager).norm()), weight_move=float((w_eager - w_start).norm()), normalizer_err=norm_err)
print("TRAINING-GRAPH-OK", precision, "eager", [f"{x:.5f}" for x in eager[:4]], "graphed", [f"{x:.5f}" for x in graphed[:4]], "frozen", [f"{x:.5f}" for x in frozen[:4]],
      f"weights: gap {float((w_graph - w_eager).norm()):.3e} of a move of {float((w_eager - w_start).norm()):.3e}; normaliser statistics {norm_err:.1e}")
