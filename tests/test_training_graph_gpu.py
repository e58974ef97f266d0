"""the whole training iteration -- encoder, processor, decoder, normalisers, masked loss, backward, Adam -- as
one CUDA graph (``hgn_b200.graphed.FlagTrainingGraph``) follows the reference's eager loop (the reference's own ``FlagModel`` on the
installed modules: build_graph / training_step / backward / optimizer.step per frame, MeshSimulator.py:131-139) over ten frames: the
losses, the weights after ten Adam steps and the normalisers' statistics.  Runs tests/training_graph_worker.py in a subprocess."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_graphed_training_iteration_follows_the_reference_loop(precision, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_shim
    if not reference_shim.available():
        pytest.skip("no reference tree: neither /root/reference nor oracle/_ref")
    out = str(tmp_path / "tg.npz")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "training_graph_worker.py"), precision, out], capture_output=True, text=True,
                         timeout=900, env=dict(os.environ, WANDB_MODE="disabled"))
    assert run.returncode == 0 and "TRAINING-GRAPH-OK" in run.stdout, run.stdout[-3000:] + run.stderr[-5000:]
    print("\n" + [ln for ln in run.stdout.splitlines() if "TRAINING-GRAPH-OK" in ln][-1][:600])
    z = np.load(out)
    eager, graphed, frozen = z["eager"], z["graphed"], z["frozen"]
    assert abs(graphed[0] - eager[0]) <= 2e-3 * abs(eager[0])                  # same weights, same frame, same statistics
    for k in range(1, len(eager)):
        moved = abs(eager[k] - frozen[k])                                       # what the optimizer steps changed on this frame
        assert abs(graphed[k] - eager[k]) <= max(0.1 * moved, 5e-3 * abs(eager[k])), (k, graphed[k], eager[k], frozen[k])
    assert float(z["weight_gap"]) <= 0.1 * float(z["weight_move"])
    assert float(z["normalizer_err"]) < 1e-4
