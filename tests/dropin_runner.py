"""TEST INFRASTRUCTURE -- one arm of the drop-in parity tests (tests/test_dropin_gpu.py), run as a subprocess because the two
arms need different ``sys.modules`` (and the reference captures ``src.util.device`` at import time).

    python tests/dropin_runner.py ARM CASE OUT.npz [precision [REFERENCE.npz]]

ARM ``reference``: the UNMODIFIED reference (``/root/reference`` when mounted, else the staged ``oracle/_ref``; torch_scatter & co
stubbed by ``oracle/reference_shim.py``) on the CPU -- the caller hides the GPUs so that ``src.util.device`` is ``cpu``.
ARM ``ours``: ``hgn_b200.install_as_reference_modules()`` FIRST (the documented order), then the reference's own
system model classes on CUDA: ``src.migration.*`` / ``Normalizer`` / ``util.unsorted_segment_operation`` are hgn_b200's.

CASE ``flag`` (FlagModel, GraphNet pna, 15 layers, 40x40 cloth -- BASELINE.json configs[1]), ``plate`` (PlateModel with the
plateCluster.yaml model section: spectral clustering, 31 clusters, HeteroGraphNet, 5 edge sets incl. world edges; batches of 2
through the reference's own ``MeshSimulator._get_batched`` -- configs[2]), ``cylinder`` (CylinderModel, GraphNet pna, 5 layers,
batches of 2 -- configs[3]), ``flag_hyper`` (FlagModel with flag.yaml's spectral + hyper connector, HyperGraphNet).

With REFERENCE.npz (the reference arm's output) the ``ours`` arm first checks that the graphs it built itself have the reference's
index lists bit for bit and its features to fp32 rounding, and then runs the training steps on the REFERENCE's feature tensors, so
that the loss / gradient comparison sees identical inputs (the reference's own torch code gives slightly different features on
the two devices -- its normalisers' E[x^2] - E[x]^2 amplifies the last bit); the rollout always builds its own graphs.

Both arms execute the same script below: accumulate the normalisers on three frames, load weights that are a pure function of the
state_dict key, two ``training_step`` + ``loss.backward()`` + ``Adam.step()`` iterations (MeshSimulator.py:131-139), then
``model.rollout(trajectory, K)`` (flag.py:194-246 / plate.py:264-312 / cylinder.py:175-209).  Everything the test compares goes
into OUT.npz.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))

WEIGHT_SEED = 13
ROLLOUT_STEPS = {"flag": 20, "flag_hyper": 6, "plate": 6, "cylinder": 8}


def model_params(case):
    rmp = {"num_clusters": 31, "hyper_noise": "none", "hyper_node_features": True, "frequency": 1, "clustering": "none",
           "connector": "none", "fully_connect": False,
           "intra_cluster_sampling": {"enabled": False, "alpha": 0.1, "spotter_threshold": 0},
           "hdbscan": {"max_cluster_size": 50, "min_cluster_size": 20, "min_samples": 1, "spotter_threshold": 0.9}}
    params = {"size": 3, "aggregation": "pna", "message_passing_steps": 5, "rmp": rmp,
              "graph_balancer": {"algorithm": "none", "frequency": 1, "remove_edges": True,
                                 "ricci": {"loops": 150, "tau": 150}, "random": {"edge_amount": 100}}}
    if case == "flag":
        params["message_passing_steps"] = 15                      # BASELINE.json configs[1]: 15 MP layers
    elif case == "flag_hyper":                                    # configs/flag.yaml:38-41
        rmp.update(clustering="spectral", connector="hyper", num_clusters=10)
    elif case == "plate":                                         # configs/plateCluster.yaml:36-43
        rmp.update(clustering="spectral", connector="hetero", num_clusters=31)
    return params


def make_frames(case, count):
    """`count` consecutive frames of one synthetic trajectory (SURVEY.md s8d shapes), CPU tensors."""
    from hgn_b200 import synthetic
    frames = []
    if case in ("flag", "flag_hyper"):
        w, h = (40, 40) if case == "flag" else (20, 15)
        base = synthetic.cloth_frame(w, h, seed=1)
        n = w * h
        pos = [base["world_pos"] + 0.002 * t * synthetic.seeded_tensor("drift", (n, 3), 2) + 0.0005 * synthetic.seeded_tensor(f"jit{t}", (n, 3), 2)
               for t in range(-1, count + 1)]
        for t in range(count):
            frames.append({**base, "prev|world_pos": pos[t], "world_pos": pos[t + 1], "target|world_pos": pos[t + 2]})
    elif case == "plate":
        base = synthetic.plate_frame(plate=(13, 13, 3), obstacle=(5, 5, 2), seed=3)
        n = base["world_pos"].shape[0]
        obstacle = (base["node_type"][:, 0] == 1).float().unsqueeze(1)
        push = torch.tensor([0.0, 0.0, -0.0006])                 # the obstacle sinks into the plate: the contact set changes
        pos = [base["world_pos"] + t * push * obstacle + 0.0002 * t * synthetic.seeded_tensor("drift", (n, 3), 4) * (1 - obstacle)
               for t in range(count + 1)]
        for t in range(count):
            frames.append({**base, "world_pos": pos[t], "target|world_pos": pos[t + 1]})
    elif case == "cylinder":
        base = synthetic.cylinder_frame(32, 24, seed=5)
        n = base["velocity"].shape[0]
        vel = [base["velocity"] + 0.01 * t * synthetic.seeded_tensor("dv", (n, 2), 6) for t in range(count + 1)]
        for t in range(count):
            frames.append({**base, "velocity": vel[t], "target|velocity": vel[t + 1],
                           "pressure": base["pressure"] + 0.01 * t * synthetic.seeded_tensor("dp", (n, 1), 6)})
    else:
        raise SystemExit(f"unknown case {case}")
    return frames


def main():
    arm, case, out_path = sys.argv[1:4]
    precision = sys.argv[4] if len(sys.argv) > 4 else "fp32"
    given = dict(np.load(sys.argv[5])) if len(sys.argv) > 5 else None
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    import reference_shim
    if not reference_shim.available():
        raise SystemExit("SKIP: no reference tree (neither /root/reference nor oracle/_ref)")
    reference_shim._install_stub_modules()
    sys.path.insert(0, reference_shim.REFERENCE_ROOT)
    os.chdir(reference_shim.REFERENCE_ROOT)
    if arm == "ours":
        assert torch.cuda.is_available(), "the 'ours' arm needs the GPU"
        import hgn_b200
        hgn_b200.set_precision(precision)
        hgn_b200.install_as_reference_modules()
        import src.migration.graphnet as installed
        assert installed.__name__.startswith("hgn_b200."), installed.__name__
    else:
        assert not torch.cuda.is_available(), "run the reference arm with the GPUs hidden (src.util.device must be cpu)"
    import src.util
    from src.algorithms.MeshSimulator import MeshSimulator
    from src.model.cylinder import CylinderModel
    from src.model.flag import FlagModel
    from src.model.plate import PlateModel
    from hgn_b200 import synthetic
    if arm == "ours":
        import hgn_b200.util
        assert src.util.unsorted_segment_operation is hgn_b200.util.unsorted_segment_operation
        assert hasattr(src.util, "read_yaml") and hasattr(src.util, "detach")
    device = src.util.device

    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    cls = {"flag": FlagModel, "flag_hyper": FlagModel, "plate": PlateModel, "cylinder": CylinderModel}[case]
    model = cls(model_params(case))
    if getattr(model, "_rmp", False):
        model._remote_graph._clustering_algorithm.visualize_cluster = lambda *a, **k: None   # wandb.Object3D upload only
    net = model.learned_model
    if arm == "ours":
        import hgn_b200.migration.meshgraphnet as ours_mgn
        assert type(net) is ours_mgn.MeshGraphNet, type(net)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4)     # MeshSimulator.py:109-110: before the first forward
    batch = 1 if case in ("flag", "flag_hyper") else 2
    warm = 3
    steps = 2
    total = warm + steps * batch
    frames = [{k: v.to(device) for k, v in f.items()} for f in make_frames(case, total + ROLLOUT_STEPS[case] + 1)]

    rec = {}
    model.train()
    graphs = []
    for t in range(total):                                        # fetch_data (MeshSimulator.py:252-259): normalisers accumulate
        g = model.build_graph(frames[t], True)
        g = model.expand_graph(g, t, total, True)
        graphs.append(g)
    with torch.no_grad():                                         # materialise the lazy linears, then seeded weights
        model(graphs[0])
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict({k: v.to(device) for k, v in synthetic.seeded_state_dict(shapes, WEIGHT_SEED).items()})
    g0 = graphs[0]
    rec["n_nodes"] = np.asarray([int(x.shape[0]) for x in g0.node_features])
    for es in g0.edge_sets:                                       # graph construction is compared bit for bit
        rec[f"graph_{es.name}_senders"] = es.senders.detach().cpu().numpy().astype(np.int64)
        rec[f"graph_{es.name}_receivers"] = es.receivers.detach().cpu().numpy().astype(np.int64)
        rec[f"graph_{es.name}_features"] = es.features.detach().cpu().numpy()
    for i, nf in enumerate(g0.node_features):
        rec[f"graph_node_features_{i}"] = nf.detach().cpu().numpy()

    for i in range(warm, total):                                  # the graphs the training steps consume
        g = graphs[i]
        for j, nf in enumerate(g.node_features):
            rec[f"tg{i}_nf{j}"] = nf.detach().cpu().numpy()
        for es in g.edge_sets:
            rec[f"tg{i}_ef_{es.name}"] = es.features.detach().cpu().numpy()
            rec[f"tg{i}_es_{es.name}"] = np.stack([es.senders.detach().cpu().numpy(), es.receivers.detach().cpu().numpy()]).astype(np.int64)
    if given is not None:
        worst = 0.0
        for i in range(warm, total):
            g = graphs[i]
            nfs = []
            for j, nf in enumerate(g.node_features):
                want = given[f"tg{i}_nf{j}"]
                assert want.shape == tuple(nf.shape), (i, j, want.shape, tuple(nf.shape))
                worst = max(worst, float(np.abs(nf.detach().cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-30)))
                nfs.append(torch.from_numpy(want).to(device))
            sets = []
            for es in g.edge_sets:
                want = given[f"tg{i}_ef_{es.name}"]
                assert np.array_equal(rec[f"tg{i}_es_{es.name}"], given[f"tg{i}_es_{es.name}"]), f"graph {i}: {es.name} index lists differ from the reference"
                assert want.shape == tuple(es.features.shape)
                if want.size:
                    worst = max(worst, float(np.abs(es.features.detach().cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-30)))
                sets.append(es._replace(features=torch.from_numpy(want).to(device)))
            graphs[i] = g._replace(node_features=nfs, edge_sets=sets)
        rec["own_graph_feature_error"] = np.asarray(worst)

    data = list(zip(graphs[warm:], frames[warm:total]))
    batches = MeshSimulator._get_batched(data, batch) if batch > 1 else data
    assert len(batches) == steps
    losses = []
    for it, (graph, frame) in enumerate(batches):
        loss = model.training_step(graph, frame)
        loss.backward()
        if it == 0:
            names, norms, projs = [], [], []
            for key, p in net.named_parameters():
                if p.grad is None:
                    continue
                g64 = p.grad.detach().double().cpu().reshape(-1)
                names.append(key)
                norms.append(float(g64.norm()))
                projs.append([float(torch.dot(g64, synthetic.seeded_tensor(f"proj{i}:{key}", g64.shape, 11).double())) for i in range(3)])
            rec["grad_names"] = np.frombuffer(json.dumps(names).encode(), dtype=np.uint8)
            rec["grad_norms"] = np.asarray(norms)
            rec["grad_projs"] = np.asarray(projs)
            rec["edge_set_sizes"] = np.frombuffer(json.dumps({es.name: int(es.senders.shape[0]) for es in graph.edge_sets}).encode(), dtype=np.uint8)
        optimizer.step()
        optimizer.zero_grad()
        losses.append(float(loss.detach()))
    rec["losses"] = np.asarray(losses)

    model.eval()
    k = ROLLOUT_STEPS[case]
    traj = synthetic.trajectory(frames[total:total + k])
    if getattr(model, "_rmp", False):
        model._remote_graph.reset_clusters()
    traj_ops, mse = model.rollout(traj, k)
    pred_key = "pred_velocity" if case == "cylinder" else "pred_pos"
    rec["rollout_pred"] = traj_ops[pred_key].detach().cpu().numpy()
    rec["rollout_mse"] = mse.detach().cpu().numpy()
    rec["rollout_gt"] = (traj["velocity"] if case == "cylinder" else traj["world_pos"]).cpu().numpy()
    np.savez_compressed(out_path, **rec)
    print(f"DROPIN-RUNNER-OK {arm} {case} {precision} losses={losses} rollout_mse_last={float(mse[-1]):.4e}")


if __name__ == "__main__":
    main()
