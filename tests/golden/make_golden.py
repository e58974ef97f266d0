"""Generate the committed golden fixtures from the LIVE reference (build container only).

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

Imports ``/root/reference`` through ``oracle/reference_shim.py`` (stubbed ``torch_scatter`` etc.),
builds graphs with the reference's own ``FlagModel.build_graph`` / ``expand_graph`` (spectral
clustering + ``HierarchicalConnector``), loads weights that are a pure function of the state_dict key
(``hgn_b200.synthetic.seeded_state_dict``) and records inputs, outputs, latents and gradient
projections.  The fixtures are small because weights are re-derived from their keys, not stored.
The reference cannot travel to the GPU box; these ``.npz`` files can.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))

import reference_shim  # noqa: E402
from hgn_b200 import synthetic  # noqa: E402

WEIGHT_SEED = 7
N_PROJ = 4


def seed_all():
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)


def flag_params(aggregation, steps, clustering="none", connector="none", num_clusters=4):
    return {
        "size": 3, "aggregation": aggregation, "message_passing_steps": steps,
        "rmp": {"num_clusters": num_clusters, "hyper_noise": 0.005, "hyper_node_features": True, "frequency": 1,
                "clustering": clustering, "connector": connector, "fully_connect": False,
                "intra_cluster_sampling": {"enabled": False, "alpha": 0.1, "spotter_threshold": 0},
                "hdbscan": {"max_cluster_size": 50, "min_cluster_size": 20, "min_samples": 1, "spotter_threshold": 0.9}},
        "graph_balancer": {"algorithm": "none", "frequency": 1, "remove_edges": True,
                           "ricci": {"loops": 150, "tau": 150}, "random": {"edge_amount": 100}},
    }


def grad_projections(named_grads):
    out = {}
    for key, g in named_grads.items():
        g64 = g.detach().double().reshape(-1)
        proj = [float(torch.dot(g64, synthetic.seeded_tensor(f"proj{i}:{key}", g64.shape, 11).double())) for i in range(N_PROJ)]
        out[key] = proj + [float(g64.norm())]
    return out


def record_graph(graph):
    rec = {}
    for i, nf in enumerate(graph.node_features):
        rec[f"node_features_{i}"] = nf.detach().cpu().numpy().astype(np.float32)
    names = []
    for es in graph.edge_sets:
        names.append(es.name)
        rec[f"edge_{es.name}_features"] = es.features.detach().cpu().numpy().astype(np.float32)
        rec[f"edge_{es.name}_senders"] = es.senders.detach().cpu().numpy().astype(np.int64)
        rec[f"edge_{es.name}_receivers"] = es.receivers.detach().cpu().numpy().astype(np.int64)
    return rec, names


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]     # optional: regenerate only the named cases


def run_case(name, src, graph, aggregation, steps, architecture, edge_sets, lean=False):
    if ONLY and name not in ONLY:
        return
    """graph: reference MultiGraph (raw features); runs reference MeshGraphNet with seeded weights."""
    from src.migration.meshgraphnet import MeshGraphNet
    from src.util import MultiGraph

    rec, set_names = record_graph(graph)

    def fresh_graph(requires_grad):
        nfs = [torch.from_numpy(rec[f"node_features_{i}"]).clone().requires_grad_(requires_grad)
               for i in range(len(graph.node_features))]
        sets = [es._replace(features=torch.from_numpy(rec[f"edge_{es.name}_features"]).clone().requires_grad_(requires_grad))
                for es in graph.edge_sets]
        return MultiGraph(nfs, sets)

    model = MeshGraphNet(output_size=3, latent_size=128, num_layers=2, message_passing_aggregator=aggregation,
                         message_passing_steps=steps, architecture=architecture, edge_sets=edge_sets)
    with torch.no_grad():
        model(fresh_graph(False))  # materialise the LazyLinear parameters
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synthetic.seeded_state_dict(shapes, WEIGHT_SEED))

    g = fresh_graph(True)
    latent_in = model.encoder(g)
    enc_nodes = [t.detach().clone() for t in latent_in.node_features]
    enc_edges = {es.name: es.features.detach().clone() for es in latent_in.edge_sets}
    latent_out = model.processor(latent_in)
    out = model.decoder(latent_out._replace(node_features=latent_out.node_features[0]))
    coef = synthetic.seeded_tensor("loss_coef", out.shape, 3)
    loss = (out * coef).sum()
    loss.backward()

    rec["output"] = out.detach().numpy()
    for i, t in enumerate(enc_nodes):
        rec[f"enc_node_{i}"] = t.numpy()
    for i, t in enumerate(latent_out.node_features):
        rec[f"proc_node_{i}"] = t.detach().numpy()
    proc_names = []
    for es in latent_out.edge_sets:
        proc_names.append(es.name)
        rec[f"proc_edge_{es.name}"] = es.features.detach().numpy()
    for nm, t in enc_edges.items():
        rec[f"enc_edge_{nm}"] = t.numpy()
    if lean:
        for key in ("proc_edge_mesh_edges", "enc_edge_mesh_edges"):
            rec.pop(key, None)
    for i, nf in enumerate(g.node_features):
        rec[f"grad_node_features_{i}"] = nf.grad.numpy()
    for es in g.edge_sets:
        if es.features.grad is not None:
            rec[f"grad_edge_{es.name}_features"] = es.features.grad.numpy()
    meta = {
        "name": name, "aggregation": aggregation, "steps": steps, "architecture": architecture,
        "edge_sets": list(edge_sets), "graph_edge_sets": set_names, "proc_edge_sets": proc_names,
        "weight_seed": WEIGHT_SEED, "shapes": {k: list(v) for k, v in shapes.items()},
        "loss": float(loss),
        "grad_proj": grad_projections({k: p.grad for k, p in model.named_parameters() if p.grad is not None}),
        "no_grad_params": [k for k, p in model.named_parameters() if p.grad is None],
        "n_node_lists": len(graph.node_features),
        "torch": torch.__version__,
    }
    rec["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: N={[tuple(x.shape) for x in graph.node_features]} sets={set_names} loss={float(loss):.6f} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def flag_graph(src, width, height, aggregation, clustering="none", connector="none", num_clusters=4):
    from src.model.flag import FlagModel
    seed_all()
    model = FlagModel(flag_params(aggregation, 1, clustering, connector, num_clusters))
    frame = synthetic.cloth_frame(width, height, seed=1)
    graph = model.build_graph(frame, is_training=True)   # accumulates + applies the normalisers
    graph = model.expand_graph(graph, 0, 1, is_training=False)  # no hyper-node noise
    return model, graph, frame


def segment_case(src):
    from src.util import unsorted_segment_operation
    rng = np.random.default_rng(5)
    E, S, D = 41, 9, 6
    data = rng.standard_normal((E, D)).astype(np.float32)
    data[7] = data[3]           # exact ties inside one segment -> single first winner
    ids = rng.integers(0, S - 2, size=E).astype(np.int64)   # segments S-2, S-1 stay empty
    ids[7] = ids[3]
    data1 = rng.standard_normal(E).astype(np.float32)
    rec = {"data": data, "ids": ids, "data1": data1, "num_segments": np.int64(S)}
    gup = rng.standard_normal((S, D)).astype(np.float32)
    rec["grad_up"] = gup
    for op in ("sum", "mean", "max", "min", "std"):
        x = torch.from_numpy(data).clone().requires_grad_(True)
        out = unsorted_segment_operation(x, torch.from_numpy(ids), S, op)
        rec[f"out_{op}"] = out.detach().numpy()
        if op != "std":
            (out * torch.from_numpy(gup)).sum().backward()
            rec[f"grad_{op}"] = x.grad.numpy()
        out1 = unsorted_segment_operation(torch.from_numpy(data1), torch.from_numpy(ids), S, op)
        rec[f"out1_{op}"] = out1.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "segment_ops.npz"), **rec)
    print("segment_ops done")


def edges_case(src):
    from src.util import triangles_to_edges
    rec = {}
    tri = torch.from_numpy(synthetic.grid_triangles(7, 5))
    perm = torch.from_numpy(np.random.default_rng(2).permutation(tri.shape[0]))
    tri = tri[perm]
    d = triangles_to_edges(tri)
    rec["tri_cells"] = tri.numpy()
    rec["tri_senders"], rec["tri_receivers"] = (t.numpy() for t in d["two_way_connectivity"])
    tet = torch.from_numpy(synthetic.box_tetrahedra(3, 3, 2))
    d = triangles_to_edges(tet, deform=True)
    rec["tet_cells"] = tet.numpy()
    rec["tet_senders"], rec["tet_receivers"] = (t.numpy() for t in d["two_way_connectivity"])
    np.savez_compressed(os.path.join(HERE, "mesh_edges.npz"), **rec)
    print("mesh_edges done")


def main():
    src = reference_shim.load()
    os.chdir(reference_shim.REFERENCE_ROOT)
    if not ONLY:
        segment_case(src)
        edges_case(src)

    # MeshGraphNets (no remote path): graph from the reference's FlagModel.build_graph
    fm, graph, _ = flag_graph(src, 6, 5, "sum")
    run_case("mgn_sum_L2", src, graph, "sum", 2, "none", ["mesh_edges"])
    run_case("mgn_pna_L1", src, graph, "pna", 1, "none", ["mesh_edges"])
    run_case("mgn_max_L1", src, graph, "max", 1, "none", ["mesh_edges"])
    run_case("repeated_sum_L1", src, graph, "sum", 1, "repeated", ["mesh_edges"])
    run_case("multi_mean_L1", src, graph, "mean", 1, "multi", ["mesh_edges"])

    # HyperGraphNets: spectral clustering + HierarchicalConnector from the reference (CPU)
    hyper_sets = ["mesh_edges", "intra_cluster_to_mesh", "intra_cluster_to_cluster", "inter_cluster"]
    fm, graph, _ = flag_graph(src, 8, 6, "pna", "spectral", "hyper", 4)
    run_case("hgn_hyper_pna_L1", src, graph, "pna", 1, "hyper", hyper_sets)
    run_case("hgn_hyper_sum_L2", src, graph, "sum", 2, "hyper", hyper_sets)
    run_case("hgn_hetero_pna_L1", src, graph, "pna", 1, "hetero", hyper_sets)
    run_case("hgn_multiscale_sum_L1", src, graph, "sum", 1, "multiscale", hyper_sets)

    # the same three remote architectures above one 128-row tile (round 2): 20 x 15 cloth = 300 nodes / 1 658 mesh edges (13 tiles),
    # 16 clusters; the mesh-edge latents are left out of the fixture (LEAN) to keep it small -- node latents, outputs, the other
    # edge sets, input gradients and every parameter-gradient projection are kept
    fm, graph, _ = flag_graph(src, 20, 15, "pna", "spectral", "hyper", 16)
    run_case("hgn_hyper_pna_L2_300", src, graph, "pna", 2, "hyper", hyper_sets, lean=True)
    run_case("hgn_hetero_pna_L2_300", src, graph, "pna", 2, "hetero", hyper_sets, lean=True)
    run_case("hgn_multiscale_pna_L1_300", src, graph, "pna", 1, "multiscale", hyper_sets, lean=True)


if __name__ == "__main__":
    main()
