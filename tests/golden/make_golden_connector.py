"""Golden vectors for the hierarchical connector from the LIVE reference (build container only).

    PYTHONHASHSEED=0 python tests/golden/make_golden_connector.py

Calls the reference's own ``HierarchicalConnector.run`` (src/rmp/hierarchical_connector.py:27-143) with fresh reference
``Normalizer``s on synthetic flag- and plate-typed graphs whose clustering is a block partition of the lattice, twice per case
(a training call that accumulates the normaliser statistics, then an evaluation call that only applies them), and records the
inputs and everything the connector returns -> ``tests/golden/connector.npz``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))

import reference_shim  # noqa: E402
from hgn_b200 import synthetic  # noqa: E402


def block_clusters(width, height, bw, bh):
    """Clusters = bw x bh blocks of a width x height lattice (node id = j * width + i); neighbours = edge-adjacent block pairs."""
    ids = np.arange(width * height).reshape(height, width)
    nbx, nby = -(-width // bw), -(-height // bh)
    clusters, neighbors = [], []
    for by in range(nby):
        for bx in range(nbx):
            clusters.append(torch.tensor(ids[by * bh:(by + 1) * bh, bx * bw:(bx + 1) * bw].reshape(-1).tolist()))
            k = by * nbx + bx
            if bx + 1 < nbx:
                neighbors.append(torch.tensor([k, k + 1]))
            if by + 1 < nby:
                neighbors.append(torch.tensor([k, k + nbx]))
    return clusters, neighbors


CASES = {
    # name: (model_type, lattice w, h, block w, h, fully_connect, hyper_node_features)
    "flag": ("flag", 12, 10, 4, 5, False, True),
    "flag_full": ("flag", 9, 6, 3, 3, True, False),
    "plate": ("plate", 10, 8, 5, 3, False, True),
    "few": ("flag", 6, 6, 3, 6, False, True),            # 2 clusters (< 4): fully connected whatever the flag says
}


def case_inputs(model_type, w, h, seed):
    g = torch.Generator().manual_seed(seed)
    frame = synthetic.cloth_frame(w, h, seed=seed)
    n = w * h
    world = frame["world_pos"].float()
    mesh = frame["mesh_pos"].float()
    if model_type == "plate":
        mesh = torch.cat([mesh, 0.05 * torch.rand(n, 1, generator=g)], dim=1)      # 3-D rest positions
    node_features = torch.randn(n, 5 if model_type == "flag" else 6, generator=g)
    return node_features, world, mesh, frame


def main():
    reference_shim.load()
    os.chdir(reference_shim.REFERENCE_ROOT)
    from src.migration.normalizer import Normalizer
    from src.rmp.hierarchical_connector import HierarchicalConnector
    from src.util import EdgeSet, MultiGraphWithPos, triangles_to_edges
    rec = {}
    for seed, (name, (model_type, w, h, bw, bh, full, hyper_feats)) in enumerate(CASES.items()):
        node_features, world, mesh, frame = case_inputs(model_type, w, h, seed + 20)
        clusters, neighbors = block_clusters(w, h, bw, bh)
        s, r = triangles_to_edges(frame["cells"].long())["two_way_connectivity"]
        f_edge = 7 if model_type == "flag" else 8
        class Recording(Normalizer):
            """The reference normaliser, remembering its un-normalised inputs: |E[x^2] - E[x]^2| cancels in fp32 when a feature
            barely varies across clusters, so normalised columns are compared loosely and the raw ones tightly."""
            def forward(self, batched_data, accumulate=True):
                self.seen = getattr(self, "seen", []) + [batched_data.detach().clone()]
                return super().forward(batched_data, accumulate)

        conn = HierarchicalConnector(full, None, hyper_feats)
        norms = {"intra": Recording(f_edge, "intra"), "inter": Recording(f_edge, "inter"), "hyper": Recording(3, "hyper")}
        conn.initialize(norms["intra"], norms["inter"], norms["hyper"])
        rec[f"{name}_node_features"] = node_features.numpy()
        rec[f"{name}_world"] = world.numpy()
        rec[f"{name}_mesh"] = mesh.numpy()
        rec[f"{name}_mesh_senders"], rec[f"{name}_mesh_receivers"] = s.numpy(), r.numpy()
        rec[f"{name}_cluster_members"] = torch.cat(clusters).numpy()
        rec[f"{name}_cluster_sizes"] = np.asarray([len(c) for c in clusters], np.int64)
        rec[f"{name}_neighbors"] = torch.stack(neighbors).numpy()
        for call, is_training in (("train", True), ("eval", False)):
            wpos = world + (0.01 if call == "eval" else 0.0)             # a different frame for the second call
            rec[f"{name}_{call}_world"] = wpos.numpy()
            graph = MultiGraphWithPos(node_features=node_features, edge_sets=[EdgeSet("mesh_edges", torch.zeros(s.numel(), f_edge), s, r)],
                                      target_feature=wpos, mesh_features=mesh, model_type=model_type, node_dynamic=None,
                                      unnormalized_edges=None, obstacle_nodes=None)
            for nz in norms.values():
                nz.seen = []
            out = conn.run(graph, clusters, neighbors, is_training)
            raw = {"intra_cluster_to_cluster": norms["intra"].seen[0], "intra_cluster_to_mesh": norms["intra"].seen[1],
                   "inter_cluster": norms["inter"].seen[0]}
            for k, v in raw.items():
                rec[f"{name}_{call}_{k}_raw"] = v.cpu().numpy().astype(np.float32)
            if hyper_feats:
                rec[f"{name}_{call}_augmentation_raw"] = norms["hyper"].seen[0].cpu().numpy().astype(np.float32)
            assert [e.name for e in out.edge_sets] == ["mesh_edges", "intra_cluster_to_cluster", "intra_cluster_to_mesh", "inter_cluster"]
            rec[f"{name}_{call}_hyper_nodes"] = out.node_features[1].detach().cpu().numpy().astype(np.float32)
            for e in out.edge_sets[1:]:
                rec[f"{name}_{call}_{e.name}_features"] = e.features.detach().cpu().numpy().astype(np.float32)
                rec[f"{name}_{call}_{e.name}_senders"] = e.senders.cpu().numpy().astype(np.int64)
                rec[f"{name}_{call}_{e.name}_receivers"] = e.receivers.cpu().numpy().astype(np.int64)
        print(name, "clusters", len(clusters), "hyper", rec[f"{name}_eval_hyper_nodes"].shape,
              {e.name: int(e.senders.numel()) for e in out.edge_sets[1:]})
    np.savez_compressed(os.path.join(HERE, "connector.npz"), **rec)
    print(os.path.getsize(os.path.join(HERE, "connector.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
