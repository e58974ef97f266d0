"""Golden vectors for the world-edge search from the LIVE reference (build container only).

    PYTHONHASHSEED=0 python tests/golden/make_golden_world_edges.py

Runs the reference's own ``PlateModel.build_graph`` (src/model/plate.py:69-190: torch.cdist matrix, radius 0.03, mesh edges /
non-obstacle rows / non-normal columns cleared, torch.nonzero) on synthetic two-body plate frames and records the inputs next
to the ``mesh_edges`` and ``world_edges`` index lists it built -> ``tests/golden/world_edges.npz``.
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
sys.path.insert(0, HERE)

import reference_shim  # noqa: E402
from hgn_b200 import synthetic  # noqa: E402
from make_golden import flag_params, seed_all  # noqa: E402

CASES = {"a": dict(plate=(9, 9, 2), obstacle=(4, 4, 2), gap=0.012, seed=0),
         "b": dict(plate=(12, 7, 3), obstacle=(5, 3, 2), gap=0.02, seed=4),
         "c": dict(plate=(6, 6, 2), obstacle=(3, 3, 2), gap=0.2, seed=2)}      # c: bodies apart, no world edges


def main():
    reference_shim.load()
    os.chdir(reference_shim.REFERENCE_ROOT)
    from src.model.plate import PlateModel
    rec = {}
    for name, kw in CASES.items():
        seed_all()
        model = PlateModel(flag_params("sum", 1))
        frame = synthetic.plate_frame(**kw)
        graph = model.build_graph(frame, is_training=True)
        sets = {es.name: es for es in graph.edge_sets}
        for key in ("world_pos", "node_type", "cells"):
            rec[f"{name}_{key}"] = frame[key].numpy()
        for es_name in ("mesh_edges", "world_edges"):
            rec[f"{name}_{es_name}_senders"] = sets[es_name].senders.cpu().numpy().astype(np.int64)
            rec[f"{name}_{es_name}_receivers"] = sets[es_name].receivers.cpu().numpy().astype(np.int64)
        print(name, "nodes", frame["world_pos"].shape[0], "mesh edges", sets["mesh_edges"].senders.numel(),
              "world edges", sets["world_edges"].senders.numel())
    rec["torch_version"] = np.frombuffer(torch.__version__.encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "world_edges.npz"), **rec)


if __name__ == "__main__":
    main()
