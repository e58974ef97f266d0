"""Device-side hierarchical connector (hgn_b200/rmp/hierarchical_connector.py: segment kernels + gathers) against the golden
vectors recorded from the live reference's ``HierarchicalConnector.run`` -- a training call that accumulates the normaliser
statistics followed by an evaluation call on another frame.  Indices bit-exact; features to fp32 summation order (the reference
averages per cluster with torch.mean on the CPU, the kernels sum a CSR segment in edge order)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "connector.npz"))
CASES = {"flag": ("flag", False, True), "flag_full": ("flag", True, False), "plate": ("plate", False, True), "few": ("flag", False, True)}


def _close(a, b, what, tol=2e-6):
    a, b = a.detach().float().cpu().numpy(), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max()))
    assert float(np.abs(a - b).max()) <= tol * scale, (what, float(np.abs(a - b).max()), scale)


class Recording:
    """Wraps one of our normalisers and keeps its un-normalised inputs (the golden file holds the reference's)."""

    def __init__(self, inner):
        self.inner, self.seen = inner, []

    def __call__(self, batched_data, accumulate=True):
        self.seen.append(batched_data.detach().clone())
        return self.inner(batched_data, accumulate)


@pytest.mark.parametrize("name", sorted(CASES))
def test_connector_matches_reference_golden(name):
    from hgn_b200.migration.normalizer import Normalizer
    from hgn_b200.rmp.hierarchical_connector import HierarchicalConnector
    from hgn_b200.util import EdgeSet, MultiGraphWithPos
    model_type, full, hyper_feats = CASES[name]
    dev = torch.device("cuda")
    members = torch.from_numpy(GOLD[f"{name}_cluster_members"])
    clusters = list(torch.split(members, GOLD[f"{name}_cluster_sizes"].tolist()))
    neighbors = [torch.from_numpy(x) for x in GOLD[f"{name}_neighbors"]]
    node_features = torch.from_numpy(GOLD[f"{name}_node_features"]).to(dev)
    mesh = torch.from_numpy(GOLD[f"{name}_mesh"]).to(dev)
    s = torch.from_numpy(GOLD[f"{name}_mesh_senders"]).to(dev)
    r = torch.from_numpy(GOLD[f"{name}_mesh_receivers"]).to(dev)
    f_edge = 7 if model_type == "flag" else 8
    conn = HierarchicalConnector(full, None, hyper_feats)
    intra, inter, hyper = Recording(Normalizer(f_edge, "intra")), Recording(Normalizer(f_edge, "inter")), Recording(Normalizer(3, "hyper"))
    names = conn.initialize(intra, inter, hyper)
    assert names == ['intra_cluster_to_mesh', 'intra_cluster_to_cluster', 'inter_cluster']
    for call, is_training in (("train", True), ("eval", False)):
        world = torch.from_numpy(GOLD[f"{name}_{call}_world"]).to(dev)
        sets = [EdgeSet("mesh_edges", torch.zeros(s.numel(), f_edge, device=dev), s, r)]
        graph = MultiGraphWithPos(node_features=node_features, edge_sets=sets, target_feature=world, mesh_features=mesh,
                                  model_type=model_type, node_dynamic=None, unnormalized_edges=None, obstacle_nodes=None)
        intra.seen, inter.seen, hyper.seen = [], [], []
        out = conn.run(graph, clusters, neighbors, is_training)
        # un-normalised quantities: tight (fp32 summation order only)
        _close(intra.seen[0], GOLD[f"{name}_{call}_intra_cluster_to_cluster_raw"], f"{call} raw to-cluster features")
        _close(intra.seen[1], GOLD[f"{name}_{call}_intra_cluster_to_mesh_raw"], f"{call} raw to-mesh features")
        _close(inter.seen[0], GOLD[f"{name}_{call}_inter_cluster_raw"], f"{call} raw inter features")
        if hyper_feats:
            _close(hyper.seen[0], GOLD[f"{name}_{call}_augmentation_raw"], f"{call} raw augmentation (size, mesh spread, world spread)")
        assert out.edge_sets is sets and [e.name for e in sets] == ["mesh_edges", "intra_cluster_to_cluster", "intra_cluster_to_mesh", "inter_cluster"]
        assert out.node_features[0] is node_features or torch.equal(out.node_features[0], node_features)
        # normalised outputs: the normaliser's variance |E[x^2] - E[x]^2| cancels in fp32 where a column barely varies across
        # clusters (equal block clusters: spreads equal to ~1e-3), which amplifies summation-order noise; hence the loose bound
        n_plain = GOLD[f"{name}_node_features"].shape[1]
        _close(out.node_features[1][:, :n_plain], GOLD[f"{name}_{call}_hyper_nodes"][:, :n_plain], f"{call} hyper node means")
        _close(out.node_features[1], GOLD[f"{name}_{call}_hyper_nodes"], f"{call} hyper nodes", tol=5e-2)
        for e in sets[1:]:
            assert np.array_equal(e.senders.cpu().numpy(), GOLD[f"{name}_{call}_{e.name}_senders"])
            assert np.array_equal(e.receivers.cpu().numpy(), GOLD[f"{name}_{call}_{e.name}_receivers"])
            _close(e.features, GOLD[f"{name}_{call}_{e.name}_features"], f"{call} {e.name}", tol=1e-4)


def test_connector_feeds_the_hyper_processor():
    # the connector's output is what the HyperGraphNet blocks consume: one forward of the drop-in model on it
    from hgn_b200.migration.meshgraphnet import MeshGraphNet
    from hgn_b200.migration.normalizer import Normalizer
    from hgn_b200.rmp.hierarchical_connector import HierarchicalConnector
    from hgn_b200.util import EdgeSet, MultiGraphWithPos
    name, dev = "flag", torch.device("cuda")
    members = torch.from_numpy(GOLD[f"{name}_cluster_members"])
    clusters = list(torch.split(members, GOLD[f"{name}_cluster_sizes"].tolist()))
    neighbors = [torch.from_numpy(x) for x in GOLD[f"{name}_neighbors"]]
    conn = HierarchicalConnector(False, None, True)
    edge_sets = ["mesh_edges"] + conn.initialize(Normalizer(7, "intra"), Normalizer(7, "inter"), Normalizer(3, "hyper"))
    s = torch.from_numpy(GOLD[f"{name}_mesh_senders"]).to(dev)
    r = torch.from_numpy(GOLD[f"{name}_mesh_receivers"]).to(dev)
    graph = MultiGraphWithPos(node_features=torch.from_numpy(GOLD[f"{name}_node_features"]).to(dev),
                              edge_sets=[EdgeSet("mesh_edges", torch.randn(s.numel(), 7, device=dev), s, r)],
                              target_feature=torch.from_numpy(GOLD[f"{name}_train_world"]).to(dev),
                              mesh_features=torch.from_numpy(GOLD[f"{name}_mesh"]).to(dev), model_type="flag", node_dynamic=None,
                              unnormalized_edges=None, obstacle_nodes=None)
    out = conn.run(graph, clusters, neighbors, True)
    torch.manual_seed(0)
    model = MeshGraphNet(3, 128, 2, "pna", 2, "hyper", edge_sets).to(dev)
    pred = model(out)
    assert pred.shape == (graph.node_features.shape[0], 3) and bool(torch.isfinite(pred).all())
