"""Host-side graph building either side of the processor: the connector's index lists against the golden vectors of the live
reference, and the in-trajectory batching against a literal restatement of the reference's list comprehension (and the live
reference's ``MeshSimulator._get_batched`` when it is mounted)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
from hgn_b200.batching import get_batched  # noqa: E402
from hgn_b200.rmp.hierarchical_connector import connector_indices  # noqa: E402
from hgn_b200.util import EdgeSet, MultiGraph  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "connector.npz"))
CASES = {"flag": False, "flag_full": True, "plate": False, "few": False}          # name -> fully_connect


def golden_clusters(name):
    members = torch.from_numpy(GOLD[f"{name}_cluster_members"])
    sizes = GOLD[f"{name}_cluster_sizes"].tolist()
    clusters = list(torch.split(members, sizes))
    neighbors = [torch.from_numpy(x) for x in GOLD[f"{name}_neighbors"]]
    return clusters, neighbors


@pytest.mark.parametrize("name", sorted(CASES))
def test_connector_indices_are_bit_exact(name):
    clusters, neighbors = golden_clusters(name)
    n = GOLD[f"{name}_node_features"].shape[0]
    idx = connector_indices(clusters, neighbors, n, CASES[name])
    for set_name, (s, r) in idx.items():
        assert s.dtype == torch.int64 and r.dtype == torch.int64
        for call in ("train", "eval"):
            assert np.array_equal(s.numpy(), GOLD[f"{name}_{call}_{set_name}_senders"]), set_name
            assert np.array_equal(r.numpy(), GOLD[f"{name}_{call}_{set_name}_receivers"]), set_name


def _reference_remap(xs, i, batch_size, num_nodes, num_hyper_nodes):
    # src/algorithms/MeshSimulator.py:205-217, verbatim semantics
    hyper_node_offset = batch_size * num_nodes
    return torch.tensor([x + i * num_nodes if x < hyper_node_offset else x + (batch_size - 1) * num_nodes + i * num_hyper_nodes
                         for x in xs.tolist()])


def _trajectory(n, c, steps, seed):
    g = torch.Generator().manual_seed(seed)
    data = []
    for t in range(steps):
        nodes = [torch.randn(n, 4, generator=g)] + ([torch.randn(c, 6, generator=g)] if c else [])
        sets = [EdgeSet("mesh_edges", torch.randn(30, 3, generator=g), torch.randint(0, n, (30,), generator=g), torch.randint(0, n, (30,), generator=g))]
        if c:
            sets.append(EdgeSet("intra_cluster_to_cluster", torch.randn(n, 3, generator=g), torch.arange(n), n + torch.randint(0, c, (n,), generator=g)))
            sets.append(EdgeSet("inter_cluster", torch.randn(5, 3, generator=g), n + torch.randint(0, c, (5,), generator=g), n + torch.randint(0, c, (5,), generator=g)))
        data.append((MultiGraph(nodes, sets), {"world_pos": torch.randn(n, 3, generator=g), "step": torch.tensor([t])}))
    return data


@pytest.mark.parametrize("n,c,steps,batch", [(12, 0, 6, 3), (12, 3, 6, 2), (9, 2, 5, 2), (7, 4, 4, 1), (5, 2, 3, 8)])
def test_get_batched_matches_the_reference_remap(n, c, steps, batch):
    data = _trajectory(n, c, steps, seed=n + c)
    out = get_batched(data, batch)
    assert len(out) == -(-steps // batch)
    for b, (graph, traj) in enumerate(out):
        chunk = data[b * batch:(b + 1) * batch]
        assert [x.shape[0] for x in graph.node_features] == [n * len(chunk)] + ([c * len(chunk)] if c else [])
        assert torch.equal(traj["world_pos"], torch.cat([t["world_pos"] for _, t in chunk]))
        for k, es in enumerate(graph.edge_sets):
            assert es.name == chunk[0][0].edge_sets[k].name and es.senders.dtype == torch.int64
            assert torch.equal(es.features, torch.cat([g.edge_sets[k].features for g, _ in chunk]))
            want_s = torch.cat([_reference_remap(g.edge_sets[k].senders, i, batch, n, c) for i, (g, _) in enumerate(chunk)])
            want_r = torch.cat([_reference_remap(g.edge_sets[k].receivers, i, batch, n, c) for i, (g, _) in enumerate(chunk)])
            assert torch.equal(es.senders, want_s) and torch.equal(es.receivers, want_r)
    if c and batch > 1 and steps >= 2:
        # the reference's quirk: graph 1's hyper indices are shifted by num_nodes, not into the hyper block
        es = out[0][0].edge_sets[1]
        second = es.receivers[n:2 * n]
        assert bool((second < batch * n + c).all()) and bool((second >= n).all())


def test_get_batched_matches_live_reference_when_mounted():
    import reference_shim
    if not reference_shim.available():
        pytest.skip("/root/reference not mounted")
    reference_shim.load()
    from src.algorithms.MeshSimulator import MeshSimulator
    for n, c, steps, batch in [(12, 3, 6, 2), (10, 0, 4, 4)]:
        data = _trajectory(n, c, steps, seed=5)
        ours, ref = get_batched(data, batch), MeshSimulator._get_batched(data, batch)
        assert len(ours) == len(ref)
        for (g0, t0), (g1, t1) in zip(ours, ref):
            assert all(torch.equal(a, b) for a, b in zip(g0.node_features, g1.node_features))
            assert all(torch.equal(t0[k], t1[k]) for k in t1)
            for e0, e1 in zip(g0.edge_sets, g1.edge_sets):
                assert e0.name == e1.name and torch.equal(e0.senders, e1.senders) and torch.equal(e0.receivers, e1.receivers)
                assert torch.equal(e0.features, e1.features)


def test_triangles_to_edges_topology_cache():
    from hgn_b200 import synthetic, util
    import hgn_oracle as orc
    cells = torch.from_numpy(synthetic.grid_triangles(7, 5)).long()
    first = util.triangles_to_edges(cells)
    again = util.triangles_to_edges(cells)
    ref = orc.triangles_to_edges(cells)
    for a, b, c in zip(first["two_way_connectivity"], again["two_way_connectivity"], ref["two_way_connectivity"]):
        assert a is b and torch.equal(a, c)                      # same object from the cache, equal to the oracle
    cells[0] = cells[0].flip(0)                                  # in-place change: the version check must miss
    changed = util.triangles_to_edges(cells)
    assert changed["senders"] is not first["senders"] and torch.equal(changed["senders"], orc.triangles_to_edges(cells)["senders"])
    clone = cells.clone()                                        # another object with equal content: computed afresh, equal result
    assert torch.equal(util.triangles_to_edges(clone)["receivers"], changed["receivers"])
    tets = torch.from_numpy(synthetic.box_tetrahedra(3, 3, 2)).long()
    d1, d2 = util.triangles_to_edges(tets, deform=True), orc.triangles_to_edges(tets, deform=True)
    assert torch.equal(d1["senders"], d2["senders"]) and torch.equal(d1["receivers"], d2["receivers"])
    for k in range(12):                                          # the cache stays small
        util.triangles_to_edges(torch.from_numpy(synthetic.grid_triangles(3 + k, 3)).long())
    assert len(util._TOPOLOGY_CACHE) <= 8


def test_receiver_sorted_edge_storage_is_a_stable_permutation_and_transparent(monkeypatch):
    """plan.EdgeStorageOrder (experimental HGN_EDGE_STORAGE=receiver_sorted): stable sort by receiver, exact inverse, and the
    processor result is the same function of the graph -- checked with the CPU oracle on the permuted and on the original edge set."""
    import hgn_oracle as orc
    from hgn_b200 import plan, synthetic
    from hgn_b200.plan import EdgeStorageOrder
    # the product permutes rows with a CUDA kernel only; this host-side check of the bookkeeping brings its own torch stand-in
    monkeypatch.setattr(plan, "_permute_rows", lambda x, index32: x.index_select(0, index32.long()))
    s, r = synthetic.grid_edges_two_way(9, 7)
    order = EdgeStorageOrder(s, r)
    e = s.numel()
    assert torch.equal(order.perm.sort().values, torch.arange(e)) and torch.equal(order.perm[order.inverse], torch.arange(e))
    assert bool((order.receivers[1:] >= order.receivers[:-1]).all())                                   # sorted by receiver
    same = order.receivers[1:] == order.receivers[:-1]
    assert bool((order.perm[1:][same] > order.perm[:-1][same]).all())                                  # stable inside a receiver
    assert torch.equal(order.senders, s[order.perm]) and torch.equal(order.receivers, r[order.perm])
    feats = synthetic.seeded_tensor("so_e", (e, 128), 4)
    assert torch.equal(order.restore(order.store(feats)), feats)
    w = synthetic.seeded_state_dict(synthetic.processor_shapes(2, ["mesh_edges"], "sum"), 9)
    v = synthetic.seeded_tensor("so_v", (63, 128), 4)
    ref = orc.processor(w, "sum", "none", orc.MultiGraph([v.clone()], [orc.EdgeSet("mesh_edges", feats, s, r)]))
    out = orc.processor(w, "sum", "none", orc.MultiGraph([v.clone()], [orc.EdgeSet("mesh_edges", order.store(feats), order.senders, order.receivers)]))
    assert torch.allclose(out.node_features[0], ref.node_features[0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(order.restore(out.edge_sets[0].features), ref.edge_sets[0].features, rtol=1e-5, atol=1e-5)
    # distinct rows per 32 consecutive stored edges: the point of the exercise
    def distinct(x):
        x = x[: (x.numel() // 32) * 32].reshape(-1, 32).sort(dim=1).values
        return float((1 + (x[:, 1:] != x[:, :-1]).sum(1)).float().mean())
    assert distinct(order.receivers) < 0.6 * distinct(r)
