"""The plan cache (hgn_b200/plan.py): identity keys (two views of one buffer never collide), entries die with their source tensor,
the bound is in bytes.  The cache machinery is device independent; the CUDA-side plans are covered by tests/test_gpu_parity.py."""
import gc

import torch

from hgn_b200 import plan


def test_views_of_one_buffer_do_not_collide_and_entries_die_with_their_source():
    plan.clear_plan_cache()
    base = torch.arange(12, dtype=torch.int64)
    a, b = base[:6], base.view(-1)[:6]                      # same data_ptr, same numel, different tensor objects
    assert a.data_ptr() == b.data_ptr()
    plan._remember(("plan", id(a), 3), (a,), "plan-a", 100)
    assert plan._lookup(("plan", id(a), 3), (a,)) == "plan-a"
    assert plan._lookup(("plan", id(b), 3), (b,)) is None    # keyed on the object, not on (data_ptr, numel)
    a.add_(0)                                               # an in-place change bumps the version: the entry is stale
    assert plan._lookup(("plan", id(a), 3), (a,)) is None
    c = torch.zeros(4, dtype=torch.int64)
    plan._remember(("plan", id(c), 2), (c,), "plan-c", 50)
    assert plan.plan_cache_stats()["entries"] == 1 and plan.plan_cache_stats()["bytes"] == 50
    del c
    gc.collect()
    d = torch.zeros(4, dtype=torch.int64)
    plan._remember(("plan", id(d), 2), (d,), "plan-d", 70)  # the insertion sweeps the entry whose source has died
    stats = plan.plan_cache_stats()
    assert stats["entries"] == 1 and stats["bytes"] == 70
    plan.clear_plan_cache()


def test_cache_is_bounded_in_bytes(monkeypatch):
    plan.clear_plan_cache()
    monkeypatch.setattr(plan, "_CACHE_BYTES", 1000)
    keep = [torch.zeros(1, dtype=torch.int64) for _ in range(6)]
    for i, t in enumerate(keep):
        plan._remember(("plan", id(t), 1), (t,), i, 300)
    stats = plan.plan_cache_stats()
    assert stats["bytes"] <= 1000 and stats["entries"] == 3
    assert plan._lookup(("plan", id(keep[0]), 1), (keep[0],)) is None        # least recently used went first
    assert plan._lookup(("plan", id(keep[5]), 1), (keep[5],)) == 5
    plan.clear_plan_cache()


def test_to_device_index_is_identity_on_the_same_device():
    t = torch.arange(5)
    assert plan.to_device_index(t, t.device) is t
