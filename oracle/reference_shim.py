"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (CemOezcan/hyper-graph-nets).

``/root/reference`` is mounted only in the build container.  ``oracle/make_ref.sh`` (run by ``__graft_entry__.build()``
there) stages the reference's unmodified Python packages under the git-ignored ``oracle/_ref/``, which travels to the
GPU box with the snapshot like a built ``.so``; this loader takes the mount when present and ``oracle/_ref`` otherwise.
Nothing at run time reads ``/root/reference`` on the GPU box.  Users: ``tests/golden/make_golden*.py`` (golden vectors),
the ``not gpu`` tests that cross-check ``oracle/hgn_oracle.py``, the ``-m gpu`` drop-in tests (the reference's own
``FlagModel`` / ``PlateModel`` / ``CylinderModel`` on the CPU as the checker) and ``bench.py``'s reference arm /
``cpu_baseline`` (kind "reference").  Never the product path.

The reference imports a handful of packages that are absent from this image.
They are replaced by in-memory stub modules *before* anything from ``/root/reference/src`` is imported:

* ``torch_scatter`` 2.0.9 (``requirements.txt:7``; sole call sites ``src/util.py:117-130``) -- restated
  on ``torch.Tensor.scatter_add_`` / ``scatter_reduce_`` with the 2.0.9 semantics: empty segments give 0
  for every reduction, ``mean`` divides by ``count.clamp(min=1)``, ``max``/``min`` return ``(out, arg)``
  and route the gradient to exactly ONE winner per (segment, column) -- the first edge in input order,
  like the torch_scatter CPU reducer.  This is the one documented deviation from "the reference's own
  torch/torch_scatter path".
* ``hdbscan``, ``seaborn``, ``colorcet``, ``matplotlib(.pyplot/.animation/.tri)``, ``tfrecord(.torch)``
  -- empty modules (plot / dataset code that is never reached on the hot path).
"""
from __future__ import annotations

import os
import sys
import types

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _find_root() -> str:
    env = os.environ.get("HGN_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isdir(os.path.join(cand, "src", "migration")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "migration"))


# ----------------------------------------------------------------------------------------------
# torch_scatter 2.0.9 shim
# ----------------------------------------------------------------------------------------------
class _ScatterArg(torch.autograd.Function):
    """scatter_max / scatter_min with torch_scatter semantics (single first-in-order winner)."""

    @staticmethod
    def forward(ctx, src, index, dim_size, is_max):
        assert src.shape == index.shape
        flat_shape = src.shape
        out = torch.zeros((dim_size,) + tuple(flat_shape[1:]), dtype=src.dtype)
        red = "amax" if is_max else "amin"
        out.scatter_reduce_(0, index, src, red, include_self=False)
        # arg = first position (in input order) that attains the winning value; dim_size-sentinel = E
        E = src.shape[0]
        pos = torch.arange(E, dtype=torch.int64).view((E,) + (1,) * (src.dim() - 1)).expand_as(src)
        winner = src == out.gather(0, index)
        cand = torch.where(winner, pos, torch.full_like(pos, E))
        arg = torch.full(out.shape, E, dtype=torch.int64)
        arg.scatter_reduce_(0, index, cand, "amin", include_self=True)
        ctx.save_for_backward(arg)
        ctx.E = E
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        E = ctx.E
        grad_src = torch.zeros((E + 1,) + tuple(grad_out.shape[1:]), dtype=grad_out.dtype)
        grad_src.scatter_(0, arg, grad_out)  # one winner per (segment, column); sentinel row E dropped
        return grad_src[:E], None, None, None


def _scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    size = list(src.shape)
    size[0] = int(dim_size) if dim_size is not None else int(index.max()) + 1
    if out is None:
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(0, index, src)


def _scatter_mean(src, index, dim=0, out=None, dim_size=None):
    total = _scatter_add(src, index, dim, None, dim_size)
    ones = torch.ones(index.shape[0], dtype=src.dtype)
    idx1 = index.reshape(index.shape[0], -1)[:, 0]
    count = torch.zeros(total.shape[0], dtype=src.dtype).scatter_add_(0, idx1, ones).clamp_(min=1)
    return total / count.view((-1,) + (1,) * (total.dim() - 1))


def _scatter_max(src, index, dim=0, out=None, dim_size=None):
    return _ScatterArg.apply(src, index, int(dim_size), True)


def _scatter_min(src, index, dim=0, out=None, dim_size=None):
    return _ScatterArg.apply(src, index, int(dim_size), False)


def _scatter_std(src, index, dim=0, out=None, dim_size=None, unbiased=True):
    mean = _scatter_mean(src, index, dim, None, dim_size)
    ones = torch.ones(index.shape[0], dtype=src.dtype)
    idx1 = index.reshape(index.shape[0], -1)[:, 0]
    count = torch.zeros(mean.shape[0], dtype=src.dtype).scatter_add_(0, idx1, ones)
    var = _scatter_add((src - mean.gather(0, index)) ** 2, index, dim, None, dim_size)
    denom = (count - 1 if unbiased else count).clamp(min=1).view((-1,) + (1,) * (var.dim() - 1))
    return (var / (denom + 1e-6)).sqrt()


def _install_stub_modules() -> None:
    ts = types.ModuleType("torch_scatter")
    ts.scatter_add = _scatter_add
    ts.scatter_sum = _scatter_add
    ts.scatter_mean = _scatter_mean
    ts.scatter_max = _scatter_max
    ts.scatter_min = _scatter_min
    ts.scatter_std = _scatter_std
    sys.modules.setdefault("torch_scatter", ts)
    for name in ("hdbscan", "seaborn", "colorcet", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.animation", "matplotlib.tri", "tfrecord", "tfrecord.torch"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            sys.modules[name] = mod
    sys.modules["tfrecord.torch"].TFRecordDataset = type("TFRecordDataset", (), {})
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].tri = sys.modules["matplotlib.tri"]
    sys.modules["tfrecord"].torch = sys.modules["tfrecord.torch"]
    sys.modules["seaborn"].color_palette = lambda *a, **k: []
    sys.modules["colorcet"].glasbey = []


_loaded = False


def load():
    """Put the reference on ``sys.path`` (with stubs) and return its ``src`` package."""
    global _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _loaded:
        os.environ.setdefault("WANDB_MODE", "disabled")
        _install_stub_modules()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _loaded = True
    import src  # noqa: F401  (the reference's top-level package)
    import src.util  # noqa: F401
    return sys.modules["src"]
