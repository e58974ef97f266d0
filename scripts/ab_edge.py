"""development: A/B of two builds of the library on the cfg5 edge update (forward + backward kernels): per-launch times from the
library's kernel timers and a bitwise comparison of every output.  usage: python scripts/ab_edge.py [old.so]   (parent mode);
the child mode (AB_CHILD=<lib path or ''>) runs one build and writes sha1 digests + times as JSON on stdout."""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if "AB_CHILD" in os.environ:
    import torch
    sys.path.insert(0, os.path.join(ROOT, "hyper-graph-nets_b200"))
    from hgn_b200 import _cabi
    if os.environ["AB_CHILD"]:
        _cabi.LIB_PATH = os.environ["AB_CHILD"]
    from hgn_b200 import ops, synthetic
    from hgn_b200.plan import segment_plan
    dev = "cuda"
    W, H = int(os.environ.get("GRID_W", 1000)), int(os.environ.get("GRID_H", 1000))
    s, r = synthetic.grid_edges_two_way(W, H)
    n, E = W * H, s.numel()
    s, r = s.to(dev), r.to(dev)
    sd = synthetic.seeded_state_dict(synthetic.mlp_shapes("m", 384), 3)
    w = [sd[f"m.0.layers.linear_{k}.{p}"].to(dev).requires_grad_(True) for k in range(3) for p in ("weight", "bias")]
    w += [sd["m.1.weight"].to(dev).requires_grad_(True), sd["m.1.bias"].to(dev).requires_grad_(True)]
    g = torch.Generator(device=dev).manual_seed(1)
    v = torch.randn(n, 128, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    e = torch.randn(E, 128, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    gup = torch.randn(E, 128, device=dev, generator=g).to(torch.bfloat16)
    gagg = torch.randn(n, 128, device=dev, generator=g).to(torch.bfloat16)
    sp, rp = segment_plan(s, n), segment_plan(r, n)
    cache = {}
    def it():
        for t in [v, e] + w:
            t.grad = None
        out, agg = ops.edge_update(w, cache, v, e, sp, rp, True)
        torch.autograd.backward([out, agg], [gup, gagg])
        return out, agg
    for _ in range(2): out, agg = it()
    _cabi.profile(True)
    for _ in range(4): out, agg = it()
    rep = _cabi.profile_report()
    torch.cuda.synchronize()
    dig = lambda t: hashlib.sha1(t.detach().contiguous().view(torch.uint8).cpu().numpy().tobytes()).hexdigest()[:16]
    res = {"ms": {k["name"]: round(k["ms"] / k["launches"], 4) for k in rep},
           "sha": {"out": dig(out), "agg": dig(agg), "dv": dig(v.grad), "de": dig(e.grad), **{f"w{i}": dig(p.grad) for i, p in enumerate(w)}}}
    print("AB-RESULT " + json.dumps(res))
    sys.exit(0)
# variants: "name=libpath" or "name=libpath,ENV=VALUE,..." (empty libpath = the in-tree library); default: build/old vs in-tree
variants = sys.argv[1:] or ["old=" + os.path.join(ROOT, "hyper-graph-nets_b200", "build", "old", "libhgn_b200_old.so"), "new="]
res = {}
for spec in variants:
    tag, rest = spec.split("=", 1)
    parts = rest.split(",")
    lib, env = parts[0], dict(kv.split("=", 1) for kv in parts[1:])
    if lib and not os.path.isabs(lib):
        lib = os.path.join(ROOT, lib)
    run = subprocess.run([sys.executable, __file__], env=dict(os.environ, AB_CHILD=lib, **env), capture_output=True, text=True, timeout=600)
    line = [ln for ln in run.stdout.splitlines() if ln.startswith("AB-RESULT ")]
    if not line:
        print(tag, "FAILED", run.stdout[-2000:], run.stderr[-3000:]); continue
    res[tag] = json.loads(line[-1][10:])
    print(tag, res[tag]["ms"])
tags = list(res)
for t in tags[1:]:
    same = {k: res[tags[0]]["sha"][k] == res[t]["sha"][k] for k in res[tags[0]]["sha"]}
    print(f"{t} bitwise identical to {tags[0]}:", all(same.values()), {k: v for k, v in same.items() if not v})
