"""cuobjdump -sass opcode histogram per kernel of the in-tree objects (evidence that the hot kernels are tcgen05 / TMA code):

    python scripts/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt
"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "hyper-graph-nets_b200", "build")
KEEP = re.compile(r"UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTCATOM|SYNCS|LDGSTS|USETMAXREG|HMMA|^LDG|^STG|^LDS|^STS|^BAR|^ATOM|^RED|SHFL|FFMA|FADD2|FMUL2|F2FP|^LDL|^STL")
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
print(f"# cuobjdump -sass opcode histogram per kernel (sm_100a objects of commit {commit}); opcodes grouped by their first two dot-fields")
for obj in ("edge_tc", "edge_fwd_tc", "mlp_tc", "segment", "world_edges", "peer", "mlp_f32"):
    path = os.path.join(BUILD, obj + ".o")
    if not os.path.exists(path):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, hist, total = None, None, 0
    def flush():
        if name is not None:
            print(f"\n== {obj}.o :: {name} ({total} instructions)")
            for op, n in sorted(hist.items(), key=lambda kv: (-kv[1], kv[0])):
                print(f"   {op:36s} {n:6d}")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            raw = m.group(1)
            mm = re.match(r"_ZN3hgn\d+([A-Za-z0-9_]+?)(?:ILb\d+EEE|E)", raw)
            name, hist, total = (mm.group(1) if mm else raw), collections.Counter(), 0
            if "ILb1E" in raw: name += "<true>"
            if "ILb0E" in raw: name += "<false>"
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name is not None:
            total += 1
            op = ".".join(m.group(1).split(".")[:3])
            if KEEP.search(op):
                hist[op] += 1
    flush()
