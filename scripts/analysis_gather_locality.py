"""Round-2 planning, CPU only: how many distinct table rows does one warp instruction of the edge kernels' thread-per-row gathers
touch on the cfg5 mesh?  (The load pipe retires about one 32-byte sector per cycle, DESIGN.md s3.2: 32 distinct rows per warp
instruction is the worst case.)  Compared for the reference's edge order (two_way_connectivity: first half sorted by sender, second
half the reversed copies = sorted by receiver) and for a receiver-sorted storage order."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hyper-graph-nets_b200"))
from hgn_b200 import synthetic

W = int(os.environ.get("GRID_W", 1000)); H = int(os.environ.get("GRID_H", 1000))
s, r = (t.numpy() for t in synthetic.grid_edges_two_way(W, H))
E = s.size

def distinct_per_warp(idx):
    pad = (-idx.size) % 32
    x = np.concatenate([idx, np.full(pad, idx[-1])]).reshape(-1, 32)
    x = np.sort(x, axis=1)
    return 1 + (np.diff(x, axis=1) != 0).sum(1)

def lines_touched_per_tile(idx, rows_per_tile=128):
    pad = (-idx.size) % rows_per_tile
    x = np.concatenate([idx, np.full(pad, idx[-1])]).reshape(-1, rows_per_tile)
    x = np.sort(x, axis=1)
    return 1 + (np.diff(x, axis=1) != 0).sum(1)

for name, order in (("reference order", np.arange(E)), ("receiver-sorted (stable)", np.argsort(r, kind="stable")),
                    ("sender-sorted (stable)", np.argsort(s, kind="stable"))):
    ss, rr = s[order], r[order]
    ds, dr = distinct_per_warp(ss), distinct_per_warp(rr)
    ts, tr = lines_touched_per_tile(ss), lines_touched_per_tile(rr)
    print(f"{name:26s} distinct rows per warp instruction: senders {ds.mean():5.1f}  receivers {dr.mean():5.1f}   "
          f"distinct rows per 128-edge tile: senders {ts.mean():6.1f}  receivers {tr.mean():6.1f}")
# sector requests per tile of the backward kernel: Ps[s] 4 instr/half-row x 2 halves, Pr[r] same, grad_agg[r] same, dense dO 8 (always 32 rows)
def sectors(ds, dr):
    per_warp = 8 * ds.mean() + 8 * dr.mean() + 8 * dr.mean() + 8 * 32
    return per_warp * 4          # 4 row groups of 32 per tile (x 2 column halves is in the 8 instructions)
for name, order in (("reference order", np.arange(E)), ("receiver-sorted", np.argsort(r, kind="stable"))):
    print(f"{name:18s}: ~{sectors(distinct_per_warp(s[order]), distinct_per_warp(r[order])):6.0f} distinct-line requests per tile (worst case 4096)")
