#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Stage the LIVE reference (CemOezcan/hyper-graph-nets, pure Python) where the GPU box can import it:
# /root/reference is mounted only in the build container, so its Python packages (src/, util/, configs/) are staged, unmodified,
# under the git-ignored oracle/_ref/ (listed in .gitignore, NOT in .gpurunignore: it travels with the gpurun snapshot exactly like
# a built .so).  Nothing is ever committed from there.  oracle/reference_shim.py loads /root/reference when mounted, else oracle/_ref.
# Used by: tests (-m gpu drop-in parity against the reference's own FlagModel / PlateModel / CylinderModel on the CPU) and
# bench.py --impl reference / cpu_baseline (kind "reference").  Never by the product path.
set -euo pipefail
SRC=${HGN_REFERENCE_SRC:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [[ ! -d "$SRC/src/migration" ]]; then
  echo "make_ref: $SRC not mounted; keeping $DST as is"; exit 0
fi
rm -rf "$DST"; mkdir -p "$DST"
for d in src util configs; do cp -r "$SRC/$d" "$DST/$d"; done
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && find src util configs -type f \( -name '*.py' -o -name '*.yaml' \) -print0 | sort -z | xargs -0 sha256sum ) > "$DST/MANIFEST.sha256"
echo "make_ref: staged $(wc -l < "$DST/MANIFEST.sha256") files from $SRC into $DST"
