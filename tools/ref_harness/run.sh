#!/bin/bash
# Build and run the reference golden-vector harness inside the CubeFS module (needs Go >= 1.18; none in the
# build image).  Nothing is written into the reference tree: the harness is mapped in with `go build -overlay`.
#   REF=/path/to/cubefs tools/ref_harness/run.sh        -> tests/golden/rs_golden.json
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${REF:-/root/reference}"
OUT="${OUT:-$REPO/tests/golden/rs_golden.json}"
command -v go >/dev/null || { echo "go toolchain not found: parity stays unpinned (see DESIGN.md section 2)" >&2; exit 3; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cat > "$TMP/overlay.json" <<JSON
{"Replace": {"$REF/blobstore/cubeec_ref_harness/main.go": "$HERE/main.go"}}
JSON
export GOFLAGS=-mod=vendor GOCACHE="$TMP/gocache" CGO_ENABLED=0
if ! (cd "$REF" && go build -overlay "$TMP/overlay.json" -o "$TMP/harness" ./blobstore/cubeec_ref_harness); then
  # older toolchains / read-only quirks: build from a scratch copy of the module instead
  cp -r "$REF" "$TMP/cubefs" && mkdir -p "$TMP/cubefs/blobstore/cubeec_ref_harness" \
    && cp "$HERE/main.go" "$TMP/cubefs/blobstore/cubeec_ref_harness/" \
    && (cd "$TMP/cubefs" && go build -o "$TMP/harness" ./blobstore/cubeec_ref_harness)
fi
"$TMP/harness" > "$OUT.tmp"
mv "$OUT.tmp" "$OUT"
echo "wrote $OUT ($(wc -c < "$OUT") bytes); now run: python -m pytest tests/test_reference_golden.py -q"
