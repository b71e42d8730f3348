#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference (NVlabs/nvdiffrast v0.4.0) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored,
# so the built files travel to the GPU box). No reference source is copied into the repository:
# the build happens on a scratch copy under /tmp (the reference tree is read-only) and only the
# installed package (pure-python front end + _nvdiffrast_c*.so) lands in oracle/_ref/.
# Used by tests/ (parity against the real reference on the GPU) and bench.py --impl reference.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${NVDR_REFERENCE:-/root/reference}
OUT="$REPO/oracle/_ref"
if [ ! -d "$SRC" ]; then echo "reference tree $SRC absent (GPU box?) - using prebuilt $OUT if any"; exit 0; fi
if ls "$OUT"/_nvdiffrast_c*.so >/dev/null 2>&1; then echo "oracle/_ref already built"; exit 0; fi
TMP=$(mktemp -d /tmp/nvdr_src.XXXXXX)
cp -r "$SRC"/. "$TMP"/
export TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=${MAX_JOBS:-8}
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$OUT" "$TMP" 
rm -rf "$TMP"
ls -la "$OUT"
