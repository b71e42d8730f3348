#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Builds the reference (NVlabs/nvdiffrast v0.4.0) from the sources where they lie under
# /root/reference, on a scratch copy under /tmp (the reference tree is read-only); only the installed package
# (pure-python front end + _nvdiffrast_c*.so) lands in the output directories, which are git-ignored but NOT
# gpurun-ignored, so the built files travel to the GPU box.  No reference source is copied into the repository.
#
#   oracle/_ref/       the UNMODIFIED reference: forward goldens, parity tests, `bench.py --impl reference`.
#   oracle/_ref_noca/  test-only variant for GRADIENT pins: identical sources except that the header's own
#                      "#if __CUDA_ARCH__ >= 700" switch around the warp-coalesced atomics
#                      (csrc/common/common.h:198) is forced to its #else branch (common.h:246-259: plain atomicAdd).
#                      The coalesced branch races on sm_100a (leader zeroes / reads the shared slot without a
#                      __syncwarp) and corrupts g_pos / g_attr / linear g_tex / antialias g_pos; the plain branch is
#                      the reference's own arithmetic with race-free accumulation.  One `sed` in /tmp, never committed.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${NVDR_REFERENCE:-/root/reference}
if [ ! -d "$SRC" ]; then echo "reference tree $SRC absent (GPU box?) - using prebuilt oracle/_ref* if any"; exit 0; fi
export TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=${MAX_JOBS:-8}

build_one() {   # $1 = output dir, $2 = "noca" to force the plain-atomics branch
  local OUT="$1"
  if ls "$OUT"/_nvdiffrast_c*.so >/dev/null 2>&1; then echo "$(basename "$OUT") already built"; return 0; fi
  local TMP; TMP=$(mktemp -d /tmp/nvdr_src.XXXXXX)
  cp -r "$SRC"/. "$TMP"/
  if [ "${2:-}" = "noca" ]; then
    sed -i 's|^#if __CUDA_ARCH__ >= 700 // Warp match instruction.*$|#if 0 // test build: plain atomicAdd branch|' "$TMP/csrc/common/common.h"
    grep -q '^#if 0 // test build: plain atomicAdd branch' "$TMP/csrc/common/common.h" || { echo "noca: switch not found in common.h"; rm -rf "$TMP"; return 1; }
  fi
  python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$OUT" "$TMP"
  rm -rf "$TMP"
  ls -la "$OUT"
}

build_one "$REPO/oracle/_ref"
build_one "$REPO/oracle/_ref_noca" noca
