#!/bin/bash
# Rebuilds librbfmap.so, then runs the given command on a B200 box: scripts/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
make -C "$ROOT/precice_b200/csrc" -j"$(nproc)" > /tmp/rbfmap_make.log 2>&1 || { tail -30 /tmp/rbfmap_make.log; exit 1; }
make -C "$ROOT/oracle" > /dev/null
cd "$ROOT"
exec /usr/local/graft/bin/gpurun "$@"
