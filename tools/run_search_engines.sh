#!/bin/bash
# Parity + timing sweep of the two search engines (tools/search_engines) on the GPU box.
# Every case runs under its own timeout; a trap or a mismatch does not stop the sweep.
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/search_engines.txt}
mkdir -p "$(dirname "$out")"
: > "$out"
run() { echo "== $*" >> "$out"; timeout 60 tools/search_engines "$@" >> "$out" 2>&1; echo "exit $?" >> "$out"; }
run 256 4 4 0 0
run 256 4 4 3 0
run 2048 32 4 3 3
run 2048 32 4 2 3
run 2048 32 4 1 3
run 1280 32 4 2 3
run 1000 8 4 3 1
run 130 3 4 3 1
run 1 1 4 3 0
run 2448 16 8 3 3
run 4096 16 8 1 3
run 600 4 12 3 1
run 600 4 16 3 1
run 2048 1536 4 2 3
run 4096 750 8 1 2
grep -c identical "$out"; grep -c "MISMATCH\|error" "$out"
