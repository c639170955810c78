#!/bin/bash
# Parity + timing of the one-pass consistency kernel (search_mma3_kernel) against the popcount engine, and the
# two-pass kernels on the same descriptors for comparison. Every case under its own timeout.
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/search_onepass.txt}
mkdir -p "$(dirname "$out")"
: > "$out"
run() { echo "== $*" >> "$out"; timeout 60 tools/search_engines "$@" >> "$out" 2>&1; echo "exit $?" >> "$out"; }
run 256 4 4 2 0 64 0 2
run 128 1 4 2 0 64 0 2
run 1 1 4 2 0 64 0 2
run 130 3 4 2 1 64 0 2
run 300 5 4 2 1 64 0 2
run 1000 8 4 2 1 64 0 2
run 1280 32 4 2 3 64 0 2
run 1920 33 4 2 3 8 0 2
run 2048 32 4 2 3 64 0 2
run 8192 3 4 2 1 64 0 2
run 2048 192 4 2 5 64 0 2
run 2048 1536 4 2 5 64 0 2
run 2048 1536 4 2 5 64 2 2
run 1920 1200 4 2 5 64 0 2
run 1280 1024 4 2 5 64 0 2
# 256-bit descriptors
run 1 1 8 2 0 64 0 2
run 130 3 8 2 1 64 0 2
run 300 5 8 2 1 8 0 2
run 1000 8 8 2 1 64 0 2
run 2448 16 8 2 3 64 0 2
run 8192 3 8 2 1 64 0 2
run 2448 512 8 2 3 64 0 2
run 2448 512 8 2 3 64 2 2
run 4096 376 8 2 3 64 0 2
run 4096 376 8 2 3 64 2 2
grep -c identical "$out"; grep -c "MISMATCH\|error\|exit [1-9]" "$out"
grep "^cols" "$out" | sed 's/forward ties.*, //'
