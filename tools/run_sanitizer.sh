#!/bin/bash
# compute-sanitizer over small shapes of both search engines and both tensor-core kernel variants, with and without
# the column term (tools/search_engines: the two engines must also agree bit for bit), then one small whole match
# through the C++ API (transform + search + refine). ONE tool per gpurun call (B200_PROFILING.md):
#   tools/run_sanitizer.sh memcheck|racecheck|synccheck|initcheck [out]
cd "$(dirname "$0")/.."
tool=${1:-memcheck}
out=${2:-gpurun_out/r02_sanitizer_${tool}.txt}
mkdir -p "$(dirname "$out")"
: > "$out"
CS="compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20"
run() { echo "== $*" >> "$out"; timeout 300 $CS "$@" >> "$out" 2>&1; echo "exit $?" >> "$out"; }
#                 cols rows K flags reps pool variant colterm
run tools/search_engines 256 4 4 3 0 64 1 0
run tools/search_engines 256 4 4 3 0 64 1 1
run tools/search_engines 1000 6 4 2 0 64 1 1
run tools/search_engines 130 3 4 3 0 64 1 0
run tools/search_engines 1 1 4 3 0 64 1 0
run tools/search_engines 600 4 8 3 0 64 1 1
run tools/search_engines 600 3 12 3 0 64 1 0
run tools/search_engines 600 3 16 1 0 64 1 0
run tools/search_engines 512 8 4 2 0 64 2 1
run tools/search_engines 520 6 4 3 0 64 2 0
run tools/search_engines 700 5 8 3 0 64 2 1
run tools/search_engines 300 4 8 1 0 64 2 0
run tools/search_engines 300 2 1 3 0 64 0 0
run tools/search_engines 300 2 2 2 0 64 0 0
# a small whole match (33 images, 24 x 272, subpixel + consistency) through BICOS::match, input written by the test helper
python - <<'PY' >> "$out" 2>&1
import importlib.util, numpy as np, sys
sys.path.insert(0, ".")
spec = importlib.util.spec_from_file_location("t", "tests/test_cpp_api.py")
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
from libbicos_b200 import synth
left, right, _ = synth.make_stacks(33, 256, 272, np.uint8, seed=3, row0=64, rows=24)
m._write_input("gpurun_out/sanitizer_api_in.bin", left, right,
               dict(nxcorr_threshold=0.96, min_variance=2.0, subpixel_step=0.1, consistency=True, max_lr_diff=1))
PY
run tests/cpp/build/api_check gpurun_out/sanitizer_api_in.bin gpurun_out/sanitizer_api_out.bin
echo "runs: $(grep -c '^== ' "$out"), clean exits: $(grep -c '^exit 0' "$out"), reports with errors: $(grep -c 'ERROR SUMMARY: [1-9]' "$out")"
grep -h "ERROR SUMMARY\|identical\|MISMATCH\|^exit" "$out" | paste - - - | head -40
