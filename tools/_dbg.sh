cd /root/repo
run() { echo "== $*"; timeout 60 tools/search_engines "$@" 2>&1 | grep "^cols\|error"; }
run 4096 1 8 1 0
run 4096 4 8 1 0
run 256 400 8 1 0
run 256 400 8 0 0
run 512 400 8 0 0
run 256 400 12 1 0
run 256 400 16 1 0
run 256 400 4 1 0
run 2448 16 8 0 0
