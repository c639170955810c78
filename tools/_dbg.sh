cd /root/repo
run() { timeout 60 tools/search_engines "$@" 2>&1 | grep "^cols\|rror\|row" | head -5 | cut -c1-185; }
run 256 4 4 3 0 64 1 1
run 1000 8 4 3 1 64 1 1
run 2448 16 8 3 1 64 1 1
run 300 500 8 2 1 64 2 1
run 384 700 4 3 1 64 2 1
run 2048 1536 4 2 3 64 2 1
run 2048 1536 4 3 3 64 2 1
run 4096 750 8 1 2 64 2 1
run 1280 1024 4 2 3 64 1 1
