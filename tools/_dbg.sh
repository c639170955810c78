cd /root/repo
run() { timeout 60 tools/search_engines "$@" 2>&1 | grep "^cols\|rror\|row" | head -8; }
for v in 2; do
run 256 4 4 3 0 64 $v
run 1000 8 4 3 1 64 $v
run 2448 16 8 3 3 64 $v
run 300 500 8 2 1 64 $v
run 384 700 4 3 1 64 $v
run 2048 1536 4 2 3 64 $v
run 2048 1536 4 1 3 64 $v
run 4096 750 8 1 2 64 $v
run 1280 1024 4 2 3 64 $v
run 1280 1024 4 2 3 64 1
run 1920 1200 4 2 3 64 $v
run 1920 1200 4 2 3 64 1
done
