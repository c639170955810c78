set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_r1g.json 2> $O/bench_r1g.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_ref_r1g.json 2> $O/bench_ref_r1g.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transform|search_kernel|refine_kernel" -c 400 --csv --log-file $O/launches_g.csv python bench.py --steps 2 --warmup 1 > $O/ncu_g.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"transform|search_kernel|refine_kernel" -c 4 -f -o $O/prof_r1g python tools/profile_one.py config2 > $O/ncu_full_g.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"transform|search_kernel" -c 3 -f -o $O/prof_r1g_c4 python tools/profile_one.py c4 > $O/ncu_full_g_c4.log 2>&1; echo "ncu c4 rc=$?"
cat $O/bench_r1g.json; cat $O/bench_ref_r1g.json
