#!/bin/bash
# Multi-GPU measurements of one box (run under gpurun --gpus N):  tools/run_scaling.sh N [full]
#   gpurun_out/r02_scale_nN_*.json[l]; `full` (8-GPU box): also the host-link topology, 4 of 8 GPUs chosen two ways,
#   and the batch of 256 frames
N=${1:-2}
FULL=${2:-}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
$TR --nproc-per-node $N --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 > $O/r02_scale_n${N}_bench.json 2> $O/r02_scale_n${N}_bench.err; echo "bench N=$N rc=$?"
for c in metric C3 C4; do
  $TR --nproc-per-node $N --master-port 29542 tools/bench_rowshard.py --config $c --iters 10 >> $O/r02_scale_n${N}_rowshard.jsonl 2>> $O/r02_scale_n${N}_rowshard.err; echo "rowshard $c rc=$?"
done
if [ -n "$FULL" ]; then
  python tools/h2d_topology.py > $O/r02_h2d_topology_n${N}.json 2> $O/r02_h2d_topology.err; echo "topology rc=$?"
  H=$((N / 2))
  BICOS_BENCH_DEVICES=first $TR --nproc-per-node $H --master-port 29544 bench.py --gpus $H --steps 8 --warmup 3 > $O/r02_scale_n${H}of${N}_first_bench.json 2> $O/r02_scale_n${H}of${N}_first.err; echo "bench $H of $N first rc=$?"
  $TR --nproc-per-node $H --master-port 29545 bench.py --gpus $H --steps 8 --warmup 3 > $O/r02_scale_n${H}of${N}_picked_bench.json 2> $O/r02_scale_n${H}of${N}_picked.err; echo "bench $H of $N picked rc=$?"
  $TR --nproc-per-node $N --master-port 29543 tools/bench_batch.py --frames 256 > $O/r02_scale_n${N}_batch.jsonl 2> $O/r02_scale_n${N}_batch.err; echo "batch rc=$?"
  for f in 1 3; do
    BICOS_BENCH_INFLIGHT=$f $TR --nproc-per-node $N --master-port 29546 bench.py --gpus $N --steps 6 --warmup 3 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('inflight $f e2e', l['e2e']['value'])"
  done
  for b in 96 384; do
    BICOS_B200_HOST_BAND_ROWS=$b $TR --nproc-per-node $N --master-port 29547 bench.py --gpus $N --steps 6 --warmup 3 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('band rows $b e2e', l['e2e']['value'])"
  done
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/r02_scale_n*bench.json")):
    try:
        l = json.load(open(f))
        print(f, "value", round(l["value"]), "e2e", round(l["e2e"]["value"]), l["e2e"].get("devices"), "row_sharded", l.get("row_sharded"))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat $O/r02_scale_n${N}_rowshard.jsonl
if [ -n "$FULL" ]; then cat $O/r02_scale_n${N}_batch.jsonl $O/r02_h2d_topology_n${N}.json; fi
