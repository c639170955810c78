#!/bin/bash
# All multi-GPU measurements of one box in one go (run under gpurun --gpus N):
#   tools/run_scaling.sh N   ->  gpurun_out/scale_nN_*.json[l]
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
O=gpurun_out
$TR --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > $O/scale_n${N}_bench.json 2> $O/scale_n${N}_bench.err; echo "bench rc=$?"
for c in metric C3 C4; do
  $TR --master-port 29542 tools/bench_rowshard.py --config $c --iters 10 >> $O/scale_n${N}_rowshard.jsonl 2>> $O/scale_n${N}_rowshard.err; echo "rowshard $c rc=$?"
done
$TR --master-port 29543 tools/bench_batch.py --frames 256 > $O/scale_n${N}_batch.jsonl 2> $O/scale_n${N}_batch.err; echo "batch rc=$?"
cat $O/scale_n${N}_bench.json $O/scale_n${N}_rowshard.jsonl $O/scale_n${N}_batch.jsonl
