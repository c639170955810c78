N=8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
rm -f $O/r02_scale_n8_rowshard.jsonl
$TR --nproc-per-node $N --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 > $O/r02_scale_n8_bench.json 2> $O/r02_scale_n8_bench.err; echo "bench rc=$?"
for c in metric C3 C4; do
  $TR --nproc-per-node $N --master-port 29542 tools/bench_rowshard.py --config $c --iters 10 >> $O/r02_scale_n8_rowshard.jsonl 2>> $O/r02_scale_n8_rowshard.err; echo "rowshard $c rc=$?"
done
$TR --nproc-per-node $N --master-port 29543 tools/bench_batch.py --frames 256 > $O/r02_scale_n8_batch.jsonl 2> $O/r02_scale_n8_batch.err; echo "batch rc=$?"
$TR --nproc-per-node 4 --master-port 29545 bench.py --gpus 4 --steps 8 --warmup 3 > $O/r02_scale_n4of8_picked_bench.json 2> $O/r02_scale_n4of8_picked.err; echo "bench 4 of 8 rc=$?"
python - <<PY
import json
for f in ("$O/r02_scale_n8_bench.json", "$O/r02_scale_n4of8_picked_bench.json"):
    l = json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value", round(l["value"]), "e2e", round(l["e2e"]["value"]), l["e2e"].get("devices"), "row_sharded", l.get("row_sharded"))
PY
cat $O/r02_scale_n8_rowshard.jsonl $O/r02_scale_n8_batch.jsonl
