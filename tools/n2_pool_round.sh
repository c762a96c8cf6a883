N=${1:-2}; TAG=${2:-pool}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 $TR --master-port 29533 tools/dist_check.py > gpurun_out/${TAG}_dist_check_n$N.log 2>&1; echo "dist_check exit $?"; grep "rank 0\|dist_check\|Error\|error" gpurun_out/${TAG}_dist_check_n$N.log | grep -v "^W1" | tail -14
timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 8 --warmup 5 --no-e2e > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench_n$N.err
ALAD_NO_POOL=1 timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 8 --warmup 5 --no-e2e > gpurun_out/${TAG}_bench_n${N}_nopool.json 2> gpurun_out/${TAG}_bench_n${N}_nopool.err; echo "nopool exit $?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_bench_n${N}_nopool.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, d["ms_per_step"], d["roofline"]["launches"], d["roofline"]["achieved"], d["roofline"]["kernel_share_of_step"], d["shard_balance"].get("work_pool"), d["recall_at_1"])
PY
