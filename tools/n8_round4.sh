N=${1:-8}; TAG=${2:-r02s5}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
for pass in 1 2; do
  timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 12 --warmup 6 --no-e2e > gpurun_out/${TAG}_bench_n${N}_pool_$pass.json 2> gpurun_out/${TAG}_bench_n${N}_pool_$pass.err; echo "pool $pass exit $?"
  ALAD_NO_POOL=1 timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 12 --warmup 6 --no-e2e > gpurun_out/${TAG}_bench_n${N}_nopool_$pass.json 2> gpurun_out/${TAG}_bench_n${N}_nopool_$pass.err; echo "nopool $pass exit $?"
done
timeout 300 $TR --master-port 29502 bench.py --gpus $N --steps 10 --warmup 6 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_n${N}*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, round(d["ms_per_step"], 3), (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["launches"], round(d["roofline"]["achieved"], 1), round(d["roofline"]["kernel_share_of_step"], 3), (d["shard_balance"] or {}).get("work_pool"))
PY
