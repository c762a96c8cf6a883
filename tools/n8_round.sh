# One 8-GPU box pass (N = 8 costs 8x the box time: keep it short).  Usage: bash tools/n8_round.sh [N] [tag]
N=${1:-8}; TAG=${2:-r02s2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 10 --warmup 6 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench exit $?"
ALAD_NO_BALANCE=1 timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 10 --warmup 6 --no-e2e > gpurun_out/${TAG}_bench_n${N}_nobal.json 2> gpurun_out/${TAG}_bench_n${N}_nobal.err; echo "nobal exit $?"
timeout 300 $TR --master-port 29502 tools/two_stage_probe.py > gpurun_out/${TAG}_two_stage_n$N.json 2> gpurun_out/${TAG}_two_stage_n$N.err; echo "two-stage exit $?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_bench_n${N}_nobal.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["avg_launch_ms"], d["roofline"]["kernel_share_of_step"], d["shard_balance"])
for l in open("gpurun_out/${TAG}_two_stage_n$N.json"):
    if l.startswith("{"): print(l.strip())
PY
