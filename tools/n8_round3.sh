N=${1:-8}; TAG=${2:-r02s4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 10 --warmup 6 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench exit $?"
ALAD_DEVICE_PHASES=0 timeout 300 $TR --master-port 29501 bench.py --gpus $N --steps 10 --warmup 6 --no-e2e > gpurun_out/${TAG}_bench_n${N}_nophases.json 2> gpurun_out/${TAG}_bench_n${N}_nophases.err; echo "nophases exit $?"
timeout 300 $TR --master-port 29510 tools/e2e_timeline.py > gpurun_out/${TAG}_tl_peer_n$N.json 2> gpurun_out/${TAG}_tl_peer_n$N.err; echo "timeline exit $?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_bench_n${N}_nophases.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["launches"], d["roofline"]["avg_launch_ms"], d["roofline"]["kernel_share_of_step"])
for l in open("gpurun_out/${TAG}_tl_peer_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l)
        if d["rank"] == 0:
            print("e2e step", d["step_ms"], "first s0", d["phases"][0]["s0"], "last s1", d["phases"][-1]["s1"])
            for r in d["resident_steps"]: print(" res", r)
PY
