# Same-box A/B at N = 2: shard balancer on/off x device-resident caption phases on/off (bench.py --no-e2e).
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 6 --warmup 4 --no-e2e > gpurun_out/r02_n2_$tag.json 2> gpurun_out/r02_n2_$tag.err; python - <<PY
import json
for l in open("gpurun_out/r02_n2_$tag.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$tag", d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["launches"], d["roofline"]["kernel_share_of_step"], d["shard_balance"])
PY
}
mkdir -p gpurun_out
run default A=1
run nobal ALAD_NO_BALANCE=1
run phases ALAD_DEVICE_PHASES=1
run phases_nobal ALAD_NO_BALANCE=1 ALAD_DEVICE_PHASES=1
