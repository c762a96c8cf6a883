N=${1:-8}; TAG=${2:-r02s3}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 200 $TR --master-port 29520 tools/h2d_multi_probe.py > gpurun_out/${TAG}_h2d_n$N.json 2> gpurun_out/${TAG}_h2d_n$N.err; echo "h2d exit $?"; cat gpurun_out/${TAG}_h2d_n$N.json
timeout 300 $TR --master-port 29510 tools/e2e_timeline.py > gpurun_out/${TAG}_tl_peer_n$N.json 2> gpurun_out/${TAG}_tl_peer_n$N.err; echo "timeline peer exit $?"
timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 10 --warmup 6 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench exit $?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["ms_per_step"], d["e2e"])
PY
