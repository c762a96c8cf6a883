# N-GPU pass: per-phase timeline of the e2e step with both caption exchanges + the bench line.  bash tools/n8_timeline.sh [N] [tag]
N=${1:-8}; TAG=${2:-r02s2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 $TR --master-port 29510 tools/e2e_timeline.py > gpurun_out/${TAG}_tl_peer_n$N.json 2> gpurun_out/${TAG}_tl_peer_n$N.err; echo "timeline peer exit $?"
ALAD_EXCHANGE=nccl timeout 300 $TR --master-port 29511 tools/e2e_timeline.py > gpurun_out/${TAG}_tl_nccl_n$N.json 2> gpurun_out/${TAG}_tl_nccl_n$N.err; echo "timeline nccl exit $?"
timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 10 --warmup 6 > gpurun_out/${TAG}_bench_peer_n$N.json 2> gpurun_out/${TAG}_bench_peer_n$N.err; echo "bench exit $?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_peer_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["ms_per_step"], d["e2e"])
PY
