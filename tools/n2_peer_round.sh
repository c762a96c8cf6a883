# 2-GPU pass for the peer-window caption exchange: parity (sharded == unsharded), per-phase timeline, bench line
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 300 $TR --master-port 29533 tools/dist_check.py > gpurun_out/peer_dist_check_n$N.log 2>&1; echo "dist_check exit $?"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/peer_dist_check_n$N.log | tail -12
timeout 300 $TR --master-port 29510 tools/e2e_timeline.py > gpurun_out/peer_tl_n$N.json 2> gpurun_out/peer_tl_n$N.err; echo "timeline exit $?"; tail -3 gpurun_out/peer_tl_n$N.err
timeout 300 $TR --master-port 29500 bench.py --gpus $N --steps 6 --warmup 4 > gpurun_out/peer_bench_n$N.json 2> gpurun_out/peer_bench_n$N.err; echo "bench exit $?"; tail -3 gpurun_out/peer_bench_n$N.err
