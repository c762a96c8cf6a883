run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29510 tools/e2e_timeline.py > gpurun_out/r02s2_tl_$tag.json 2> gpurun_out/r02s2_tl_$tag.err; grep -c . gpurun_out/r02s2_tl_$tag.json; }
run nt256 NCCL_NTHREADS=256
run nt128 NCCL_NTHREADS=128 NCCL_MAX_NCHANNELS=8
run nt64 NCCL_NTHREADS=64 NCCL_MAX_NCHANNELS=4
