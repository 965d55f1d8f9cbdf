mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py pile --quick > gpurun_out/r2_mgpu_pile_n$N.json 2> gpurun_out/r2_mgpu_pile_n$N.err; echo "mgpu pile n$N rc=$?"; tail -2 gpurun_out/r2_mgpu_pile_n$N.err; cat gpurun_out/r2_mgpu_pile_n$N.json
bash tools/gpu_run4.sh $N
