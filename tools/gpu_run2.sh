mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_box2.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -x --durations=10 > gpurun_out/r2_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_sharded.log
tail -15 gpurun_out/r2_pytest_sharded.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py slab > gpurun_out/r2_mgpu_slab_n2.json 2> gpurun_out/r2_mgpu_slab_n2.err; echo "slab rc=$?"
tail -3 gpurun_out/r2_mgpu_slab_n2.err; cat gpurun_out/r2_mgpu_slab_n2.json
