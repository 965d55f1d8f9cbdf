mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_v10_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_v10_pytest_gpu.log
bash tools/gpu_grid_scale.sh
