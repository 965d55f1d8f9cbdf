mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_box.txt; nproc >> gpurun_out/r2_box.txt; free -g >> gpurun_out/r2_box.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/r2_pytest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_1.log
timeout 300 python tools/time_np_flags.py c2 0 $((1<<23)) 0 $((1<<23)) > gpurun_out/r2_deepfirst_c2.json 2> gpurun_out/r2_deepfirst_c2.err
timeout 300 python tools/time_np_flags.py c3 0 $((1<<23)) > gpurun_out/r2_deepfirst_c3.json 2> gpurun_out/r2_deepfirst_c3.err
tail -5 gpurun_out/r2_pytest_1.log; cat gpurun_out/r2_deepfirst_c2.json
