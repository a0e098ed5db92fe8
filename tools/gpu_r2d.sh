set -x
mkdir -p gpurun_out/r2d
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -15 gpurun_out/r2d/gpu_tests.log
