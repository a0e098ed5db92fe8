mkdir -p gpurun_out/r2r
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2r/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/r2r/gpu_tests.log
