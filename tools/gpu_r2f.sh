set -x
mkdir -p gpurun_out/r2f
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2f/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2f/gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f/bench_driver_cmd.json 2> gpurun_out/r2f/bench_driver_cmd.err; echo "bench rc=$?"
tail -5 gpurun_out/r2f/bench_driver_cmd.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f/bench_reference_arm.json 2> gpurun_out/r2f/bench_reference_arm.err; echo "ref rc=$?"
