set -x
mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_openmm_shim.py -x -q > gpurun_out/r2e/shim_tests.log 2>&1; echo "shim tests rc=$?"
tail -5 gpurun_out/r2e/shim_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench_driver_cmd.json 2> gpurun_out/r2e/bench_driver_cmd.err; echo "bench rc=$?"
tail -5 gpurun_out/r2e/bench_driver_cmd.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2e/bench_reference_arm.json 2> gpurun_out/r2e/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2e/bench_default.json 2> gpurun_out/r2e/bench_default.err; echo "bench default rc=$?"
tail -5 gpurun_out/r2e/bench_default.err
