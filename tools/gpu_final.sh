# what the driver runs at round end, plus the default bench (final code): GPU tests, smoke, bench reference arm, driver command line, default
mkdir -p gpurun_out/final
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/final/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_driver_cmd.json 2> gpurun_out/final/bench_driver_cmd.err; echo "bench rc=$?"
timeout 900 python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; echo "bench default rc=$?"
