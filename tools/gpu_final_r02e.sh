# final code of round 2 (call-free square root in the generator, MMVT LOCAL kernel with its own load group in pool mode): GPU tests, smoke, the driver's bench command
mkdir -p gpurun_out/final_e
timeout 360 python -m pytest tests -m gpu -x -q > gpurun_out/final_e/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/final_e/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_e/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_e/smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/final_e/bench_driver_cmd.json 2> gpurun_out/final_e/bench_driver_cmd.err; echo "bench rc=$?"
