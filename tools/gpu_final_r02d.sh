# final code of round 2, last session (MMVT LOCAL kernel: normals of later tiles after the tile's loads): what the driver runs at round end,
# plus the decide-mode sweep
mkdir -p gpurun_out/final_d
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final_d/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/final_d/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_d/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_d/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_d/bench_reference_arm.json 2> gpurun_out/final_d/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_d/bench_driver_cmd.json 2> gpurun_out/final_d/bench_driver_cmd.err; echo "bench rc=$?"
timeout 900 python bench.py > gpurun_out/final_d/bench_default.json 2> gpurun_out/final_d/bench_default.err; echo "bench default rc=$?"
SWEEP_STEPS=2048 timeout 900 python tools/decide_mode_sweep.py > gpurun_out/final_d/decide_mode_sweep.json 2> gpurun_out/final_d/decide_mode_sweep.err; echo "sweep rc=$?"
