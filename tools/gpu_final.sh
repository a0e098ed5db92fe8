# what the driver runs at round end, plus the default bench, the decision-mode sweep and the ncu passes of the final code
mkdir -p gpurun_out/final
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/final/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_driver_cmd.json 2> gpurun_out/final/bench_driver_cmd.err; echo "bench rc=$?"
timeout 900 python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; echo "bench default rc=$?"
SWEEP_STEPS=2048 timeout 900 python tools/decide_mode_sweep.py > gpurun_out/final/decide_mode_sweep.json 2> gpurun_out/final/decide_mode_sweep.err; echo "sweep rc=$?"
B="python bench.py --steps 64 --warmup 32 --no-cpu-baseline --no-e2e --no-box --no-adapter"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final/launches.csv $B > gpurun_out/final/ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_step_kernel -s 40 -c 3 -f -o gpurun_out/final/fused_step $B > gpurun_out/final/ncu2.log 2>&1; echo "ncu fused_step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_local_kernel -s 300 -c 3 -f -o gpurun_out/final/fused_local $B > gpurun_out/final/ncu3.log 2>&1; echo "ncu fused_local rc=$?"
ls -la gpurun_out/final
