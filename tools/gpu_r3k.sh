mkdir -p gpurun_out/r3k
timeout 1500 python -m pytest tests/test_gpu_decision_mode.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_scenarios.py tests/test_gpu_known_answers.py tests/test_gpu_robustness.py -m gpu -x -q > gpurun_out/r3k/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3k/tests.log
export SWEEP_STEPS=4096 SWEEP_MODES=auto SWEEP_ANCHORS=1,2,4,8
timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3k/sweep_auto.json 2> gpurun_out/r3k/sweep_auto.err; echo "sweep rc=$?"; cut -c1-160 gpurun_out/r3k/sweep_auto.err
timeout 600 python tools/decide_mode_sweep.py 2520 > gpurun_out/r3k/sweep_auto_2520.json 2> gpurun_out/r3k/sweep_auto_2520.err; echo "sweep rc=$?"; cut -c1-160 gpurun_out/r3k/sweep_auto_2520.err
