mkdir -p gpurun_out/r3i
timeout 1500 python -m pytest tests/test_gpu_decision_mode.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py tests/test_gpu_scenarios.py tests/test_gpu_known_answers.py -m gpu -x -q > gpurun_out/r3i/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3i/tests.log
export SWEEP_STEPS=4096 SWEEP_MODES=local SWEEP_ANCHORS=1,2
for t in 128 160 192 256; do
SEEKR2_B200_LOCAL_THREADS=$t timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3i/sweep_t$t.json 2> gpurun_out/r3i/sweep_t$t.err; echo "T=$t rc=$?"; cut -c1-120 gpurun_out/r3i/sweep_t$t.err
done
export SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128
python tools/handoff_probe.py 1 23036 local 2>&1 | grep -v "^iter\|streaming blocks that" | tail -2 | tee gpurun_out/r3i/probe.txt
