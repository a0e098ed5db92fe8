set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_decision_mode.py -x -q > gpurun_out/r2a/knobs.log 2>&1; echo "knobs rc=$?"
tail -5 gpurun_out/r2a/knobs.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2a/gpu_tests.log
SWEEP=quick timeout 900 python tools/latency_sweep.py > gpurun_out/r2a/latency_sweep.jsonl 2> gpurun_out/r2a/latency_sweep.err; echo "sweep rc=$?"
grep best_plain gpurun_out/r2a/latency_sweep.jsonl
for args in "1 23036 local" "1 2520 local"; do
SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128 timeout 300 python tools/handoff_probe.py $args > "gpurun_out/r2a/probe_$(echo $args | tr ' ' '_').txt" 2>&1
SEEKR2_B200_TUNE=3 SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128 timeout 300 python tools/handoff_probe.py $args > "gpurun_out/r2a/probe_t3_$(echo $args | tr ' ' '_').txt" 2>&1
done
tail -4 gpurun_out/r2a/probe_1_23036_local.txt
