set -x
mkdir -p gpurun_out/r2b
./tools/microbench/decide_chain_latency > gpurun_out/r2b/microbench.txt 2>&1; cat gpurun_out/r2b/microbench.txt
timeout 600 python -m pytest tests/test_gpu_decision_mode.py -x -q > gpurun_out/r2b/knobs.log 2>&1; echo "knobs rc=$?"
tail -3 gpurun_out/r2b/knobs.log
SWEEP=quick timeout 900 python tools/latency_sweep.py > gpurun_out/r2b/latency_sweep.jsonl 2> gpurun_out/r2b/latency_sweep.err; echo "sweep rc=$?"
grep best_plain gpurun_out/r2b/latency_sweep.jsonl
for args in "1 23036 local" "1 2520 local"; do
SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128 timeout 300 python tools/handoff_probe.py $args > "gpurun_out/r2b/probe_$(echo $args | tr ' ' '_').txt" 2>&1
SEEKR2_B200_LOCAL_THREADS=128 SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128 timeout 300 python tools/handoff_probe.py $args > "gpurun_out/r2b/probe_T128_$(echo $args | tr ' ' '_').txt" 2>&1
done
grep LOCAL gpurun_out/r2b/probe_1_23036_local.txt | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2b/gpu_tests.log
