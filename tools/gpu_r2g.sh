set -x
mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_decision_mode.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2g/tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2g/tests.log
SWEEP_TUNES=0 SWEEP_EXTRA=",SEEKR2_B200_SEGMENTED=0" SWEEP_HANDOFF=0 SWEEP_ANCHORS=1,2,4,8,16,32 SWEEP=quick timeout 1200 python tools/latency_sweep.py > gpurun_out/r2g/latency_sweep.jsonl 2> gpurun_out/r2g/latency_sweep.err; echo "sweep rc=$?"
cut -c1-270 gpurun_out/r2g/latency_sweep.jsonl
tail -3 gpurun_out/r2g/latency_sweep.err
for args in "1 23036 local" "1 2520 local"; do
SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128 timeout 300 python tools/handoff_probe.py $args > "gpurun_out/r2g/probe_$(echo $args | tr ' ' '_').txt" 2>&1
grep LOCAL "gpurun_out/r2g/probe_$(echo $args | tr ' ' '_').txt" | tail -4
done
