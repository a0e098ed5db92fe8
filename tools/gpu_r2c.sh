set -x
mkdir -p gpurun_out/r2c
for tune in 0 8 16; do
for args in "1 23036 local" "1 2520 local"; do
SEEKR2_B200_TUNE=$tune SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128 timeout 300 python tools/handoff_probe.py $args > "gpurun_out/r2c/probe_t${tune}_$(echo $args | tr ' ' '_').txt" 2>&1
grep LOCAL "gpurun_out/r2c/probe_t${tune}_$(echo $args | tr ' ' '_').txt" | tail -2
done; done
SWEEP_TUNES=0,8,16,24 SWEEP_HANDOFF=0 SWEEP_ANCHORS=1,4 SWEEP=quick timeout 900 python tools/latency_sweep.py > gpurun_out/r2c/latency_sweep.jsonl 2> gpurun_out/r2c/latency_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r2c/latency_sweep.jsonl | cut -c1-260
