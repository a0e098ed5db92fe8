mkdir -p gpurun_out/r3g
export SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128
for n in 23036 2520 64; do
  echo "== atoms $n"; python tools/handoff_probe.py 1 $n local 2>&1 | grep -v "^iter\|streaming blocks that" | tail -4
done > gpurun_out/r3g/probe.txt 2>&1
cat gpurun_out/r3g/probe.txt
