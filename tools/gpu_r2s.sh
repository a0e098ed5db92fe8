mkdir -p gpurun_out/r2s
SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=64 timeout 300 python tools/handoff_probe.py 128 23036 handoff > gpurun_out/r2s/probe_128_handoff.txt 2>&1
tail -12 gpurun_out/r2s/probe_128_handoff.txt | cut -c1-400
