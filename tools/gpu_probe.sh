# LOCAL kernel chain stamps (probe build): one anchor of 23 036 atoms (packed slots), the host-guest system (aligned slots), 64 atoms
mkdir -p gpurun_out/probe
export SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=128
{ echo "== 1 x 23036 atoms, 11 + 9 atom groups"; python tools/handoff_probe.py 1 23036 local 2>&1 | grep -v "^iter\|streaming blocks that" | tail -2
  echo "== host-guest: 1 x 2520 atoms, 147 + 15 atom groups"; PROBE_HG=1 python tools/handoff_probe.py 1 2520 local 2>&1 | grep -v "^iter\|streaming blocks that" | tail -2
  echo "== 1 x 64 atoms"; python tools/handoff_probe.py 1 64 local 2>&1 | grep -v "^iter\|streaming blocks that" | tail -2; } > gpurun_out/probe/probe.txt 2>&1
cat gpurun_out/probe/probe.txt
