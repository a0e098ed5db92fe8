mkdir -p gpurun_out/r2w
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2w/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2w/gpu_tests.log
SWEEP_MODES=handoff SWEEP_STEPS=1024 SWEEP_ANCHORS=64,96,128,256 timeout 900 python tools/decide_mode_sweep.py > gpurun_out/r2w/sweep.json 2> gpurun_out/r2w/sweep.err
python -c "
import json
d=json.load(open('gpurun_out/r2w/sweep.json'))
for r in d['rows']: print(r['anchors'], 'mmvt_handoff', r['mmvt_handoff_us'], 'plain_handoff', r['langevin_handoff_us'], 'ratio', round(r['mmvt_handoff_us']/r['langevin_handoff_us'],4))
"
SEEKR2_B200_LIBRARY=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_probe.so PROBE_GRAPH=1 PROBE_STEPS=64 timeout 300 python tools/handoff_probe.py 128 23036 handoff 2>&1 | tail -3 | cut -c1-330
timeout 600 python tools/kernel_bench.py 128 2>/dev/null | grep -i "finish\|kick\|MMVT,LANGEVIN"
