mkdir -p gpurun_out/r2i
for r in 72 96; do
  lib=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_r$r.so
  [ $r = 96 ] && lib=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200.so
  SEEKR2_B200_LIBRARY=$lib SWEEP_STEPS=1024 SWEEP_TUNES=0 SWEEP_HANDOFF=1 SWEEP_ANCHORS=6,12,24,48,64,96,128 SWEEP_SYSTEMS=tryp SWEEP=quick timeout 1200 python tools/latency_sweep.py > gpurun_out/r2i/sweep_r$r.jsonl 2> gpurun_out/r2i/sweep_r$r.err
  echo "== maxnreg $r"; grep -v best_plain gpurun_out/r2i/sweep_r$r.jsonl | cut -c30-200
done
