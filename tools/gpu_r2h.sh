mkdir -p gpurun_out/r2h
for r in 64 72 80 96; do
  lib=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200_r$r.so
  [ $r = 96 ] && lib=$PWD/seekr2_openmm_plugin_b200/libseekr2_b200.so
  SEEKR2_B200_LIBRARY=$lib SWEEP_TUNES=0 SWEEP_HANDOFF=1 SWEEP_ANCHORS=1,8,16,32 SWEEP_SYSTEMS=tryp SWEEP=quick timeout 900 python tools/latency_sweep.py > gpurun_out/r2h/sweep_r$r.jsonl 2> gpurun_out/r2h/sweep_r$r.err
  echo "== maxnreg $r"; grep best_plain gpurun_out/r2h/sweep_r$r.jsonl | cut -c1-330
done
