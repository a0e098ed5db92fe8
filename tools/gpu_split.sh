# FOLLOW mode, undo-log size and the register-split LOCAL kernel (setmaxnreg): bit-equality under the knob test, then us per step
# SPLIT_CFGS: list of follow:regs:undo_tiles (empty = default)
mkdir -p gpurun_out/split
if [ -z "$SPLIT_SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_decision_mode.py -x -q -k "knobs" > gpurun_out/split/knob_tests.log 2>&1; echo "knob tests rc=$?"; tail -3 gpurun_out/split/knob_tests.log
fi
export SWEEP_STEPS=2048 SWEEP_MODES=local SWEEP_KINDS=mmvt SWEEP_ANCHORS=${SPLIT_ANCHORS:-4,8,12,16,24,32,48}
for cfg in ${SPLIT_CFGS:-"0::" "1::" "1:56:" "1:64:" "1:164:"}; do
  IFS=: read follow regs undo <<< "$cfg"
  echo "== LOCAL_FOLLOW=${follow:-default} LOCAL_REGS=${regs:-default} UNDO_TILES=${undo:-default}"
  if [ -z "$follow" ]; then unset SEEKR2_B200_LOCAL_FOLLOW; else export SEEKR2_B200_LOCAL_FOLLOW=$follow; fi
  if [ -z "$regs" ]; then unset SEEKR2_B200_LOCAL_REGS; else export SEEKR2_B200_LOCAL_REGS=$regs; fi
  if [ -z "$undo" ]; then unset SEEKR2_B200_LOCAL_UNDO_TILES; else export SEEKR2_B200_LOCAL_UNDO_TILES=$undo; fi
  timeout 400 python tools/decide_mode_sweep.py 2>&1 >/dev/null | cut -c1-80 | tr '\n' ' '; echo
done 2>&1 | tee gpurun_out/split/sweep.txt
