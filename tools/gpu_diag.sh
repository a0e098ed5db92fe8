# what separates the MMVT step from the plain one at 8 / 16 / 32 anchors: Elber (same kernel, slot role in one block per anchor, no undo log)
# and plain Langevin under the register caps of the MMVT kernel
mkdir -p gpurun_out/split
export SWEEP_STEPS=2048 SWEEP_MODES=local SWEEP_ANCHORS=8,16,32
for cfg in "elber,langevin:" "elber,langevin:72"; do
  kinds=${cfg%%:*}; regs=${cfg##*:}
  echo "== kinds=$kinds LOCAL_REGS=${regs:-default}"
  if [ -z "$regs" ]; then unset SEEKR2_B200_LOCAL_REGS; else export SEEKR2_B200_LOCAL_REGS=$regs; fi
  SWEEP_KINDS=$kinds timeout 400 python tools/decide_mode_sweep.py 2>&1 >/dev/null | cut -c1-120
done 2>&1 | tee gpurun_out/split/diag.txt
