mkdir -p gpurun_out/r3f
export SWEEP_STEPS=1024 SWEEP_MODES=local SWEEP_ANCHORS=8,12,16,24,32,48
for regs in 72 56 48; do for thr in 128 256; do
  tag=r${regs}_t${thr}
  SEEKR2_B200_LOCAL_REGS=$regs SEEKR2_B200_LOCAL_THREADS=$thr timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3f/sweep_$tag.json 2> gpurun_out/r3f/sweep_$tag.err; echo "$tag rc=$?"; cut -c1-110 gpurun_out/r3f/sweep_$tag.err
done; done
