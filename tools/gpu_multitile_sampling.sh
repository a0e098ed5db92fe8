# warp-state sampling (finest interval, up to 50 passes) of the LOCAL kernel in the multi-tile regime: 16 anchors x 23 036 atoms, MMVT and Elber
mkdir -p gpurun_out/mts
export SWEEP_STEPS=64 SWEEP_MODES=local SWEEP_ANCHORS=16
for kind in mmvt elber; do
  SWEEP_KINDS=$kind timeout 600 ncu --section SourceCounters --section WarpStateStats --warp-sampling-interval 0 --warp-sampling-max-passes 50 --clock-control none --import-source on -k regex:fused_local_kernel -s 150 -c 2 -f -o gpurun_out/mts/sampled_a16_$kind python tools/decide_mode_sweep.py > gpurun_out/mts/ncu_$kind.log 2>&1; echo "$kind rc=$?"
done
ls -la gpurun_out/mts
