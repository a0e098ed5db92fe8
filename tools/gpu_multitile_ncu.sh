# ncu --set full of the LOCAL kernel in the multi-tile regime (16 anchors x 23 036 atoms): MMVT against Elber (same kernel, slot role in
# one block per anchor, no undo log) and plain Langevin
mkdir -p gpurun_out/mt
export SWEEP_STEPS=64 SWEEP_MODES=local SWEEP_ANCHORS=16
for kind in mmvt elber langevin; do
  SWEEP_KINDS=$kind timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_local_kernel -s 150 -c 2 -f -o gpurun_out/mt/local_a16_$kind python tools/decide_mode_sweep.py > gpurun_out/mt/ncu_$kind.log 2>&1; echo "$kind rc=$?"
done
ls -la gpurun_out/mt
