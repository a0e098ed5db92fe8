mkdir -p gpurun_out/r3l
export SWEEP_STEPS=8192 SWEEP_MODES=auto SWEEP_ANCHORS=1,2
for rep in 1 2 3; do for tune in 0 64; do
SEEKR2_B200_TUNE=$tune timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3l/sweep_tune${tune}_$rep.json 2> gpurun_out/r3l/sweep_tune${tune}_$rep.err; echo "tune=$tune rep=$rep"; cut -c1-130 gpurun_out/r3l/sweep_tune${tune}_$rep.err
done; done
