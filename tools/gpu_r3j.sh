mkdir -p gpurun_out/r3j
export SWEEP_STEPS=4096 SWEEP_MODES=local SWEEP_ANCHORS=1,2,3,4,6,8,12
for t in 96 128 160 192 224 256; do
SEEKR2_B200_LOCAL_THREADS=$t timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3j/sweep_t$t.json 2> gpurun_out/r3j/sweep_t$t.err; echo "T=$t rc=$?"; cut -c1-100 gpurun_out/r3j/sweep_t$t.err
done
