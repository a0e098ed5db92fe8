# undo log of the LOCAL kernel: bit-equality tests, then the sweep over the number of logged tiles
mkdir -p gpurun_out/r3b
timeout 1500 python -m pytest tests/test_gpu_decision_mode.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r3b/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3b/tests.log
export SWEEP_STEPS=2048 SWEEP_MODES=local SWEEP_ANCHORS=2,4,6,8,12,16,24,32,48
for u in 1 2 3 99; do
  SEEKR2_B200_LOCAL_UNDO_TILES=$u timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3b/sweep_u$u.json 2> gpurun_out/r3b/sweep_u$u.err; echo "sweep u=$u rc=$?"
done
