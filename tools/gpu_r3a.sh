# round 2, third session: undo log of the LOCAL kernel -- bit-equality tests, then the sweep with the log on and off
mkdir -p gpurun_out/r3a
timeout 1200 python -m pytest tests/test_gpu_decision_mode.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r3a/tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3a/tests.log
export SWEEP_STEPS=2048 SWEEP_MODES=local SWEEP_ANCHORS=1,2,4,8,16,32,48
timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3a/sweep_undo.json 2> gpurun_out/r3a/sweep_undo.err; echo "sweep rc=$?"
SEEKR2_B200_LOCAL_UNDO=0 timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3a/sweep_noundo.json 2> gpurun_out/r3a/sweep_noundo.err; echo "sweep0 rc=$?"
SEEKR2_B200_TUNE=32 timeout 600 python tools/decide_mode_sweep.py > gpurun_out/r3a/sweep_lazy.json 2> gpurun_out/r3a/sweep_lazy.err; echo "sweep lazy rc=$?"
cat gpurun_out/r3a/sweep_undo.err
