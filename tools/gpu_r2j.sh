set -x
mkdir -p gpurun_out/r2j
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2j/gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r2j/gpu_tests.log
SWEEP_STEPS=2048 timeout 1500 python tools/decide_mode_sweep.py > gpurun_out/r2j/decide_mode_sweep.json 2> gpurun_out/r2j/decide_mode_sweep.err; echo "sweep rc=$?"
cut -c1-400 gpurun_out/r2j/decide_mode_sweep.err | tail -16
