mkdir -p gpurun_out/r2q
timeout 600 python -m pytest tests/test_gpu_cv_float.py -x -q > gpurun_out/r2q/cvfloat.log 2>&1; echo "cvfloat rc=$?"; tail -5 gpurun_out/r2q/cvfloat.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2q/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2q/gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-adapter > gpurun_out/r2q/bench.json 2> gpurun_out/r2q/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2q/bench.json'))
print(d['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['mmvt_over_plain_langevin'], d['long_run']['roofline_frac'], d['single_anchor_per_gpu']['us_per_step'], d['box_100k']['frac_of_hbm_peak'], d['e2e']['frac_of_pcie'])
"
