mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_openmm_shim.py -x -q -k "elber" > gpurun_out/r2o/shim_elber.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2o/shim_elber.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-box > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o/bench.json'))
a=d["adapter_vs_reference_cuda_platform"]
for size in ("c2","c4"):
    for kind in ("mmvt","elber"):
        e=a[size][kind]
        print(size,kind,{k:(round(v,2) if isinstance(v,float) else v) for k,v in e.items()})
PY
