mkdir -p gpurun_out/r2m
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2m/bench_2gpu.json 2> gpurun_out/r2m/bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2m/bench_2gpu.err
nvidia-smi topo -m > gpurun_out/r2m/topo.txt 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r2m/bench_2gpu.json'))
print(d['value'], d['n_gpus'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['pcie_ceiling_gbs'], d['e2e']['frac_of_pcie'], d['e2e']['staging_cpus'], d['gathered_anchors'])
"
