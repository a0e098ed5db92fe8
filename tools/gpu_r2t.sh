mkdir -p gpurun_out/r2t
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2t/bench_8gpu.json 2> gpurun_out/r2t/bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r2t/bench_8gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2t/bench_8gpu.json'))
print(d['value'], d['n_gpus'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['frac_of_pcie'], d['e2e']['pcie_ceiling_gbs']['alone_per_gpu'], d['e2e']['pcie_ceiling_gbs']['all_ranks_at_once_per_gpu'], d['e2e']['staging_cpus'], d['gathered_anchors'])
"
