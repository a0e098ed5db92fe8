# 8-GPU bench of the final code (weak scaling, NCCL gather at the end)
mkdir -p gpurun_out/final8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/final8/bench_8gpu.json 2> gpurun_out/final8/bench_8gpu.err; echo "8gpu rc=$?"
tail -c 600 gpurun_out/final8/bench_8gpu.json
