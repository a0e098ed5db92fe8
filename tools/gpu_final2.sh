# 2-GPU bench (NCCL path) + plain Langevin register rule check
mkdir -p gpurun_out/final2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final2/bench_2gpu.json 2> gpurun_out/final2/bench_2gpu.err; echo "2gpu rc=$?"
SWEEP_STEPS=2048 SWEEP_MODES=auto SWEEP_ANCHORS=8,12,16,24 timeout 600 python tools/decide_mode_sweep.py > gpurun_out/final2/sweep_plain_rule.json 2> gpurun_out/final2/sweep_plain_rule.err; cut -c1-200 gpurun_out/final2/sweep_plain_rule.err
