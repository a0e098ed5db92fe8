# warp-state sampling of the LOCAL kernel (one anchor x 23 036 atoms): which warps wait for what
mkdir -p gpurun_out/r2p
B="python bench.py --steps 64 --warmup 32 --no-cpu-baseline --no-e2e --no-box --no-adapter"
timeout 900 ncu --section SourceCounters --section WarpStateStats --warp-sampling-interval 0 --warp-sampling-max-passes 50 --clock-control none --import-source on -k regex:fused_local_kernel -s 300 -c 4 -f -o gpurun_out/r2p/fused_local_sampled $B > gpurun_out/r2p/ncu.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r2p/ncu.log
ls -la gpurun_out/r2p
