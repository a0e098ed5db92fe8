set -x
mkdir -p gpurun_out/r2k
B="python bench.py --steps 64 --warmup 32 --no-cpu-baseline --no-e2e --no-box --no-adapter"
timeout 600 $B > gpurun_out/r2k/bench_plain.json 2> gpurun_out/r2k/bench_plain.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2k/launches.csv $B > gpurun_out/r2k/ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_step_kernel -s 40 -c 3 -f -o gpurun_out/r2k/fused_step $B > gpurun_out/r2k/ncu2.log 2>&1; echo "ncu fused_step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_local_kernel -s 300 -c 3 -f -o gpurun_out/r2k/fused_local $B > gpurun_out/r2k/ncu3.log 2>&1; echo "ncu fused_local rc=$?"
timeout 600 python tools/kernel_bench.py 128 > gpurun_out/r2k/kernel_bench_a128.txt 2> gpurun_out/r2k/kernel_bench.err; echo "kernel_bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kick_kernel|finish_kernel|resync_kernel|ring_pack_kernel|middle_part" -s 20 -c 14 -f -o gpurun_out/r2k/two_pass python tools/kernel_bench.py 128 > gpurun_out/r2k/ncu4.log 2>&1; echo "ncu two_pass rc=$?"
timeout 600 python tools/launch_gaps.py > gpurun_out/r2k/launch_gaps.json 2> gpurun_out/r2k/launch_gaps.err; echo "gaps rc=$?"
cat gpurun_out/r2k/kernel_bench_a128.txt
tail -3 gpurun_out/r2k/launch_gaps.err
ls -la gpurun_out/r2k
