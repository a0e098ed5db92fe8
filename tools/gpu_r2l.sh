set -x
mkdir -p gpurun_out/r2l
for k in finish_kernel middle_finish_kernel resync_kernel ring_pack_kernel middle_part2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k|::$k" -s 6 -c 2 -f -o gpurun_out/r2l/$k python tools/kernel_bench.py 128 > gpurun_out/r2l/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/r2l
