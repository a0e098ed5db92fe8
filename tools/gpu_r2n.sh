mkdir -p gpurun_out/r2n
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n/bench_driver_cmd.json 2> gpurun_out/r2n/bench_driver_cmd.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2n/bench_reference_arm.json 2> gpurun_out/r2n/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2n/bench_default.json 2> gpurun_out/r2n/bench_default.err; echo "bench default rc=$?"
tail -3 gpurun_out/r2n/*.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
