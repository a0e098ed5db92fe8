# A/B of experimental library builds (SEEKR2_B200_LIBRARY) within one box: LOCAL kernel, 1 and 2 anchors x 23 036 atoms, and the bench's latency legs
mkdir -p gpurun_out/ab
export SWEEP_STEPS=8192 SWEEP_MODES=auto SWEEP_ANCHORS=1,2
for lib in "$@"; do
  echo "== $lib"
  SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/$lib timeout 600 python tools/decide_mode_sweep.py 2>&1 >/dev/null | cut -c1-140
  SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/$lib python bench.py --steps 64 --warmup 32 --no-cpu-baseline --no-e2e --no-box --no-adapter 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); hg=d['host_guest_c2']; sa=d['single_anchor_per_gpu']
print('   HG mmvt %.3f elber %.3f plain %.3f | 1-anchor mmvt %.3f plain %.3f | floor mmvt %.3f plain %.3f' % (hg['mmvt_us_per_step'],hg['elber_us_per_step'],hg['plain_langevin_us_per_step'],sa['us_per_step'],sa['plain_langevin_us_per_step'],sa['launch_overhead_us_per_step']['mmvt'],sa['launch_overhead_us_per_step']['plain_langevin']))"
done 2>&1 | tee gpurun_out/ab/result.txt
