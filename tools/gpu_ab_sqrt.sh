# A/B in one box: product library against an experimental build with the call-free square root in sqrt_inv_mass() too
# (make -C seekr2_openmm_plugin_b200/csrc exp2); one, two and 128 anchors x 23 036 atoms
mkdir -p gpurun_out/abs
export SWEEP_STEPS=2048 SWEEP_MODES=auto SWEEP_KINDS=mmvt,langevin SWEEP_ANCHORS=1,2,128
for lib in libseekr2_b200.so libseekr2_b200_exp.so; do
  echo "== $lib"
  SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/$lib timeout 60 python tools/decide_mode_sweep.py 2>&1 >/dev/null | cut -c1-140
done 2>&1 | tee gpurun_out/abs/result.txt
