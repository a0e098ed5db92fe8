# A/B in one box: product library against an experimental build in which EVERY kind of the LOCAL kernel generates the normals of later
# tiles behind the tile's loads (-DSEEKR2_NORMALS_AFTER_LOADS_ALL; build it here first: make -C seekr2_openmm_plugin_b200/csrc exp); plain
# Langevin and Elber, multi-tile regime
mkdir -p gpurun_out/abn
export SWEEP_STEPS=2048 SWEEP_MODES=auto SWEEP_KINDS=langevin,elber SWEEP_ANCHORS=8,16,32,48
for lib in libseekr2_b200.so libseekr2_b200_exp.so; do
  echo "== $lib"
  SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/$lib timeout 120 python tools/decide_mode_sweep.py 2>&1 >/dev/null | cut -c1-120
done 2>&1 | tee gpurun_out/abn/result.txt
