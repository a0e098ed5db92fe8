#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference plugin's OWN CUDA kernels (langevin.cu and langevinMiddle.cu), unmodified and from where they
# lie under /root/reference, into oracle/_ref/libref_langevin[_middle]_{single,mixed,double}.so for sm_100a.
# No reference source is copied into this repository; oracle/_ref/ is git-ignored and travels to the
# GPU box with the gpurun snapshot.  The reference's own CMake build is not used (it needs OpenMM).
# A second set built with -use_fast_math (what OpenMM passes to NVRTC) is produced for the tolerance
# test of the SQRT / rounding differences.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${SEEKR2_REFERENCE_DIR:-/root/reference}/seekr2plugin/platforms/cuda/src/kernels/langevin.cu"
refm="${SEEKR2_REFERENCE_DIR:-/root/reference}/seekr2plugin/platforms/cuda/src/kernels/langevinMiddle.cu"
if [ ! -f "$ref" ]; then
    echo "build_ref: $ref not present (expected on the GPU box) -- keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$here/_ref"
arch="-gencode arch=compute_100a,code=sm_100a"
for p in single mixed double; do
    P=$(echo "$p" | tr a-z A-Z)
    nvcc $arch -O3 -lineinfo -shared -Xcompiler -fPIC -DREF_PRECISION_$P \
        -DREF_LANGEVIN_CU="\"$ref\"" -I"$here" -o "$here/_ref/libref_langevin_$p.so" "$here/ref_launch.cu"
    nvcc $arch -O3 -lineinfo -use_fast_math -shared -Xcompiler -fPIC -DREF_PRECISION_$P \
        -DREF_LANGEVIN_CU="\"$ref\"" -I"$here" -o "$here/_ref/libref_langevin_${p}_fastmath.so" "$here/ref_launch.cu"
    nvcc $arch -O3 -lineinfo -shared -Xcompiler -fPIC -DREF_PRECISION_$P \
        -DREF_LANGEVIN_MIDDLE_CU="\"$refm\"" -I"$here" -o "$here/_ref/libref_langevin_middle_$p.so" "$here/ref_launch_middle.cu"
done
echo "build_ref: built $(ls "$here/_ref" | wc -l) libraries in $here/_ref"
