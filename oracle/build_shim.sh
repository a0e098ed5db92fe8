#!/usr/bin/env bash
# oracle/build_shim.sh -- TEST INFRASTRUCTURE ONLY.
# Builds, under oracle/_ref/shim/, everything that runs the plugin through tests/openmm_shim (a minimal stand-in
# for the OpenMM API, NOT OpenMM):
#   libopenmm_shim.so        the stand-in itself (tests/openmm_shim/src)
#   libseekr2_api.so         the reference's OWN integrator classes, unmodified, compiled from where they lie
#                            (seekr2plugin/openmmapi/src/*.cpp)
#   libref_cuda_platform.so  the reference's OWN CUDA-platform host code (CudaSeekr2Kernels.cpp, CudaSeekr2KernelFactory.cpp)
#                            plus its kernel sources embedded as strings (what the reference's CMake step
#                            EncodeCUDAFiles.cmake generates); the kernels go through NVRTC at run time, as in OpenMM
#   libb200_platform.so      this repository's platforms/b200 adapter over libseekr2_b200.so
#   ref_Test*, b200_Test*    the reference's OWN platform tests (platforms/cuda/tests/Test*.cpp, unmodified) linked
#                            against either platform library
#   ref_scenario, b200_scenario   tests/openmm_shim/src/scenario_driver.cpp against either platform library
# No reference source is copied into the repository: oracle/_ref/ is git-ignored and travels to the GPU box.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(dirname "$here")"
refroot="${SEEKR2_REFERENCE_DIR:-/root/reference}/seekr2plugin"
if [ ! -d "$refroot/openmmapi" ]; then
    echo "build_shim: $refroot not present (expected on the GPU box) -- keeping prebuilt oracle/_ref/shim" >&2
    exit 0
fi
out="$here/_ref/shim"
mkdir -p "$out/gen"
shim="$root/tests/openmm_shim"
cuda="${CUDA_HOME:-/usr/local/cuda}"
CXX="${CXX:-g++}"
flags="-O1 -g -std=c++17 -fPIC -w -I$shim/include -I$cuda/include -I$refroot/openmmapi/include"
rpath="-Wl,-rpath,\$ORIGIN -Wl,-rpath,\$ORIGIN/../../../seekr2_openmm_plugin_b200 -Wl,-rpath,$cuda/lib64"
cudalibs="-L$cuda/lib64 -L$cuda/lib64/stubs -lcudart -lnvrtc -lcuda"

# 1. the stand-in
$CXX $flags -shared -o "$out/libopenmm_shim.so" "$shim/src/shim_core.cpp" "$shim/src/shim_cuda.cpp" $cudalibs $rpath

# 2. the reference's integrator classes
$CXX $flags -shared -o "$out/libseekr2_api.so" "$refroot"/openmmapi/src/*.cpp -L"$out" -lopenmm_shim $rpath

# 3. the reference's CUDA platform; kernel sources embedded the way EncodeCUDAFiles.cmake does
python3 - "$refroot/platforms/cuda/src/kernels" "$out/gen" <<'PY'
import os, sys
src, gen = sys.argv[1], sys.argv[2]
names = sorted(f[:-3] for f in os.listdir(src) if f.endswith(".cu"))
with open(os.path.join(gen, "CudaSeekr2KernelSources.h"), "w") as h:
    h.write("#pragma once\n#include <string>\nnamespace Seekr2Plugin {\nclass CudaSeekr2KernelSources {\npublic:\n")
    for n in names:
        h.write(f"    static const std::string {n};\n")
    h.write("};\n}\n")
with open(os.path.join(gen, "CudaSeekr2KernelSources.cpp"), "w") as c:
    c.write('#include "CudaSeekr2KernelSources.h"\nusing namespace Seekr2Plugin;\n')
    for n in names:
        text = open(os.path.join(src, n + ".cu")).read()
        assert ')SEEKR2SRC"' not in text
        c.write(f'const std::string CudaSeekr2KernelSources::{n} = R"SEEKR2SRC({text})SEEKR2SRC";\n')
PY
$CXX $flags -I"$out/gen" -I"$refroot/platforms/cuda/include" -I"$refroot/platforms/cuda/src" -shared -o "$out/libref_cuda_platform.so" \
    "$refroot/platforms/cuda/src/CudaSeekr2Kernels.cpp" "$refroot/platforms/cuda/src/CudaSeekr2KernelFactory.cpp" \
    "$out/gen/CudaSeekr2KernelSources.cpp" -L"$out" -lseekr2_api -lopenmm_shim $cudalibs $rpath

# 4. this repository's adapter
$CXX $flags -I"$root/include" -I"$root/platforms/b200/include" -shared -o "$out/libb200_platform.so" \
    "$root"/platforms/b200/src/*.cpp -L"$out" -lseekr2_api -lopenmm_shim \
    -L"$root/seekr2_openmm_plugin_b200" -lseekr2_b200 $cudalibs $rpath

# 5. the reference's own platform tests and the scenario driver, once per platform library
for platform in ref_cuda b200; do
    prefix="${platform%_cuda}"
    for t in "$refroot"/platforms/cuda/tests/Test*.cpp; do
        name="$(basename "$t" .cpp)"
        $CXX $flags -o "$out/${prefix}_${name}" "$t" -L"$out" -l${platform}_platform -lseekr2_api -lopenmm_shim $cudalibs $rpath
    done
    $CXX $flags -I"$root/include" -o "$out/${prefix}_scenario" "$shim/src/scenario_driver.cpp" -L"$out" -l${platform}_platform -lseekr2_api -lopenmm_shim $cudalibs $rpath
done
echo "build_shim: built $(ls "$out" | grep -vc '^gen$') files in $out"
