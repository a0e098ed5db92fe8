/*
 * oracle/ref_launch.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * Host launchers for the reference plugin's own CUDA kernels.  The kernels themselves are NOT in
 * this repository: REF_LANGEVIN_CU is a -D'ed absolute path to
 *   /root/reference/seekr2plugin/platforms/cuda/src/kernels/langevin.cu
 * which is #included below, unmodified, after oracle/ref_prelude.h supplies the typedefs OpenMM
 * would inject.  oracle/build_ref.sh builds one shared library per CudaPrecision into oracle/_ref/
 * (git-ignored; travels to the GPU box with the snapshot).
 *
 * Launch shape restated from CudaContext::executeKernel as used at CudaSeekr2Kernels.cpp:223,238,357:
 * 128-thread blocks, grid = min(ceil(numAtoms/128), numThreadBlocks), grid-stride loop inside the
 * kernel.  numThreadBlocks is passed in by the caller (OpenMM uses 4 x multiprocessor count... we
 * expose it so tests can exercise the grid-stride path).
 */
#include "ref_prelude.h"
#include REF_LANGEVIN_CU

static int grid_for(int numAtoms, int numThreadBlocks) {
    int g = (numAtoms + 127) / 128;
    if (numThreadBlocks > 0 && g > numThreadBlocks) g = numThreadBlocks;
    return g < 1 ? 1 : g;
}

extern "C" {

int ref_precision(void) {
#if defined(REF_PRECISION_DOUBLE)
    return 2;
#elif defined(REF_PRECISION_MIXED)
    return 1;
#else
    return 0;
#endif
}

int ref_part1(int numAtoms, int paddedNumAtoms, void* velm, const long long* force, void* posDelta,
              const void* params, const void* dt, const void* random, unsigned int randomIndex,
              int numThreadBlocks, void* stream) {
    integrateMmvtLangevinPart1<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, paddedNumAtoms, (mixed4*) velm, force, (mixed4*) posDelta, (const mixed*) params,
        (const mixed2*) dt, (const float4*) random, randomIndex);
    return (int) cudaGetLastError();
}

int ref_part1_elber(int numAtoms, int paddedNumAtoms, void* velm, const long long* force, void* posDelta,
                    const void* params, const void* dt, const void* random, unsigned int randomIndex,
                    int numThreadBlocks, void* stream) {
    integrateElberLangevinPart1<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, paddedNumAtoms, (mixed4*) velm, force, (mixed4*) posDelta, (const mixed*) params,
        (const mixed2*) dt, (const float4*) random, randomIndex);
    return (int) cudaGetLastError();
}

int ref_part2_mmvt(int numAtoms, void* posq, void* posqCorrection, const void* posDelta, void* velm,
                   const void* dt, void* oldPosq, void* oldVelm, int numThreadBlocks, void* stream) {
    integrateMmvtLangevinPart2<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, (real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta, (mixed4*) velm,
        (const mixed2*) dt, (real4*) oldPosq, (mixed4*) oldVelm);
    return (int) cudaGetLastError();
}

int ref_part2_elber(int numAtoms, void* posq, void* posqCorrection, const void* posDelta, void* velm,
                    const void* dt, int numThreadBlocks, void* stream) {
    integrateElberLangevinPart2<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, (real4*) posq, (real4*) posqCorrection, (const mixed4*) posDelta, (mixed4*) velm,
        (const mixed2*) dt);
    return (int) cudaGetLastError();
}

int ref_bounce(int numAtoms, int paddedNumAtoms, void* posq, void* velm, const void* oldPosq,
               const void* oldVelm, int numThreadBlocks, void* stream) {
    mmvtBounce<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, paddedNumAtoms, (real4*) posq, (mixed4*) velm, (const real4*) oldPosq,
        (const mixed4*) oldVelm);
    return (int) cudaGetLastError();
}

}  // extern "C"
