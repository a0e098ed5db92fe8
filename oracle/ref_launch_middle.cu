/*
 * oracle/ref_launch_middle.cu -- TEST INFRASTRUCTURE ONLY.
 * Host launchers for the reference plugin's own LangevinMiddle CUDA kernels
 * (/root/reference/seekr2plugin/platforms/cuda/src/kernels/langevinMiddle.cu, #included unmodified via
 * -DREF_LANGEVIN_MIDDLE_CU after oracle/ref_prelude.h).  Launch shape as in ref_launch.cu
 * (CudaSeekr2Kernels.cpp:767,783,798,907).
 */
#include "ref_prelude.h"
#include REF_LANGEVIN_MIDDLE_CU

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

int ref_middle_part1(int numAtoms, int paddedNumAtoms, void* velm, const long long* force, const void* dt,
                     int numThreadBlocks, void* stream) {
    integrateMmvtLangevinMiddlePart1<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, paddedNumAtoms, (mixed4*) velm, force, (const mixed2*) dt);
    return (int) cudaGetLastError();
}

int ref_middle_part2(int numAtoms, void* velm, void* posDelta, void* oldDelta, const void* params, const void* dt,
                     void* oldVelm, const void* random, unsigned int randomIndex, int numThreadBlocks, void* stream) {
    integrateMmvtLangevinMiddlePart2<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, (mixed4*) velm, (mixed4*) posDelta, (mixed4*) oldDelta, (const mixed*) params, (const mixed2*) dt,
        (mixed4*) oldVelm, (const float4*) random, randomIndex);
    return (int) cudaGetLastError();
}

int ref_middle_part3(int numAtoms, void* posq, void* velm, const void* posDelta, void* oldDelta, const void* dt,
                     void* oldPosq, void* posqCorrection, int numThreadBlocks, void* stream) {
    integrateMmvtLangevinMiddlePart3<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, (real4*) posq, (mixed4*) velm, (const mixed4*) posDelta, (mixed4*) oldDelta, (const mixed2*) dt,
        (real4*) oldPosq, (real4*) posqCorrection);
    return (int) cudaGetLastError();
}

int ref_middle_bounce(int numAtoms, int paddedNumAtoms, void* posq, void* velm, const void* oldPosq,
                      const void* oldVelm, int numThreadBlocks, void* stream) {
    mmvtBounce<<<grid_for(numAtoms, numThreadBlocks), 128, 0, (cudaStream_t) stream>>>(
        numAtoms, paddedNumAtoms, (real4*) posq, (mixed4*) velm, (const real4*) oldPosq, (const mixed4*) oldVelm);
    return (int) cudaGetLastError();
}

}  // extern "C"
