// cv.cuh -- milestone boundary evaluation and the crossing bookkeeping that the reference does on
// the host after a blocking energy readback (CudaSeekr2Kernels.cpp:248-347 and :517-590), done here
// on the device inside the step kernel.
//
// Boundary-value arithmetic contract (shared with the CPU oracle, oracle/seekr2_oracle.c):
//   P      = (double) posq                      (what OpenMM's force kernels see)
//   acc_g  = sum_i llrint((w_i * P_i) * 2^40)   int64 fixed point => independent of reduction order
//   C_g    = (double) acc_g * 2^-40
//   SPHERE_PAIR / SPHERE_FIXED: u = fma(dz,dz, fma(dy,dy, dx*dx)) - r*r ;  PLANE: u = C[axis] - thr
//   PAIR_SUM: u = sum of squared pair distances - r*r ;  ANGLE: u = cos(ref)*sqrt(|a|^2 |b|^2) - a.b
//   term = (k*u >= 0 ? bitcode : 0) (STEP) or k*u (RAW); value = in-order double sum of the terms
#pragma once
#include "arith.cuh"
#include "../../include/seekr2_b200.h"

namespace s2 {

using Boundary = seekr2_b200_boundary;
using AnchorState = seekr2_b200_anchor_state;
using Record = seekr2_b200_record;

constexpr double kFixedScale = 1099511627776.0;           // 2^40
constexpr double kFixedInvScale = 1.0 / 1099511627776.0;  // 2^-40
constexpr unsigned kFullWarp = 0xffffffffu;

// everything a step kernel needs besides the caller's buffers; passed by value
struct EngineDev {
    int num_atoms, padded_num_atoms, num_anchors;
    int num_slots, num_groups, num_boundaries, num_milestone_groups;
    int variant, flags_bits;
    int num_src, num_dest, end_on_src;
    int src_groups[SEEKR2_B200_MAX_ELBER];
    int dest_groups[SEEKR2_B200_MAX_ELBER];
    int ring_capacity;
    const int* slot_atom;        // [num_slots] atom index of each group slot (slots sorted by group)
    const int* slot_group;       // [num_slots]
    const double* slot_weight;   // [num_slots] normalised weights
    const Boundary* boundaries;  // [A][num_boundaries]
    AnchorState* state;          // [A]
    Record* ring;                // [A][ring_capacity]
    void* slab;                  // [2][A][num_slots] SlabEntry<P>
    int* decision;               // [A] bounce flag of the two-pass path
    long long* step_shadow;      // [4][A] step index of the fused launch with that phase (launch count mod 4): launch k reads
                                 // phase k, writes phase k+2 -- stable during the launch and already final when launch k starts early (PDL)
    unsigned long long rng_seed; // in-kernel Philox key (used when the caller passes no random pool)
    unsigned rng_anchor_base;    // Philox stream of local anchor a = rng_anchor_base + a * rng_anchor_stride: the GLOBAL anchor id, whatever
    unsigned rng_anchor_stride;  // way anchors are dealt to ranks (round-robin: base = rank, stride = world size; blocks: base = first, stride 1)
    long long* probe;            // optional timing probe (debug, seekr2_b200_debug_set_probe)
    long long* hint;             // mapped pinned host memory [2][A]: ring head, clock resets -- written when they change (seekr2_b200_poll)
    int* head_start;             // [2] decide blocks that have issued their gathers (per launch parity)
    int first_wave_blocks;       // streaming blocks resident at kernel start (they wait for head_start)
    int head_start_target;       // A x warps per decide block that hold group slots
    int prefetch_ahead;          // first-wave blocks prefetch the tile of block (own + this) into L2 while they wait (0 = off)
    int* flags;                  // [2][A] decision hand-off of the fused kernel: 0 = not yet; bit0 valid, bit1 bounce,
                                 // bit2 save the advanced state into snapshot slot (flag >> 4)
    int mmvt_any_term;           // 1: every force-group-1 term is STEP style with a bitcode in [1, 2^31], so the MMVT value
                                 // is positive exactly when one of them is active (LOCAL mode skips the in-order sum)
    int snap_slots;              // save_state_slots (0 = off)
    void* snap_posq;             // real4 [A][slots][Np]   pre-bounce positions of the last crossings
    void* snap_corr;             // real4 [A][slots][Np]   (MIXED only)
    void* snap_velm;             // mixed4[A][slots][Np]
    // LOCAL decision mode (fused_local_kernel, finish_local_kernel)
    int slot_layout;             // kLayoutAligned: every group padded to whole warps; kLayoutPacked: all groups in one warp
    int slot_threads;            // slot threads per block: 32 (packed, or no slots at all) or num_slots (aligned, a multiple of 32)
    int local_stream_threads;    // streaming threads per block (a multiple of 32); one atom per thread and tile
    int local_tiles;             // ceil(num_atoms / local_stream_threads): a block walks tiles blockIdx.x, + gridDim.x, ...
    int local_undo_tiles;        // MMVT: tiles per block whose pre-step position and bounce velocity are logged in shared memory, so that
                                 // the streaming warps store optimistically and never wait for the decision (0 = off; sized at the first launch)
    int group_warp_begin[SEEKR2_B200_MAX_GROUPS + 1];   // aligned layout: slot warps [begin[g], begin[g+1]) hold group g
    int has_rare_terms;          // a PAIR_SUM / ANGLE boundary exists (their sums go through shared memory)
    int pair_fast;               // LOCAL mode and every boundary of every anchor is a SPHERE_PAIR on the same two groups
    int pair_ga, pair_gb;        // (pair_ga, pair_gb): the LOCAL kernel reduces the DIFFERENCE of the two centroid sums directly
    int group_slot_begin[SEEKR2_B200_MAX_GROUPS];   // first slot of group g and the number of its REAL slots (padding excluded):
    int group_slot_count[SEEKR2_B200_MAX_GROUPS];   // FLAG_CV_FLOAT sums a group's float terms sequentially in this order
    unsigned spin_limit;         // polls (64 ns apart) after which a block waiting for an in-launch decision traps (0xffffffff = never)
    int late_step;               // 1: this launch follows seekr2_b200_resync under programmatic dependent launch -- the step-index shadow is
                                 // read AFTER griddepcontrol.wait (the resync kernel has just written it); only with a Gaussian pool
    int tune;                    // A/B knobs (SEEKR2_B200_TUNE): bit 5 the LOCAL kernel's slot warp joins the undo barrier on bounce steps only
};
constexpr int kLayoutAligned = 0, kLayoutPacked = 1;
constexpr int kTuneLazyUndoSync = 32;
constexpr int kMaxPackedGroups = 4;

constexpr int kCvRound = 256;   // slots handled per round by a decide block (= its thread count)

// shared-memory scratch of the CV phase
struct CvShared {
    long long acc[SEEKR2_B200_MAX_GROUPS * 3];
    float facc[SEEKR2_B200_MAX_GROUPS * 3];        // FLAG_CV_FLOAT: float group centres ...
    float fprod[3][kCvRound];                      // ... and this round's weighted float positions, by slot within the round
    double term[SEEKR2_B200_MAX_BOUNDARIES];
    int force_group[SEEKR2_B200_MAX_BOUNDARIES];
    int bounce;
};

__device__ __forceinline__ void cv_clear(CvShared& sh, int num_groups) {
    if (threadIdx.x < num_groups * 3) { sh.acc[threadIdx.x] = 0; sh.facc[threadIdx.x] = 0.0f; }
}

// ---- FLAG_CV_FLOAT: the boundary value in single precision, restating OpenMM's CUDA CustomCentroidBondForce evaluation
// (oracle: group_value_float) -- float weights, float centres summed SEQUENTIALLY in group order, sqrtf, float d^2 - r^2.
// One round = the kCvRound slots [r0, r0 + kCvRound) held one per thread: every thread deposits its weighted position,
// the block synchronises, thread g adds the real slots of group g that fall into the round, in slot order.
__device__ __forceinline__ void cvf_round(CvShared& sh, const EngineDev& e, int r0, bool valid, double w, float x, float y, float z) {
    const float wf = (float) w;
    sh.fprod[0][threadIdx.x] = valid ? __fmul_rn(wf, x) : 0.0f;
    sh.fprod[1][threadIdx.x] = valid ? __fmul_rn(wf, y) : 0.0f;
    sh.fprod[2][threadIdx.x] = valid ? __fmul_rn(wf, z) : 0.0f;
    __syncthreads();
    const int g = threadIdx.x;
    if (g < e.num_groups) {
        const int lo = max(e.group_slot_begin[g], r0), hi = min(e.group_slot_begin[g] + e.group_slot_count[g], r0 + kCvRound);
        float a0 = sh.facc[g * 3], a1 = sh.facc[g * 3 + 1], a2 = sh.facc[g * 3 + 2];
        for (int s = lo; s < hi; s++) {
            a0 = __fadd_rn(a0, sh.fprod[0][s - r0]); a1 = __fadd_rn(a1, sh.fprod[1][s - r0]); a2 = __fadd_rn(a2, sh.fprod[2][s - r0]);
        }
        sh.facc[g * 3] = a0; sh.facc[g * 3 + 1] = a1; sh.facc[g * 3 + 2] = a2;
    }
    __syncthreads();
}
__device__ __forceinline__ double cvf_term(const CvShared& sh, const Boundary& b) {
    const float* ca = sh.facc + 3 * b.groups[0];
    float u;
    if (b.type == SEEKR2_B200_CV_PLANE) {
        u = __fsub_rn(ca[b.axis], (float) b.threshold);
    } else {
        const bool pair = b.type == SEEKR2_B200_CV_SPHERE_PAIR;
        const float* cb = sh.facc + 3 * b.groups[1];
        const float d0 = __fsub_rn(ca[0], pair ? cb[0] : (float) b.center[0]);
        const float d1 = __fsub_rn(ca[1], pair ? cb[1] : (float) b.center[1]);
        const float d2 = __fsub_rn(ca[2], pair ? cb[2] : (float) b.center[2]);
        const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
        const float dist = __fsqrt_rn(r2);
        const float r = (float) b.threshold;
        u = __fsub_rn(__fmul_rn(dist, dist), __fmul_rn(r, r));
    }
    const float ku = __fmul_rn((float) b.k, u);
    if (b.style == SEEKR2_B200_STYLE_RAW) return (double) ku;
    return ku >= 0.0f ? b.bitcode : 0.0;
}

// sum over the 32 lanes of a converged warp, modulo 2^64 (exact whenever the true sum fits in int64): three limbs
// of 21 / 21 / 22 (signed) bits, each limb sum fits in 27 bits -> three independent REDUX.SUM
__device__ __forceinline__ long long warp_sum_s64(long long q) {
    const unsigned s0 = __reduce_add_sync(kFullWarp, (unsigned) ((unsigned long long) q & 0x1FFFFFu));
    const unsigned s1 = __reduce_add_sync(kFullWarp, (unsigned) (((unsigned long long) q >> 21) & 0x1FFFFFu));
    const int s2 = __reduce_add_sync(kFullWarp, (int) (q >> 42));                   // signed top limb
    return (long long) (((unsigned long long) (long long) s2 << 42) + ((unsigned long long) s1 << 21) + (unsigned long long) s0);
}

// Adds one slot's weighted position to its group's fixed-point accumulator.  Must be called by all
// 32 lanes of a warp (invalid lanes pass valid=false).  The host pads every group to a multiple of 32 slots, so
// a warp covers a single group: the three sums are REDUX-reduced and committed by lane 0 -- with a plain store
// when the whole group lives in this one warp (g carries kGroupSingleWarp: the common case, no CAS loop and no
// need for a cleared accumulator), with one shared-memory atomic per component otherwise.  A warp that does mix
// groups falls back to per-lane atomics.
constexpr int kGroupSingleWarp = 0x4000;
__device__ __forceinline__ void cv_accumulate(CvShared& sh, bool valid, int g, double w, double px,
                                              double py, double pz) {
    long long qx = 0, qy = 0, qz = 0;
    if (valid) {
        qx = __double2ll_rn(__dmul_rn(__dmul_rn(w, px), kFixedScale));
        qy = __double2ll_rn(__dmul_rn(__dmul_rn(w, py), kFixedScale));
        qz = __double2ll_rn(__dmul_rn(__dmul_rn(w, pz), kFixedScale));
    }
    const unsigned vmask = __ballot_sync(kFullWarp, valid);
    if (vmask == 0) return;
    const int gref = __shfl_sync(kFullWarp, g, __ffs(vmask) - 1);
    const bool uniform = __all_sync(kFullWarp, !valid || g == gref);
    const int gi = gref & (kGroupSingleWarp - 1);
    if (uniform) {
        qx = warp_sum_s64(qx); qy = warp_sum_s64(qy); qz = warp_sum_s64(qz);
        if ((threadIdx.x & 31) == 0) {
            if (gref & kGroupSingleWarp) {
                sh.acc[gi * 3 + 0] = qx; sh.acc[gi * 3 + 1] = qy; sh.acc[gi * 3 + 2] = qz;
            } else {
                atomicAdd((unsigned long long*) &sh.acc[gi * 3 + 0], (unsigned long long) qx);
                atomicAdd((unsigned long long*) &sh.acc[gi * 3 + 1], (unsigned long long) qy);
                atomicAdd((unsigned long long*) &sh.acc[gi * 3 + 2], (unsigned long long) qz);
            }
        }
    } else if (valid) {
        const int gl = g & (kGroupSingleWarp - 1);
        atomicAdd((unsigned long long*) &sh.acc[gl * 3 + 0], (unsigned long long) qx);
        atomicAdd((unsigned long long*) &sh.acc[gl * 3 + 1], (unsigned long long) qy);
        atomicAdd((unsigned long long*) &sh.acc[gl * 3 + 2], (unsigned long long) qz);
    }
}

// one slot's contribution to its group's fixed-point centroid sums
__device__ __forceinline__ void cv_fixed(double w, double px, double py, double pz, long long& qx, long long& qy, long long& qz) {
    qx = __double2ll_rn(__dmul_rn(__dmul_rn(w, px), kFixedScale));
    qy = __double2ll_rn(__dmul_rn(__dmul_rn(w, py), kFixedScale));
    qz = __double2ll_rn(__dmul_rn(__dmul_rn(w, pz), kFixedScale));
}

// cv_accumulate() for callers that know the host's slot layout holds (every group padded to whole warps, so the 32 lanes
// of a warp always belong to one group): no votes, straight to the REDUX sums
__device__ __forceinline__ void cv_accumulate_aligned(CvShared& sh, int g, double w, double px, double py, double pz) {
    long long qx = __double2ll_rn(__dmul_rn(__dmul_rn(w, px), kFixedScale));
    long long qy = __double2ll_rn(__dmul_rn(__dmul_rn(w, py), kFixedScale));
    long long qz = __double2ll_rn(__dmul_rn(__dmul_rn(w, pz), kFixedScale));
    qx = warp_sum_s64(qx); qy = warp_sum_s64(qy); qz = warp_sum_s64(qz);
    if ((threadIdx.x & 31) == 0) {
        const int gi = g & (kGroupSingleWarp - 1);
        if (g & kGroupSingleWarp) {
            sh.acc[gi * 3 + 0] = qx; sh.acc[gi * 3 + 1] = qy; sh.acc[gi * 3 + 2] = qz;
        } else {
            atomicAdd((unsigned long long*) &sh.acc[gi * 3 + 0], (unsigned long long) qx);
            atomicAdd((unsigned long long*) &sh.acc[gi * 3 + 1], (unsigned long long) qy);
            atomicAdd((unsigned long long*) &sh.acc[gi * 3 + 2], (unsigned long long) qz);
        }
    }
}

// the two rarer forms of the examples (demo4 pair-sum, demo5 angle): kept out of line so that the common path
// below stays short, straight-line code (it is executed once per launch by a single warp, on the critical path
// of the decision hand-off)
__device__ __noinline__ double cv_term_rare(const long long* acc, const Boundary* bp) {
    const Boundary& b = *bp;
    auto cen = [&](int g, int c) { return __dmul_rn(__ll2double_rn(acc[g * 3 + c]), kFixedInvScale); };
    auto sqdist = [&](int ga, int gb) {
        const double d0 = __dsub_rn(cen(ga, 0), cen(gb, 0)), d1 = __dsub_rn(cen(ga, 1), cen(gb, 1)), d2 = __dsub_rn(cen(ga, 2), cen(gb, 2));
        return __fma_rn(d2, d2, __fma_rn(d1, d1, __dmul_rn(d0, d0)));
    };
    if (b.type == SEEKR2_B200_CV_PAIR_SUM) {
        double sum = sqdist(b.groups[0], b.groups[1]);
        for (int p = 1; p < b.axis; p++) sum = __dadd_rn(sum, sqdist(b.groups[2 * p], b.groups[2 * p + 1]));
        return __dsub_rn(sum, __dmul_rn(b.threshold, b.threshold));
    }
    double a[3], v[3];                                   // SEEKR2_B200_CV_ANGLE
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double mid = cen(b.groups[1], c);
        a[c] = __dsub_rn(cen(b.groups[0], c), mid);
        v[c] = __dsub_rn(cen(b.groups[2], c), mid);
    }
    const double dot = __fma_rn(a[2], v[2], __fma_rn(a[1], v[1], __dmul_rn(a[0], v[0])));
    const double na2 = __fma_rn(a[2], a[2], __fma_rn(a[1], a[1], __dmul_rn(a[0], a[0])));
    const double nb2 = __fma_rn(v[2], v[2], __fma_rn(v[1], v[1], __dmul_rn(v[0], v[0])));
    return __dsub_rn(__dmul_rn(b.center[0], __dsqrt_rn(__dmul_rn(na2, nb2))), dot);   // center[0] = cos(ref)
}

__device__ __forceinline__ double fixed_to_nm(long long a) { return __dmul_rn(__ll2double_rn(a), kFixedInvScale); }

// Surface function u of the common forms -- sphere (pair or fixed centre) and plane, what MMVT anchors are made of -- from
// the fixed-point centroid sums of the boundary's first two groups, held in registers; computed without a branch and selected.
__device__ __forceinline__ double cv_u_common(const Boundary& b, long long a0, long long a1, long long a2,
                                              long long b0, long long b1, long long b2) {
    const bool pair = b.type == SEEKR2_B200_CV_SPHERE_PAIR;
    const double ax = fixed_to_nm(a0), ay = fixed_to_nm(a1), az = fixed_to_nm(a2);
    const double bx = pair ? fixed_to_nm(b0) : b.center[0];
    const double by = pair ? fixed_to_nm(b1) : b.center[1];
    const double bz = pair ? fixed_to_nm(b2) : b.center[2];
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    const double d2 = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
    const double u_sphere = __dsub_rn(d2, __dmul_rn(b.threshold, b.threshold));
    const double along = b.axis == 0 ? ax : (b.axis == 1 ? ay : az);
    const double u_plane = __dsub_rn(along, b.threshold);
    return b.type == SEEKR2_B200_CV_PLANE ? u_plane : u_sphere;
}

// term contributed to the force-group energy: bitcode * step(k u) (step(0) = 1), or k u for the older non-step() style
__device__ __forceinline__ double cv_term_of(const Boundary& b, double u) {
    const double ku = __dmul_rn(b.k, u);
    if (b.style == SEEKR2_B200_STYLE_RAW) return ku;
    return ku >= 0.0 ? b.bitcode : 0.0;
}

// one boundary term; evaluated by lane b of warp 0 after the accumulators in shared memory are complete.  b_mem is the same
// boundary in memory (the rare forms index groups[] dynamically; b itself may live in registers: only static indices below).
__device__ __forceinline__ double cv_term_regs(const CvShared& sh, const Boundary& b, const Boundary* b_mem) {
    if (b.type > SEEKR2_B200_CV_PLANE) return cv_term_of(b, cv_term_rare(sh.acc, b_mem));
    const int ga = b.groups[0], gb = b.groups[1];         // groups[1] is a valid index (0) when unused
    return cv_term_of(b, cv_u_common(b, sh.acc[ga * 3], sh.acc[ga * 3 + 1], sh.acc[ga * 3 + 2],
                                     sh.acc[gb * 3], sh.acc[gb * 3 + 1], sh.acc[gb * 3 + 2]));
}
__device__ __forceinline__ double cv_term(const CvShared& sh, const Boundary& b) { return cv_term_regs(sh, b, &b); }

// warp 0: lane b evaluates boundary b of this anchor into shared memory
__device__ __forceinline__ void cv_terms(CvShared& sh, const EngineDev& e, const Boundary& b, int lane = threadIdx.x) {
    if (lane < e.num_boundaries) {
        sh.term[lane] = (e.flags_bits & SEEKR2_B200_FLAG_CV_FLOAT) ? cvf_term(sh, b) : cv_term(sh, b);
        sh.force_group[lane] = b.force_group;
    }
    __syncwarp();
}
__device__ __forceinline__ void cv_terms(CvShared& sh, const EngineDev& e, int anchor) {
    Boundary b = {};
    if (threadIdx.x < e.num_boundaries) b = e.boundaries[(size_t) anchor * e.num_boundaries + threadIdx.x];
    cv_terms(sh, e, b);
}

// energy of one force group == context.calcForcesAndEnergy(false, true, 1 << group); one lane
__device__ __forceinline__ double cv_group_value(const CvShared& sh, int num_boundaries, int force_group) {
    double value = 0.0;
    for (int b = 0; b < num_boundaries; b++)
        if (sh.force_group[b] == force_group) value = __dadd_rn(value, sh.term[b]);
    return value;
}

constexpr int kFlagValid = 1, kFlagBounce = 2, kFlagSnapshot = 4, kFlagSlotShift = 4;

// If crossing states are kept and exactly one surface is crossed, claims the next snapshot slot:
// returns the flag bits for the streaming blocks and the snapshot's sequence number (0 = none).
__device__ __forceinline__ int claim_snapshot(const EngineDev& e, AnchorState& st, int num_surfaces, int& seq) {
    seq = 0;
    if (e.snap_slots <= 0 || num_surfaces != 1) return 0;
    seq = ++st.num_snapshots;
    return kFlagSnapshot | (((seq - 1) % e.snap_slots) << kFlagSlotShift);
}

// number of logged surfaces in an MMVT value (CudaSeekr2Kernels.cpp:264-270)
__device__ __forceinline__ int mmvt_num_surfaces(const EngineDev& e, float value) {
    int bitcode = (int) value, n = 0;
    for (int i = 0; i < e.num_milestone_groups; i++) { if ((bitcode % 2) != 0) n++; bitcode = bitcode >> 1; }
    return n;
}

__device__ __forceinline__ void ring_append(const EngineDev& e, AnchorState& st, int anchor, int kind,
                                            int index, int counter, int nsurf, double time, int snapshot = 0) {
    Record r;
    r.anchor = anchor; r.kind = kind; r.index = index; r.counter = counter;
    r.num_surfaces = nsurf; r.snapshot = snapshot; r.step = st.step_count; r.time = time;
    e.ring[(size_t) anchor * e.ring_capacity + (size_t) (st.ring_head % e.ring_capacity)] = r;
    st.ring_head++;
    if (e.hint) e.hint[anchor] = st.ring_head;                      // crossing steps only: one posted write over PCIe
}

// The host block of CudaIntegrateMmvtLangevinStepKernel::execute (CudaSeekr2Kernels.cpp:250-347),
// run by one thread of the anchor's first block.  Returns whether the step bounces.
__device__ __forceinline__ bool mmvt_bookkeeping(const EngineDev& e, AnchorState& st, int anchor, float value, int snapshot = 0) {
    st.last_value_positive = value > 0.0f;
    if (!(value > 0.0f)) return false;
    int bitcode = (int) value;                                     // :262
    int num_bounced_surfaces = 0;
    for (int i = 0; i < e.num_milestone_groups; i++) {              // :264-270
        if ((bitcode % 2) != 0) num_bounced_surfaces++;
        bitcode = bitcode >> 1;
    }
    bitcode = (int) value;
    for (int i = 0; i < e.num_milestone_groups; i++) {              // :273-310
        if ((bitcode % 2) == 0) { bitcode = bitcode >> 1; continue; }
        bitcode = bitcode >> 1;
        ring_append(e, st, anchor, SEEKR2_B200_REC_MMVT, i, st.bounce_counter, num_bounced_surfaces, st.time, snapshot);
        if (e.variant == SEEKR2_B200_VARIANT_REFERENCE || st.previous_milestone != -1)   // :294-296 / Reference :290
            st.N_alpha_beta[i] += 1;
        if (st.previous_milestone != i) {                           // :297-305
            if (st.previous_milestone != -1) {
                st.Nij_alpha[st.previous_milestone * SEEKR2_B200_MAX_MILESTONES + i] += 1;
                st.Ri_alpha[st.previous_milestone] = __dadd_rn(st.Ri_alpha[st.previous_milestone], st.incubation_time);
            } else {
                st.first_crossing_time = st.time;
            }
            st.incubation_time = 0.0;
        }
        st.T_alpha = __dsub_rn(st.time, st.first_crossing_time);    // :306
        st.previous_milestone = i;
        st.bounce_counter++;
    }
    st.num_bounce_steps++;
    return true;
}

// The crossing block of CudaIntegrateElberLangevinStepKernel::execute (CudaSeekr2Kernels.cpp:517-590).
// elber_values holds the stored source values then destination values; elber_sampled marks which
// have their baseline (the reference uses -INFINITY as "unsampled").
// Returns the snapshot flag bits (0 = none) when the run just ended on exactly one surface (:575).
__device__ __forceinline__ int elber_bookkeeping(const EngineDev& e, AnchorState& st, int anchor, const CvShared& sh) {
    if (st.end_simulation) return 0;                                  // :524
    const long long head0 = st.ring_head;
    int num_bounced_surfaces = 0;
    for (int i = 0; i < e.num_src; i++) {                           // :526-549
        const float value = (float) cv_group_value(sh, e.num_boundaries, e.src_groups[i]);
        if (!(st.elber_sampled & (1 << i))) { st.elber_values[i] = value; st.elber_sampled |= 1 << i; }
        const float oldvalue = st.elber_values[i];
        const bool time_ok = e.variant == SEEKR2_B200_VARIANT_REFERENCE ? true : (st.time != 0.0);   // :533 / Reference :472
        if (((value - oldvalue) != 0.0f) && time_ok) {
            if (e.end_on_src) {
                st.end_simulation = 1;
                num_bounced_surfaces++;
                ring_append(e, st, anchor, SEEKR2_B200_REC_ELBER_SRC, i, st.crossing_counter, 0, st.time);
            } else {
                st.crossed_src = 1;
                st.time = 0.0;                                      // context.setTime(0.0) :545
                st.clock_resets++;
                if (e.hint) e.hint[e.num_anchors + anchor] = st.clock_resets;
                st.elber_values[i] = value;
            }
        }
    }
    for (int i = 0; i < e.num_dest; i++) {                          // :551-572
        const int j = e.num_src + i;
        const float value = (float) cv_group_value(sh, e.num_boundaries, e.dest_groups[i]);
        if (!(st.elber_sampled & (1 << j))) { st.elber_values[j] = value; st.elber_sampled |= 1 << j; }
        const float oldvalue = st.elber_values[j];
        const bool time_ok = e.variant == SEEKR2_B200_VARIANT_REFERENCE ? true : (st.time != 0.0);   // :558 / Reference :498
        if (((value - oldvalue) != 0.0f) && time_ok) {
            st.end_simulation = 1;
            num_bounced_surfaces++;
            const bool star = !(st.crossed_src || e.end_on_src);    // :564-568
            ring_append(e, st, anchor, star ? SEEKR2_B200_REC_ELBER_DEST_STAR : SEEKR2_B200_REC_ELBER_DEST, i,
                        st.crossing_counter, 0, st.time);
        }
    }
    int snap_bits = 0;
    if (st.end_simulation) {                                        // :573-589
        int seq = 0;
        snap_bits = claim_snapshot(e, st, num_bounced_surfaces, seq);
        for (long long h = head0; h < st.ring_head; h++) {
            Record& r = e.ring[(size_t) anchor * e.ring_capacity + (size_t) (h % e.ring_capacity)];
            r.num_surfaces = num_bounced_surfaces;
            r.snapshot = seq;
        }
        st.crossing_counter++;
    }
    return snap_bits;
}

// step epilogue: CudaSeekr2Kernels.cpp:362-364 (MMVT) / :592-593 (Elber)
__device__ __forceinline__ void advance_clock(AnchorState& st, double dt, bool mmvt) {
    st.time = __dadd_rn(st.time, dt);
    st.step_count++;
    if (mmvt) st.incubation_time = __dadd_rn(st.incubation_time, dt);
}

}  // namespace s2
