/*
 * seekr2_b200.h -- C ABI of the B200-native SEEKR2 integrator step (MMVT / Elber Langevin).
 *
 * This is the drop-in seam for the per-timestep hot path of the reference plugin
 * (seekrcentral/seekr2_openmm_plugin).  Every entry point below replaces one piece of the
 * reference's CUDA-platform kernel implementation and is what a `platforms/b200` OpenMM
 * KernelImpl (platforms/b200/B200Seekr2Kernels.cpp) binds; the same symbols are bound by
 * ctypes (seekr2_openmm_plugin_b200/_lib.py) for the standalone harness and the tests.
 *
 * Reference interfaces replaced (paths relative to seekr2plugin/ in the reference tree):
 *   IntegrateMmvtLangevinStepKernel / IntegrateElberLangevinStepKernel
 *       openmmapi/include/Seekr2Kernels.h:53-110                (initialize / execute)
 *   CudaIntegrateMmvtLangevinStepKernel::{allocateMemory,initialize,execute}
 *       platforms/cuda/src/CudaSeekr2Kernels.cpp:82-366
 *   CudaIntegrateElberLangevinStepKernel::{initialize,execute}
 *       platforms/cuda/src/CudaSeekr2Kernels.cpp:391-595
 *   integrate{Mmvt,Elber}LangevinPart1/Part2, mmvtBounce
 *       platforms/cuda/src/kernels/langevin.cu:7-179
 *   IntegrateMmvtLangevinMiddleStepKernel / IntegrateElberLangevinMiddleStepKernel (scheme MIDDLE)
 *       openmmapi/include/Seekr2Kernels.h:115-172, platforms/cuda/src/kernels/langevinMiddle.cu:7-214
 *
 * Conventions
 *   - plain C, no C++/torch types; all pointers named d_* are CUDA device pointers, `stream`
 *     is a cudaStream_t passed as void*.
 *   - every function returns an int status (SEEKR2_B200_OK == 0); nothing throws.  The text of
 *     the last error of the calling thread is available from seekr2_b200_last_error().
 *   - buffers use the OpenMM CUDA-platform layouts:
 *       posq            real4 [A][Np]   (x,y,z,q)
 *       posqCorrection  real4 [A][Np]   low-order position bits, mixed precision only (else NULL)
 *       velm            mixed4[A][Np]   (vx,vy,vz,1/m); 1/m == 0 marks an immobile particle
 *       force           int64 [A][3][Np] fixed point, value = F * 2^32, SoA per anchor
 *       posDelta        mixed4[A][Np]   (two-pass path only)
 *       random          float4[A][random_anchor_stride], consumed at random_index + atom, .w unused
 *     A = number of anchors batched in one engine (A == 1 inside OpenMM), Np = padded atom count.
 *     real/mixed: SINGLE float/float, MIXED float/double, DOUBLE double/double.
 */
#ifndef SEEKR2_B200_H_
#define SEEKR2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEEKR2_B200_ABI_VERSION 9

#define SEEKR2_B200_MAX_MILESTONES 32   /* labels / bits tracked by the MMVT statistics          */
#define SEEKR2_B200_MAX_BOUNDARIES 32   /* boundary terms per anchor                              */
#define SEEKR2_B200_MAX_GROUPS 16       /* centroid groups                                        */
#define SEEKR2_B200_MAX_ELBER 16        /* src + dest milestones of an Elber integrator           */

/* status codes */
enum {
    SEEKR2_B200_OK = 0,
    SEEKR2_B200_ERR_INVALID = 1,        /* bad argument / configuration                           */
    SEEKR2_B200_ERR_CUDA = 2,           /* a CUDA runtime call failed                             */
    SEEKR2_B200_ERR_FIRST_STEP = 3,     /* MMVT: system is behind a boundary before the first step
                                           (CudaSeekr2Kernels.cpp:205-210)                        */
    SEEKR2_B200_ERR_RING_OVERFLOW = 4,  /* crossing records were dropped: drain more often        */
    SEEKR2_B200_ERR_IO = 5,             /* crossing / statistics file could not be written        */
    SEEKR2_B200_ERR_NO_FORCE_GROUP = 6, /* Elber: a milestone's force group has no boundary
                                           (CudaSeekr2Kernels.cpp:415-436)                        */
    SEEKR2_B200_ERR_SNAPSHOT_LOST = 7   /* a crossing-state snapshot was overwritten before it was
                                           fetched: drain more often or raise save_state_slots     */
};

/* CudaPrecision of the owning context */
enum { SEEKR2_B200_SINGLE = 0, SEEKR2_B200_MIXED = 1, SEEKR2_B200_DOUBLE = 2 };

/* which of the reference's two platforms the bounce bookkeeping follows          */
enum {
    SEEKR2_B200_VARIANT_CUDA = 0,       /* bounce restores -v(post-kick); N_alpha_beta skips the
                                           first crossing (CudaSeekr2Kernels.cpp:294-296)         */
    SEEKR2_B200_VARIANT_REFERENCE = 1   /* bounce restores -v(step start); N_alpha_beta always
                                           counts (ReferenceSeekr2Kernels.cpp:205-206,290,329-332)*/
};

/* flags */
enum {
    SEEKR2_B200_FLAG_RESTORE_CORRECTION = 1, /* on a bounce also restore posqCorrection.  The
                                                reference leaves the new correction in place
                                                (langevin.cu:167-179); default 0 is bit-faithful   */
    /* How the fused step (seekr2_b200_step) gets the crossing decision to the threads that store the
       atoms.  Results are bit-identical either way; default (neither flag) picks by launch size:
       LOCAL when the whole launch is resident at once (one or a few anchors per GPU, latency-bound),
       HANDOFF for batched, bandwidth-bound launches.                                                 */
    SEEKR2_B200_FLAG_DECIDE_LOCAL = 2,       /* every block evaluates the boundaries itself (needs
                                                save_state_slots == 0 and <= 256 padded group slots)  */
    SEEKR2_B200_FLAG_DECIDE_HANDOFF = 4,     /* one decide block per anchor publishes the decision    */
    /* Boundary value in SINGLE precision, the way OpenMM's CUDA platform evaluates a CustomCentroidBondForce in single
       and mixed precision (the reference reads that value, CudaSeekr2Kernels.cpp:248): float weights, float group
       centres accumulated sequentially in group order, float distance = sqrtf(dx*dx + dy*dy + dz*dz), float
       distance*distance - radius*radius, float k * u.  OpenMM itself is absent here, so this restates its evaluation
       (tools/cv_float_contract.py) rather than reproducing it bit for bit; the default (0) is the double / fixed-point
       contract of DESIGN.md section 3, which differs from this on ~1e-5 of the steps of a constantly bouncing run.
       SPHERE_PAIR / SPHERE_FIXED / PLANE boundaries, SINGLE or MIXED precision; the fused step runs the hand-off kernel. */
    SEEKR2_B200_FLAG_CV_FLOAT = 8
};

/* integrator kind */
enum { SEEKR2_B200_KIND_LANGEVIN = 0, SEEKR2_B200_KIND_MMVT = 1, SEEKR2_B200_KIND_ELBER = 2 };

/* discretisation of the Langevin equation.
 *   LANGEVIN : {Mmvt,Elber}LangevinIntegrator       -- kick (v,F,noise) then drift (langevin.cu)
 *   MIDDLE   : {Mmvt,Elber}LangevinMiddleIntegrator -- force kick, half drift, O step, half drift
 *              (langevinMiddle.cu:7-101,104-196; host CudaSeekr2Kernels.cpp:725-915,1032-1173).
 *              A bounce negates the velocity saved after the force kick (langevinMiddle.cu:41,
 *              :202-214).  Fused: seekr2_b200_step.  With constraints: kick / middle_drift / finish.       */
enum { SEEKR2_B200_SCHEME_LANGEVIN = 0, SEEKR2_B200_SCHEME_MIDDLE = 1 };

/* boundary surface types: the typed form of the CustomCentroidBondForce expressions the
 * reference's examples use.  u is the surface function; the term contributed to
 * the force-group energy is  bitcode*step(k*u)  (STEP style) or  k*u  (RAW, the older style).   */
enum {
    SEEKR2_B200_CV_SPHERE_PAIR = 0,   /* u = distance(g_a,g_b)^2 - threshold^2   (demo3:37)        */
    SEEKR2_B200_CV_SPHERE_FIXED = 1,  /* u = |c(g_a) - center|^2 - threshold^2   (demo1:59)        */
    SEEKR2_B200_CV_PLANE = 2,         /* u = c(g_a)[axis] - threshold            (demo2:61)        */
    SEEKR2_B200_CV_PAIR_SUM = 3,      /* u = sum_{p < axis} distance(g_2p, g_2p+1)^2 - threshold^2, axis = number
                                         of pairs, 1..3                            (demo4:27)        */
    SEEKR2_B200_CV_ANGLE = 4          /* angle(g_0, g_1, g_2) - threshold, evaluated without acos as the
                                         sign-equivalent  u = cos(threshold)*|a||b| - a.b  with a = c0 - c1,
                                         b = c2 - c1 (u >= 0 <=> angle >= threshold)  (demo5:24)        */
};
enum { SEEKR2_B200_STYLE_STEP = 0, SEEKR2_B200_STYLE_RAW = 1 };

typedef struct seekr2_b200_boundary {
    int32_t type;         /* SEEKR2_B200_CV_*                                                      */
    int32_t style;        /* SEEKR2_B200_STYLE_*                                                   */
    int32_t groups[6];    /* centroid group indices: [0] (and [1] for SPHERE_PAIR); [0..2] for ANGLE;
                             [0..2*axis-1] for PAIR_SUM                                             */
    int32_t axis;         /* 0,1,2 (PLANE); number of pairs (PAIR_SUM); else ignored               */
    int32_t force_group;  /* OpenMM force group of the Force carrying the term.  MMVT sums group 1
                             only (hard-wired mask 2, CudaSeekr2Kernels.cpp:206,248)               */
    double bitcode;       /* STEP style multiplier (1,2,4,...)                                     */
    double k;             /* sign selects the side; magnitude matters only for RAW style          */
    double threshold;     /* radius (nm) or plane position (nm)                                    */
    double center[3];     /* SPHERE_FIXED centre (nm).  ANGLE: center[0] is overwritten by the library
                             with cos(threshold)                                                    */
} seekr2_b200_boundary;

typedef struct seekr2_b200_config {
    int32_t abi_version;          /* SEEKR2_B200_ABI_VERSION                                       */
    int32_t kind;                 /* SEEKR2_B200_KIND_*                                            */
    int32_t precision;            /* SEEKR2_B200_SINGLE / MIXED / DOUBLE                           */
    int32_t variant;              /* SEEKR2_B200_VARIANT_*                                         */
    int32_t flags;                /* SEEKR2_B200_FLAG_*                                            */
    int32_t scheme;               /* SEEKR2_B200_SCHEME_*                                          */
    int32_t device;               /* CUDA device ordinal; -1 = current device                      */
    int32_t num_anchors;          /* A: independent replicas of the system batched per launch      */
    int32_t num_atoms;            /* N                                                             */
    int32_t padded_num_atoms;     /* Np >= N                                                       */

    /* centroid groups, shared by all anchors (same molecular system) */
    int32_t num_groups;
    const int32_t* group_offsets; /* [num_groups+1] into group_atoms / group_weights               */
    const int32_t* group_atoms;   /* atom indices                                                  */
    const double* group_weights;  /* unnormalised weights (masses); normalised by the library      */

    /* boundary terms: [num_anchors][num_boundaries]; anchors share the structure (type, groups,
       force_group) and differ only in the numbers (bitcode, k, threshold, center)                 */
    int32_t num_boundaries;
    const seekr2_b200_boundary* boundaries;

    /* MMVT: number of addMilestoneGroup() calls = number of bits that are logged and counted
       (MmvtLangevinIntegrator.cpp:133-136); initial bounceCounter per anchor (may be NULL = 0)    */
    int32_t num_milestone_groups;
    const int32_t* bounce_counter0;     /* [num_anchors] or NULL                                   */

    /* Elber: force groups of the source and destination milestones, in add order                  */
    int32_t num_src;
    const int32_t* src_groups;          /* [num_src]                                               */
    int32_t num_dest;
    const int32_t* dest_groups;         /* [num_dest]                                              */
    int32_t end_on_src_milestone;
    const int32_t* crossing_counter0;   /* [num_anchors] or NULL                                   */

    int32_t ring_capacity;        /* crossing records kept per anchor between drains (0 = 4096)    */
    int32_t save_state_slots;     /* > 0: keep the positions / velocities of the last `slots` single-
                                     surface crossings per anchor as they were BEFORE the bounce
                                     (saveStateFileName, CudaSeekr2Kernels.cpp:282-293, :575-587);
                                     costs slots * 64 B per atom and anchor of device memory         */
} seekr2_b200_config;

/* device buffers of one step call (borrowed, never owned) */
typedef struct seekr2_b200_buffers {
    void* d_posq;
    void* d_posq_correction;      /* NULL unless MIXED                                             */
    void* d_velm;
    const long long* d_force;
    void* d_pos_delta;            /* two-pass path only; may be NULL for the fused step            */
    const void* d_random;         /* float4 pool; NULL = in-kernel Philox (seekr2_b200_set_rng)    */
    int64_t random_anchor_stride; /* float4 elements between consecutive anchors' pools            */
    void* d_old_velm;             /* multi-pass scratch, mixed4[A][Np]: VARIANT_REFERENCE (start-of-step
                                     velocity) and every MIDDLE MMVT engine (the velocity Part2 saves
                                     before the O step, langevinMiddle.cu:41)                        */
    void* d_old_delta;            /* MIDDLE multi-pass path only: mixed4[A][Np] (oldDelta,
                                     CudaSeekr2Kernels.cpp:630-643)                                  */
} seekr2_b200_buffers;

/* one crossing event, as appended by the device to the per-anchor ring */
enum {
    SEEKR2_B200_REC_MMVT = 0,       /* MMVT bounce on bit `index`                                  */
    SEEKR2_B200_REC_ELBER_SRC = 1,  /* Elber: ended on source milestone `index`                    */
    SEEKR2_B200_REC_ELBER_DEST = 2, /* Elber: ended on destination milestone `index`               */
    SEEKR2_B200_REC_ELBER_DEST_STAR = 3 /* same, source never crossed: label gets a '*' suffix     */
};
typedef struct seekr2_b200_record {
    int32_t anchor;
    int32_t kind;        /* SEEKR2_B200_REC_*                                                      */
    int32_t index;       /* bit index i (MMVT) or position in the src / dest list (Elber)          */
    int32_t counter;     /* bounceCounter / crossingCounter written on the line                    */
    int32_t num_surfaces;/* surfaces crossed in the same step (state is saved only when == 1)      */
    int32_t snapshot;    /* sequence number (1,2,...) of the saved crossing state, 0 = none          */
    int64_t step;        /* stepCount of the step that crossed                                     */
    double time;         /* context time written on the line (ps)                                  */
} seekr2_b200_record;

/* per-anchor integrator state kept on the device (read back by seekr2_b200_get_state) */
typedef struct seekr2_b200_anchor_state {
    double time;                 /* cu.getTime()                                                   */
    double incubation_time;
    double first_crossing_time;
    double T_alpha;
    double Ri_alpha[SEEKR2_B200_MAX_MILESTONES];
    int64_t step_count;
    int64_t ring_head;           /* records appended so far (monotonic)                            */
    int32_t N_alpha_beta[SEEKR2_B200_MAX_MILESTONES];
    int32_t Nij_alpha[SEEKR2_B200_MAX_MILESTONES * SEEKR2_B200_MAX_MILESTONES];
    int32_t bounce_counter;
    int32_t previous_milestone;  /* -1 before the first bounce                                     */
    int32_t num_bounce_steps;    /* steps that ended in a bounce (diagnostic)                      */
    int32_t last_value_positive; /* result of the most recent CV evaluation (first-step check)     */
    /* Elber */
    int32_t crossing_counter;
    int32_t crossed_src;
    int32_t end_simulation;
    int32_t elber_sampled;       /* bit i set once milestone i has its baseline value              */
    float elber_values[SEEKR2_B200_MAX_ELBER]; /* src values then dest values                      */
    int32_t num_snapshots;       /* crossing states saved so far (save_state_slots > 0)            */
    int32_t clock_resets;        /* Elber: times the step set the context time back to 0 (:545)     */
} seekr2_b200_anchor_state;

typedef struct seekr2_b200_engine* seekr2_b200_handle;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* replaces Cuda*StepKernel::allocateMemory + initialize (CudaSeekr2Kernels.cpp:82-173,391-463):
   allocates the group slab, per-anchor state and rings; nothing is allocated per step.           */
int seekr2_b200_create(const seekr2_b200_config* config, seekr2_b200_handle* out);
int seekr2_b200_destroy(seekr2_b200_handle h);
const char* seekr2_b200_last_error(void);
int seekr2_b200_abi_version(void);

/* 1 if the fused step of this engine runs in LOCAL decision mode, 0 for the decide-block hand-off   */
int seekr2_b200_decision_mode(seekr2_b200_handle h);

/* ---- parameters ---------------------------------------------------------------------------- */

/* replaces the parameter block of execute (CudaSeekr2Kernels.cpp:188-203 / :474-489):
   vscale = exp(-dt*gamma), fscale = gamma==0 ? dt : (1-vscale)/gamma,
   noisescale = sqrt(BOLTZ*T*(1-vscale^2)).  Cheap; call whenever T, gamma or dt change.          */
int seekr2_b200_set_params(seekr2_b200_handle h, double temperature, double friction, double step_size);
int seekr2_b200_get_params(seekr2_b200_handle h, double out_vscale_fscale_noisescale[3]);

/* replaces integration.initRandomNumberGenerator(seed) (CudaSeekr2Kernels.cpp:109,396) for callers
   that pass d_random == NULL: the step kernels then generate their Gaussian numbers themselves
   (Philox4x32-10 keyed by `seed`, counter = (atom, anchor_base + anchor * anchor_id_stride, step)) instead of reading
   OpenMM's pool.  The counter's anchor word must be the GLOBAL anchor id, so that one seed serves all ranks without two
   anchors sharing a stream: anchors dealt round-robin -> anchor_base = rank, anchor_id_stride = world size; contiguous
   blocks -> anchor_base = first anchor of the rank, anchor_id_stride = 1 (0 is read as 1).                    */
int seekr2_b200_set_rng(seekr2_b200_handle h, uint64_t seed, uint32_t anchor_base, uint32_t anchor_id_stride);

/* ---- state the kernels need besides OpenMM's buffers ---------------------------------------- */

/* (re)gather the boundary-group atoms' state into the engine's double-buffered slab and evaluate
   the boundary value at the current positions.  Call once before the first fused step and after
   anything else modified posq/velm (setPositions, setVelocities, reorderAtoms, a barostat).      */
int seekr2_b200_sync_groups(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, void* stream);

/* The light form of sync_groups for a caller that steps through seekr2_b200_step but cannot know whether something
   else touched posq / velm since its last step (inside OpenMM: Context::setPositions / setVelocities, a barostat,
   CMMotionRemover, atom reordering -- the reference is immune because it re-reads everything every step,
   CudaSeekr2Kernels.cpp:214-246): ONE small launch that re-gathers the boundary-group atoms into the slab and takes the
   step index from the anchor clock.  Queue it in front of every fused step; the pair is still a single pass over the
   atoms (no posDelta round trip).  Not needed between back-to-back fused steps of a caller that owns the buffers.
   When the step that follows draws its normals from a Gaussian pool (bufs->d_random != NULL, OpenMM's convention), it is
   launched as a programmatic dependent of this kernel and overlaps it (it reads slab and step index behind its own wait). */
int seekr2_b200_resync(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, void* stream);

/* OpenMM's CUDA platform reorders atoms now and then (cu.reorderAtoms(), CudaSeekr2Kernels.cpp:365): the buffers then
   hold atom cu.getAtomIndex()[i] at index i.  index_of_atom[a] = current buffer index of the atom the config called `a`
   ([num_atoms], host memory; the inverse of getAtomIndex()).  Remaps the group slots; stream-ordered behind the steps
   already queued.  Follow with seekr2_b200_resync / sync_groups before the next fused step.                           */
int seekr2_b200_set_atom_order(seekr2_b200_handle h, const int32_t* index_of_atom, void* stream);

/* replaces the first-step test (CudaSeekr2Kernels.cpp:205-210): evaluates the MMVT boundary value
   at the current positions; returns SEEKR2_B200_ERR_FIRST_STEP if any anchor has value > 0 and
   writes per-anchor flags to out_flags[num_anchors] when non-NULL.  Synchronises the stream.     */
int seekr2_b200_check_first_step(seekr2_b200_handle h, const seekr2_b200_buffers* bufs,
                                 int32_t* out_flags, void* stream);

/* ---- the step: fused single-pass form (systems without constraints / harness) ---------------- */

/* one full integrator step in ONE launch: Langevin kick + drift (langevin.cu:7-32,65-112), boundary
   CV evaluation, crossing test, predicated bounce (langevin.cu:167-179), crossing-record append and
   the statistics / time / step-count epilogue (CudaSeekr2Kernels.cpp:250-365).  kind of the engine
   selects MMVT, Elber (CudaSeekr2Kernels.cpp:517-594) or plain Langevin.  Asynchronous on `stream`;
   capturable into a CUDA graph.                                                                    */
int seekr2_b200_step(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, uint32_t random_index,
                     void* stream);

/* ---- the step: two-pass form (constraints act on posDelta between the passes) ---------------- */

/* replaces integrate{Mmvt,Elber}LangevinPart1 (langevin.cu:7-59)                                  */
int seekr2_b200_kick(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, uint32_t random_index,
                     void* stream);
/* MIDDLE scheme only: replaces integrate{Mmvt,Elber}LangevinMiddlePart2 (langevinMiddle.cu:28-61,
   :125-157): half drift, O step, half drift -> posDelta / oldDelta; call between the velocity
   constraints and the position constraints.  For the MIDDLE scheme seekr2_b200_kick is Part1 (force
   kick, :7-21) and seekr2_b200_finish is decision + Part3 (:68-101) with the predicated bounce.       */
int seekr2_b200_middle_drift(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, uint32_t random_index,
                             void* stream);
/* replaces Part2 + calcForcesAndEnergy(false,true,mask) + the host crossing block + mmvtBounce
   (CudaSeekr2Kernels.cpp:231-365 / :509-594) in ONE launch: the grid's first num_anchors blocks evaluate the boundaries
   at the pending positions (posq + posDelta) and hand the decision to the streaming blocks of the same launch, which
   have loaded and drifted their atoms meanwhile and store with the bounce predicated on it.  A captured graph must
   hold an even number of finish launches (the decision word alternates between two halves).        */
int seekr2_b200_finish(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, void* stream);

/* ---- CUDA graph of a sequence of steps -------------------------------------------------------- */

/* begin capturing on `stream`; every seekr2_b200_* step call (and any kernels of the caller's own,
   e.g. a force provider) issued on that stream until graph_end becomes part of the graph.          */
int seekr2_b200_graph_begin(seekr2_b200_handle h, void* stream);
int seekr2_b200_graph_end(seekr2_b200_handle h, void* stream);
int seekr2_b200_graph_launch(seekr2_b200_handle h, void* stream);

/* ---- results ---------------------------------------------------------------------------------- */

/* The crossing log of the reference is appended inside execute() (CudaSeekr2Kernels.cpp:257-311); here records
   wait in a per-anchor device ring and are drained ASYNCHRONOUSLY:
     drain_begin  enqueues, on `stream`, one small kernel that packs every record appended since the previous drain
                  (all anchors, anchor-major, each anchor's records in order, at most max_records and at most 65536)
                  straight into mapped pinned host memory, plus an event; it returns at once and further steps may be
                  queued behind it;
     drain_end    waits for that event only (never for work queued after drain_begin), copies the records to `out`,
                  *out_n = number written, *out_waiting (optional) = records still in the rings (call again).
   Returns SEEKR2_B200_ERR_RING_OVERFLOW from drain_end if a ring wrapped over undrained records.  One drain in
   flight per engine.                                                                                    */
int seekr2_b200_drain_begin(seekr2_b200_handle h, int64_t max_records, void* stream);
int seekr2_b200_drain_end(seekr2_b200_handle h, seekr2_b200_record* out, int64_t max_records, int64_t* out_n,
                          int64_t* out_waiting);
/* 1 if seekr2_b200_drain_end would return without waiting (the drain in flight has landed in host memory, or none is
   in flight), 0 if the device has not got there yet, -1 on error.  Never blocks: lets a caller that issues one step
   per call (the OpenMM adapter) collect records whenever they happen to be there.                            */
int seekr2_b200_drain_ready(seekr2_b200_handle h);
/* Every drain also carries each anchor's clock: the context time and step count the device held when the pack kernel ran
   (cu.getTime() / cu.getStepCount() of the reference, which Elber resets on the device, CudaSeekr2Kernels.cpp:545).  Valid
   after seekr2_b200_drain_end; reads host memory only, never blocks.  Lets the OpenMM adapter mirror Context time without a
   read-back per step.                                                                                              */
int seekr2_b200_drain_clock(seekr2_b200_handle h, int32_t anchor, double* time, int64_t* step_count);
/* What the device has reported so far WITHOUT any CUDA call: the step kernels mirror each anchor's ring head and (Elber)
   the number of clock resets into mapped pinned host memory when they change, i.e. on crossing steps only.
   *records_pending = records appended by steps that have already run, minus records handed out by drain_end;
   *clock_resets = context-time resets so far (CudaSeekr2Kernels.cpp:545), summed over anchors.  Either may be NULL.
   A hint (steps still queued have not reported yet): lets a one-step-per-call caller queue a drain only when there is
   something to fetch instead of after every step.                                                              */
int seekr2_b200_poll(seekr2_b200_handle h, int64_t* records_pending, int64_t* clock_resets);
/* drain_begin + drain_end: blocks until everything queued on `stream` before the call has run.             */
int seekr2_b200_drain(seekr2_b200_handle h, seekr2_b200_record* out, int64_t max_records,
                      int64_t* out_n, void* stream);

int seekr2_b200_get_state(seekr2_b200_handle h, int32_t anchor, seekr2_b200_anchor_state* out,
                          void* stream);
/* Context::setTime / setStepCount analogue (used by reinitialize, the OpenMM adapter and tests): one small
   launch on `stream` carrying the values as parameters -- stream-ordered, never blocks the host.     */
int seekr2_b200_set_time(seekr2_b200_handle h, int32_t anchor, double time, int64_t step_count,
                         void* stream);

/* copy crossing-state snapshot `snapshot` (from a record) of `anchor` to the host as doubles:
   positions[N][3] = posq (+ posqCorrection), velocities[N][3].  SEEKR2_B200_ERR_SNAPSHOT_LOST if more
   than save_state_slots snapshots of that anchor were taken since.  Synchronises `stream`.          */
int seekr2_b200_fetch_snapshot(seekr2_b200_handle h, int32_t anchor, int32_t snapshot, double* positions,
                               double* velocities, void* stream);
/* write positions / velocities as an OpenMM `State` XML document (version 1: PeriodicBoxVectors,
   Positions, Velocities), the file the reference produces with XmlSerializer::serialize<State>
   (CudaSeekr2Kernels.cpp:283-292).  box = 3x3 row vectors (nm) or NULL.  The exact number formatting of
   OpenMM's serializer is not reproduced (OpenMM absent here); values are written with 17 digits.       */
int seekr2_b200_write_state_xml(const char* path, int32_t num_atoms, const double* positions,
                                const double* velocities, double time, const double* box);

/* ---- crossing / statistics files (host side of CudaSeekr2Kernels.cpp:158-171,257-345,459-460) -- */

/* write the header iff `path` does not exist (append mode otherwise untouched)                     */
int seekr2_b200_write_header(int32_t kind, const char* path);
/* append records of one anchor formatted exactly like the reference:
   MMVT  "<label>,<counter>,<time %.3f>\n"; Elber "<label>[*],<counter>,<time %g>\n".
   labels[index] is the milestone group id of bit / list position `index` (src labels first, then
   dest labels, for Elber).                                                                         */
int seekr2_b200_append_records(int32_t kind, const char* path, const seekr2_b200_record* recs,
                               int64_t n, int32_t anchor, const int32_t* labels, int32_t num_labels,
                               int32_t num_src);
/* rewrite the MMVT statistics file (CudaSeekr2Kernels.cpp:325-343)                                 */
int seekr2_b200_write_statistics(const char* path, const seekr2_b200_anchor_state* st,
                                 const int32_t* labels, int32_t num_labels);

/* debug: when d_probe (int64[8*A + 4 + 16], device) is set the fused kernel records %globaltimer stamps of
   each decide block ([0] start, [1] loads returned, [2] atoms advanced, [3] accumulated, [4] groups
   reduced, [5] decision published, [6] earliest streaming arrival of the anchor) and the streaming
   blocks' total wait ([8A..]: blocks that waited, polls, ns).  NULL disables it.                     */
int seekr2_b200_debug_set_probe(seekr2_b200_handle h, void* d_probe);

/* ---- harness kernels (stand-ins for what OpenMM supplies; not part of the drop-in path) -------- */

/* harmonic tether force  F = -k (x - x0)  written as int64 fixed point (value * 2^32)              */
int seekr2_b200_harness_tether_force(int32_t precision, int32_t num_anchors, int32_t num_atoms,
                                     int32_t padded_num_atoms, const void* d_posq,
                                     const void* d_posq_correction, const void* d_x0 /* real4 */,
                                     double k, long long* d_force, void* stream);
/* harmonic bond force over a bond list (i, j, r0, k), accumulated with 64-bit atomics             */
int seekr2_b200_harness_bond_force(int32_t precision, int32_t num_atoms, int32_t padded_num_atoms,
                                   const void* d_posq, const void* d_posq_correction,
                                   int32_t num_bonds, const int32_t* d_bond_atoms /* [nb][2] */,
                                   const double* d_bond_params /* [nb][2] = r0,k */,
                                   long long* d_force, void* stream);
/* fill a Gaussian pool with exactly the numbers the in-kernel generator would use:
   pool[a*anchor_stride + s*padded + i] = N(seed; atom i, anchor anchor_base + a*anchor_id_stride, step step0 + s)   */
int seekr2_b200_harness_fill_pool(void* d_random, int32_t num_anchors, int32_t padded_num_atoms,
                                  int64_t anchor_stride, int32_t num_steps, uint64_t seed,
                                  uint32_t anchor_base, uint32_t anchor_id_stride, uint64_t step0, void* stream);

/* test hook of the in-kernel generator: *d_mismatches = the number of float bit patterns (of all 2^32) on which the call-free
   square root the Box-Muller step uses (kernels.cuh, sqrt_no_call) differs from sqrtf; must be 0                              */
int seekr2_b200_harness_check_sqrt(unsigned long long* d_mismatches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEEKR2_B200_H_ */
