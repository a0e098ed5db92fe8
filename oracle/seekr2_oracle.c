/*
 * oracle/seekr2_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the SEEKR2 OpenMM plugin's per-timestep hot path, used as the
 * parity checker for the B200 kernels and as the timed CPU baseline of bench.py.  Nothing in the
 * product path (seekr2_openmm_plugin_b200/, include/, platforms/) may call into this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY PINNING STATUS:
 *   - Langevin arithmetic (Part1 / Part2 / bounce): PINNED against the reference's own CUDA
 *     kernels, compiled unmodified from /root/reference by oracle/build_ref.sh into oracle/_ref/
 *     and run on the GPU by tests/test_gpu_reference_kernels.py, and against the analytic
 *     known-answer cases of the reference's platform tests (testSingleBond, testTemperature,
 *     testConstrainedMasslessParticles; TestReferenceMmvtLangevinIntegrator.cpp:63-225).
 *   - CV evaluation, crossing bookkeeping, log format, statistics, Elber logic: PARITY UNPINNED.
 *     The reference evaluates CVs inside OpenMM (absent here, un-vendored, un-pinned version) and
 *     ships no test, fixture or golden file that exercises a crossing.  These parts restate
 *     CudaSeekr2Kernels.cpp:205-365 and :517-594 line by line and are anchored only on that text.
 *
 * What is restated (paths relative to /root/reference/seekr2plugin):
 *   platforms/cuda/src/kernels/langevin.cu:7-179          per-atom arithmetic (seekr2_oracle_kernels.inc)
 *   platforms/cuda/src/CudaSeekr2Kernels.cpp:188-203      integration parameters
 *   platforms/cuda/src/CudaSeekr2Kernels.cpp:205-365      MMVT step control flow, log, statistics
 *   platforms/cuda/src/CudaSeekr2Kernels.cpp:517-594      Elber step control flow
 *   platforms/reference/src/ReferenceSeekr2Kernels.cpp:203-208,290,328-334   VARIANT_REFERENCE
 *   OpenMM CustomCentroidBondForce semantics for the three expression forms the examples use
 *   (examples/mmvt_demonstration{1,2,3}*.py) -- restated from the published behaviour: weighted
 *   centroid with normalised weights, Euclidean distance, step(u) = 1 for u >= 0.
 *
 * Boundary-value arithmetic contract (ours; shared with the CUDA kernels, cv.cuh):
 *   P      = (double) posq  (the float position in single / mixed precision: OpenMM's force
 *            kernels never read posqCorrection)
 *   acc_g  = sum over atoms of group g of llrint((w_i * P_i) * 2^40)      (int64, order-free)
 *   C_g    = (double) acc_g * 2^-40
 *   SPHERE_PAIR  u = fma(dz,dz, fma(dy,dy, dx*dx)) - r*r,  d = C_a - C_b
 *   SPHERE_FIXED same with d = C_a - center
 *   PLANE        u = C_a[axis] - threshold
 *   PAIR_SUM     u = (sum over pairs, left to right, of |C_2p - C_2p+1|^2) - r*r
 *   ANGLE        u = cos(ref)*sqrt(|a|^2*|b|^2) - a.b,  a = C_0 - C_1, b = C_2 - C_1   (no acos: u >= 0 <=> angle >= ref)
 *   term = (k*u >= 0 ? bitcode : 0)   (STEP)   or   k*u   (RAW)
 *   value(force group) = sum of terms in boundary order, in double, starting from +0.0
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "../include/seekr2_b200.h"

/* OpenMM SimTKOpenMMRealType.h: BOLTZ = RGAS / KILO, RGAS = BOLTZMANN * AVOGADRO (CODATA 2018
   exact values, OpenMM >= 7.5); used at CudaSeekr2Kernels.cpp:191. */
#define ORACLE_BOLTZMANN 1.380649e-23
#define ORACLE_AVOGADRO 6.02214076e23
#define ORACLE_RGAS (ORACLE_BOLTZMANN * ORACLE_AVOGADRO)
#define ORACLE_BOLTZ (ORACLE_RGAS / 1000.0)

#define FIXED_SCALE 1099511627776.0            /* 2^40 */
#define FIXED_INV_SCALE (1.0 / 1099511627776.0)

/* ---- three instantiations of the per-atom arithmetic ---------------------------------------- */
#define REAL float
#define MIXED float
#define SUFFIX s
#define REAL_IS_DOUBLE 0
#define MIXED_IS_DOUBLE 0
#define HAS_CORRECTION 0
#include "seekr2_oracle_kernels.inc"
#undef REAL
#undef MIXED
#undef SUFFIX
#undef REAL_IS_DOUBLE
#undef MIXED_IS_DOUBLE
#undef HAS_CORRECTION

#define REAL float
#define MIXED double
#define SUFFIX m
#define REAL_IS_DOUBLE 0
#define MIXED_IS_DOUBLE 1
#define HAS_CORRECTION 1
#include "seekr2_oracle_kernels.inc"
#undef REAL
#undef MIXED
#undef SUFFIX
#undef REAL_IS_DOUBLE
#undef MIXED_IS_DOUBLE
#undef HAS_CORRECTION

#define REAL double
#define MIXED double
#define SUFFIX d
#define REAL_IS_DOUBLE 1
#define MIXED_IS_DOUBLE 1
#define HAS_CORRECTION 0
#include "seekr2_oracle_kernels.inc"
#undef REAL
#undef MIXED
#undef SUFFIX
#undef REAL_IS_DOUBLE
#undef MIXED_IS_DOUBLE
#undef HAS_CORRECTION

/* ---- oracle object ---------------------------------------------------------------------------- */

typedef struct seekr2_oracle {
    int kind, precision, variant, flags, scheme;
    int numAtoms, paddedNumAtoms;
    int numGroups;
    int* groupOffsets;
    int* groupAtoms;
    double* groupWeights;             /* normalised */
    int numBoundaries;
    seekr2_b200_boundary* boundaries; /* this anchor's */
    int numMilestoneGroups;
    int numSrc, numDest;
    int srcGroups[SEEKR2_B200_MAX_ELBER], destGroups[SEEKR2_B200_MAX_ELBER];
    int endOnSrcMilestone;
    double params[3];
    double temperature, friction, stepSize;
    int anchor;
    /* host-side state of the reference kernel object */
    seekr2_b200_anchor_state st;
    double srcValues[SEEKR2_B200_MAX_ELBER], destValues[SEEKR2_B200_MAX_ELBER]; /* vector<double>, -INFINITY = unsampled (CudaSeekr2Kernels.h:151-152) */
    /* scratch: posDelta + old state (the reference's oldPosq / oldVelm, CudaSeekr2Kernels.cpp:82-102) */
    void* posDelta;
    void* oldDelta;                   /* LangevinMiddle (CudaSeekr2Kernels.cpp:630-643) */
    void* oldPosq;
    void* oldPosqCorrection;
    void* oldVelm;
    void* startVelm;                  /* VARIANT_REFERENCE: velocities at step start (ReferenceSeekr2Kernels.cpp:205-206) */
    void* startForce;                 /* VARIANT_REFERENCE: the third per-step copy (:207-208) */
    /* crossing records */
    seekr2_b200_record* records;
    int64_t numRecords, capRecords;
    /* crossing states (saveStateFileName, CudaSeekr2Kernels.cpp:282-293,575-587): every one is kept,
       as doubles [numSnapshots][2][numAtoms][3] = what getState(Positions | Velocities) returns */
    int saveStates;
    double* snapshots;
    int64_t capSnapshots;
} seekr2_oracle;

static size_t real4_size(int precision) { return precision == SEEKR2_B200_DOUBLE ? 32 : 16; }
static size_t mixed4_size(int precision) { return precision == SEEKR2_B200_SINGLE ? 16 : 32; }

void seekr2_oracle_destroy(seekr2_oracle* o) {
    if (!o) return;
    free(o->groupOffsets); free(o->groupAtoms); free(o->groupWeights); free(o->boundaries);
    free(o->posDelta); free(o->oldDelta); free(o->oldPosq); free(o->oldPosqCorrection); free(o->oldVelm);
    free(o->startVelm); free(o->startForce); free(o->records); free(o->snapshots);
    free(o);
}

/* Builds the oracle for one anchor of a configuration (the same struct the product takes). */
seekr2_oracle* seekr2_oracle_create(const seekr2_b200_config* cfg, int anchor) {
    if (!cfg || anchor < 0 || anchor >= cfg->num_anchors) return NULL;
    /* the same limits the engine enforces (seekr2_b200_create): fixed-size tables below */
    if (cfg->num_groups < 0 || cfg->num_groups > SEEKR2_B200_MAX_GROUPS || cfg->num_boundaries < 0 || cfg->num_boundaries > SEEKR2_B200_MAX_BOUNDARIES ||
        cfg->num_milestone_groups < 0 || cfg->num_milestone_groups > SEEKR2_B200_MAX_MILESTONES ||
        cfg->num_src < 0 || cfg->num_dest < 0 || cfg->num_src + cfg->num_dest > SEEKR2_B200_MAX_ELBER) return NULL;
    seekr2_oracle* o = (seekr2_oracle*) calloc(1, sizeof(seekr2_oracle));
    o->kind = cfg->kind; o->precision = cfg->precision; o->variant = cfg->variant; o->flags = cfg->flags; o->scheme = cfg->scheme;
    o->numAtoms = cfg->num_atoms; o->paddedNumAtoms = cfg->padded_num_atoms;
    o->anchor = anchor;
    o->numGroups = cfg->num_groups;
    o->groupOffsets = (int*) calloc((size_t) cfg->num_groups + 1, sizeof(int));
    int total = 0;
    if (cfg->num_groups > 0) {
        memcpy(o->groupOffsets, cfg->group_offsets, sizeof(int) * ((size_t) cfg->num_groups + 1));
        total = cfg->group_offsets[cfg->num_groups];
    }
    o->groupAtoms = (int*) calloc((size_t) total + 1, sizeof(int));
    o->groupWeights = (double*) calloc((size_t) total + 1, sizeof(double));
    for (int g = 0; g < cfg->num_groups; g++) {
        /* OpenMM normalises the group weights so they sum to one (CustomCentroidBondForce) */
        double sum = 0.0;
        for (int j = cfg->group_offsets[g]; j < cfg->group_offsets[g + 1]; j++) sum += cfg->group_weights[j];
        for (int j = cfg->group_offsets[g]; j < cfg->group_offsets[g + 1]; j++) {
            o->groupAtoms[j] = cfg->group_atoms[j];
            o->groupWeights[j] = cfg->group_weights[j] / sum;
        }
    }
    o->numBoundaries = cfg->num_boundaries;
    o->boundaries = (seekr2_b200_boundary*) calloc((size_t) cfg->num_boundaries + 1, sizeof(seekr2_b200_boundary));
    if (cfg->num_boundaries > 0)
        memcpy(o->boundaries, cfg->boundaries + (size_t) anchor * cfg->num_boundaries,
               sizeof(seekr2_b200_boundary) * (size_t) cfg->num_boundaries);
    for (int b = 0; b < cfg->num_boundaries; b++)
        if (o->boundaries[b].type == SEEKR2_B200_CV_ANGLE) o->boundaries[b].center[0] = cos(o->boundaries[b].threshold);
    o->numMilestoneGroups = cfg->num_milestone_groups;
    o->numSrc = cfg->num_src; o->numDest = cfg->num_dest;
    for (int i = 0; i < cfg->num_src; i++) { o->srcGroups[i] = cfg->src_groups[i]; o->srcValues[i] = -INFINITY; }
    for (int i = 0; i < cfg->num_dest; i++) { o->destGroups[i] = cfg->dest_groups[i]; o->destValues[i] = -INFINITY; }
    o->endOnSrcMilestone = cfg->end_on_src_milestone;
    o->saveStates = cfg->save_state_slots > 0;
    /* initialize(): CudaSeekr2Kernels.cpp:118-133,151-154 and :403-404 */
    memset(&o->st, 0, sizeof(o->st));
    o->st.bounce_counter = cfg->bounce_counter0 ? cfg->bounce_counter0[anchor] : 0;
    o->st.crossing_counter = cfg->crossing_counter0 ? cfg->crossing_counter0[anchor] : 0;
    o->st.previous_milestone = -1;
    size_t np = (size_t) cfg->padded_num_atoms;
    o->posDelta = calloc(np, mixed4_size(o->precision));
    o->oldDelta = calloc(np, mixed4_size(o->precision));
    o->oldPosq = calloc(np, real4_size(o->precision));
    o->oldPosqCorrection = calloc(np, real4_size(o->precision));
    o->oldVelm = calloc(np, mixed4_size(o->precision));
    o->startVelm = calloc(np, mixed4_size(o->precision));
    o->startForce = calloc(3 * np, sizeof(long long));
    o->params[0] = o->params[1] = o->params[2] = 0.0;
    return o;
}

/* CudaSeekr2Kernels.cpp:188-203 */
void seekr2_oracle_set_params(seekr2_oracle* o, double temperature, double friction, double stepSize) {
    double kT = ORACLE_BOLTZ * temperature;
    double vscale = exp(-stepSize * friction);
    double fscale = (friction == 0 ? stepSize : (1 - vscale) / friction);
    double noisescale = sqrt(kT * (1 - vscale * vscale));
    o->params[0] = vscale; o->params[1] = fscale; o->params[2] = noisescale;
    o->temperature = temperature; o->friction = friction; o->stepSize = stepSize;
}
void seekr2_oracle_get_params(const seekr2_oracle* o, double out[3]) { memcpy(out, o->params, sizeof(double) * 3); }

/* ---- raw kernels, exposed so tests can compare each against oracle/_ref and the B200 kernels -- */

void seekr2_oracle_part1(seekr2_oracle* o, void* velm, const long long* force, void* posDelta,
                         const float* random4, unsigned int randomIndex) {
    switch (o->precision) {
    case SEEKR2_B200_SINGLE: part1_s(o->numAtoms, o->paddedNumAtoms, (mixed4_s*) velm, force, (mixed4_s*) posDelta, o->params, o->stepSize, random4, randomIndex); break;
    case SEEKR2_B200_MIXED:  part1_m(o->numAtoms, o->paddedNumAtoms, (mixed4_m*) velm, force, (mixed4_m*) posDelta, o->params, o->stepSize, random4, randomIndex); break;
    default:                 part1_d(o->numAtoms, o->paddedNumAtoms, (mixed4_d*) velm, force, (mixed4_d*) posDelta, o->params, o->stepSize, random4, randomIndex); break;
    }
}
void seekr2_oracle_part2(seekr2_oracle* o, void* posq, void* posqCorrection, const void* posDelta,
                         void* velm, void* oldPosq, void* oldVelm) {
    switch (o->precision) {
    case SEEKR2_B200_SINGLE: part2_s(o->numAtoms, (real4_s*) posq, (real4_s*) posqCorrection, (const mixed4_s*) posDelta, (mixed4_s*) velm, o->stepSize, (real4_s*) oldPosq, (mixed4_s*) oldVelm); break;
    case SEEKR2_B200_MIXED:  part2_m(o->numAtoms, (real4_m*) posq, (real4_m*) posqCorrection, (const mixed4_m*) posDelta, (mixed4_m*) velm, o->stepSize, (real4_m*) oldPosq, (mixed4_m*) oldVelm); break;
    default:                 part2_d(o->numAtoms, (real4_d*) posq, (real4_d*) posqCorrection, (const mixed4_d*) posDelta, (mixed4_d*) velm, o->stepSize, (real4_d*) oldPosq, (mixed4_d*) oldVelm); break;
    }
}
void seekr2_oracle_bounce(seekr2_oracle* o, void* posq, void* velm, const void* oldPosq, const void* oldVelm) {
    switch (o->precision) {
    case SEEKR2_B200_SINGLE: bounce_s(o->numAtoms, (real4_s*) posq, (mixed4_s*) velm, (const real4_s*) oldPosq, (const mixed4_s*) oldVelm); break;
    case SEEKR2_B200_MIXED:  bounce_m(o->numAtoms, (real4_m*) posq, (mixed4_m*) velm, (const real4_m*) oldPosq, (const mixed4_m*) oldVelm); break;
    default:                 bounce_d(o->numAtoms, (real4_d*) posq, (mixed4_d*) velm, (const real4_d*) oldPosq, (const mixed4_d*) oldVelm); break;
    }
}

/* ---- boundary value: stands in for context.calcForcesAndEnergy(false, true, 1<<forceGroup) ---- */

static void atom_position(const seekr2_oracle* o, const void* posq, const void* corr, int i, double out[3]) {
    switch (o->precision) {
    case SEEKR2_B200_SINGLE: position_s((const real4_s*) posq, i, out); break;
    case SEEKR2_B200_MIXED:  position_m((const real4_m*) posq, i, out); break;
    default:                 position_d((const real4_d*) posq, i, out); break;
    }
}

void seekr2_oracle_centroids(const seekr2_oracle* o, const void* posq, const void* corr, double* out /* [numGroups][3] */) {
    for (int g = 0; g < o->numGroups; g++) {
        long long acc[3] = { 0, 0, 0 };
        for (int j = o->groupOffsets[g]; j < o->groupOffsets[g + 1]; j++) {
            double p[3];
            atom_position(o, posq, corr, o->groupAtoms[j], p);
            for (int c = 0; c < 3; c++) acc[c] += llrint((o->groupWeights[j] * p[c]) * FIXED_SCALE);
        }
        for (int c = 0; c < 3; c++) out[3 * g + c] = (double) acc[c] * FIXED_INV_SCALE;
    }
}

static double sqdist(const double* ca, const double* cb) {
    double d0 = ca[0] - cb[0], d1 = ca[1] - cb[1], d2 = ca[2] - cb[2];
    return fma(d2, d2, fma(d1, d1, d0 * d0));
}

static double boundary_term(const seekr2_b200_boundary* b, const double* cen) {
    double u;
    const double* ca = cen + 3 * b->groups[0];
    if (b->type == SEEKR2_B200_CV_PLANE) {
        u = ca[b->axis] - b->threshold;
    } else if (b->type == SEEKR2_B200_CV_PAIR_SUM) {
        /* demo4:27: distance(g1,g2)^2 + distance(g3,g4)^2 + distance(g5,g6)^2 - radius^2, summed left to right */
        double sum = sqdist(cen + 3 * b->groups[0], cen + 3 * b->groups[1]);
        for (int p = 1; p < b->axis; p++) sum = sum + sqdist(cen + 3 * b->groups[2 * p], cen + 3 * b->groups[2 * p + 1]);
        u = sum - b->threshold * b->threshold;
    } else if (b->type == SEEKR2_B200_CV_ANGLE) {
        /* demo5:24: angle(g1,g2,g3) - ref_angle, through its cosine (center[0] = cos(ref_angle)) */
        const double* c1 = cen + 3 * b->groups[1];
        const double* c2 = cen + 3 * b->groups[2];
        double a[3], v[3];
        for (int c = 0; c < 3; c++) { a[c] = ca[c] - c1[c]; v[c] = c2[c] - c1[c]; }
        double dot = fma(a[2], v[2], fma(a[1], v[1], a[0] * v[0]));
        double na2 = fma(a[2], a[2], fma(a[1], a[1], a[0] * a[0]));
        double nb2 = fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]));
        u = b->center[0] * sqrt(na2 * nb2) - dot;
    } else {
        double d[3];
        if (b->type == SEEKR2_B200_CV_SPHERE_PAIR) {
            const double* cb = cen + 3 * b->groups[1];
            for (int c = 0; c < 3; c++) d[c] = ca[c] - cb[c];
        } else {
            for (int c = 0; c < 3; c++) d[c] = ca[c] - b->center[c];
        }
        double d2 = fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0]));
        u = d2 - b->threshold * b->threshold;
    }
    double ku = b->k * u;
    if (b->style == SEEKR2_B200_STYLE_RAW) return ku;
    return ku >= 0 ? b->bitcode : 0.0;   /* step(x) = 0 if x < 0, 1 otherwise */
}

/* FLAG_CV_FLOAT: the boundary value the way OpenMM's CUDA platform evaluates it in single / mixed precision -- float weights,
   float centres summed in group order, sqrtf, float distance^2 - radius^2 (restated: tools/cv_float_contract.py, "sequential") */
static double group_value_float(const seekr2_oracle* o, const void* posq, int forceGroup) {
    float cen[3 * SEEKR2_B200_MAX_GROUPS];
    const real4_s* q = (const real4_s*) posq;
    for (int g = 0; g < o->numGroups; g++) {
        float acc[3] = { 0.0f, 0.0f, 0.0f };
        for (int j = o->groupOffsets[g]; j < o->groupOffsets[g + 1]; j++) {
            const float w = (float) o->groupWeights[j];
            const real4_s p = q[o->groupAtoms[j]];
            acc[0] = acc[0] + w * p.x; acc[1] = acc[1] + w * p.y; acc[2] = acc[2] + w * p.z;     /* -ffp-contract=off: no fma */
        }
        for (int c = 0; c < 3; c++) cen[3 * g + c] = acc[c];
    }
    double value = 0.0;
    for (int i = 0; i < o->numBoundaries; i++) {
        const seekr2_b200_boundary* b = &o->boundaries[i];
        if (b->force_group != forceGroup) continue;
        const float* ca = cen + 3 * b->groups[0];
        float u;
        if (b->type == SEEKR2_B200_CV_PLANE) {
            u = ca[b->axis] - (float) b->threshold;
        } else {
            float d[3];
            for (int c = 0; c < 3; c++) d[c] = ca[c] - (b->type == SEEKR2_B200_CV_SPHERE_PAIR ? cen[3 * b->groups[1] + c] : (float) b->center[c]);
            const float r2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
            const float dist = sqrtf(r2);
            const float r = (float) b->threshold;
            u = dist * dist - r * r;
        }
        const float ku = (float) b->k * u;
        value += b->style == SEEKR2_B200_STYLE_RAW ? (double) ku : (ku >= 0.0f ? b->bitcode : 0.0);
    }
    return value;
}

double seekr2_oracle_group_value(const seekr2_oracle* o, const void* posq, const void* corr, int forceGroup) {
    if (o->flags & SEEKR2_B200_FLAG_CV_FLOAT) return group_value_float(o, posq, forceGroup);
    double cen[3 * SEEKR2_B200_MAX_GROUPS];
    seekr2_oracle_centroids(o, posq, corr, cen);
    double value = 0.0;
    for (int b = 0; b < o->numBoundaries; b++)
        if (o->boundaries[b].force_group == forceGroup) value += boundary_term(&o->boundaries[b], cen);
    return value;
}

/* ---- crossing records ---------------------------------------------------------------------- */

/* `State myState = context.getOwner().getState(State::Positions | State::Velocities)` at the moment of
   the crossing, i.e. after the position update and before the bounce (CudaSeekr2Kernels.cpp:283, :576):
   positions = posq (+ posqCorrection in mixed precision), velocities = velm.  Returns the sequence
   number (1, 2, ...) that the crossing's records carry. */
static int take_snapshot(seekr2_oracle* o, const void* posq, const void* corr, const void* velm) {
    const size_t per = 6 * (size_t) o->numAtoms;
    if (o->st.num_snapshots == o->capSnapshots) {
        o->capSnapshots = o->capSnapshots ? 2 * o->capSnapshots : 16;
        o->snapshots = (double*) realloc(o->snapshots, sizeof(double) * per * (size_t) o->capSnapshots);
    }
    double* x = o->snapshots + per * (size_t) o->st.num_snapshots;
    double* v = x + 3 * (size_t) o->numAtoms;
    for (int i = 0; i < o->numAtoms; i++) {
        if (o->precision == SEEKR2_B200_DOUBLE) {
            const real4_d* q = (const real4_d*) posq; x[3 * i] = q[i].x; x[3 * i + 1] = q[i].y; x[3 * i + 2] = q[i].z;
        } else {
            const real4_s* q = (const real4_s*) posq; x[3 * i] = q[i].x; x[3 * i + 1] = q[i].y; x[3 * i + 2] = q[i].z;
            if (o->precision == SEEKR2_B200_MIXED) {
                const real4_s* c = (const real4_s*) corr;
                x[3 * i] += (double) c[i].x; x[3 * i + 1] += (double) c[i].y; x[3 * i + 2] += (double) c[i].z;
            }
        }
        if (o->precision == SEEKR2_B200_SINGLE) {
            const mixed4_s* w = (const mixed4_s*) velm; v[3 * i] = w[i].x; v[3 * i + 1] = w[i].y; v[3 * i + 2] = w[i].z;
        } else {
            const mixed4_d* w = (const mixed4_d*) velm; v[3 * i] = w[i].x; v[3 * i + 1] = w[i].y; v[3 * i + 2] = w[i].z;
        }
    }
    return ++o->st.num_snapshots;
}
/* positions[numAtoms][3], velocities[numAtoms][3] of crossing state `snapshot` (1-based); 0 on success */
int seekr2_oracle_get_snapshot(const seekr2_oracle* o, int snapshot, double* positions, double* velocities) {
    if (snapshot < 1 || snapshot > o->st.num_snapshots) return SEEKR2_B200_ERR_INVALID;
    const size_t per = 6 * (size_t) o->numAtoms;
    const double* x = o->snapshots + per * (size_t) (snapshot - 1);
    if (positions) memcpy(positions, x, sizeof(double) * 3 * (size_t) o->numAtoms);
    if (velocities) memcpy(velocities, x + 3 * (size_t) o->numAtoms, sizeof(double) * 3 * (size_t) o->numAtoms);
    return SEEKR2_B200_OK;
}

static void push_record(seekr2_oracle* o, int kind, int index, int counter, int nsurf, double time) {
    if (o->numRecords == o->capRecords) {
        o->capRecords = o->capRecords ? 2 * o->capRecords : 1024;
        o->records = (seekr2_b200_record*) realloc(o->records, sizeof(seekr2_b200_record) * (size_t) o->capRecords);
    }
    seekr2_b200_record* r = &o->records[o->numRecords++];
    memset(r, 0, sizeof(*r));
    r->anchor = o->anchor; r->kind = kind; r->index = index; r->counter = counter;
    r->num_surfaces = nsurf; r->step = o->st.step_count; r->time = time;
    o->st.ring_head = o->numRecords;
}
int64_t seekr2_oracle_num_records(const seekr2_oracle* o) { return o->numRecords; }
void seekr2_oracle_get_records(const seekr2_oracle* o, seekr2_b200_record* out) {
    if (o->numRecords) memcpy(out, o->records, sizeof(seekr2_b200_record) * (size_t) o->numRecords);
}
void seekr2_oracle_get_state(const seekr2_oracle* o, seekr2_b200_anchor_state* out) {
    *out = o->st;
    for (int i = 0; i < o->numSrc; i++) out->elber_values[i] = (float) o->srcValues[i];
    for (int i = 0; i < o->numDest; i++) out->elber_values[o->numSrc + i] = (float) o->destValues[i];
}
void seekr2_oracle_set_time(seekr2_oracle* o, double time, int64_t stepCount) { o->st.time = time; o->st.step_count = stepCount; }

/* ---- the steps ------------------------------------------------------------------------------ */

/* LangevinMiddle: Part1, (velocity constraints), Part2, (constraints), Part3
   (CudaSeekr2Kernels.cpp:762-802; Elber twin :1057-1096) */
static void langevin_middle_parts(seekr2_oracle* o, void* posq, void* corr, void* velm, const long long* force,
                                  const float* random4, unsigned int randomIndex, int saveOld) {
    const double p2[2] = { o->params[0], o->params[2] };   /* {vscale, noisescale} (:741-747) */
    switch (o->precision) {
    case SEEKR2_B200_SINGLE:
        middle_part1_s(o->numAtoms, o->paddedNumAtoms, (mixed4_s*) velm, force, o->stepSize);
        middle_part2_s(o->numAtoms, (mixed4_s*) velm, (mixed4_s*) o->posDelta, (mixed4_s*) o->oldDelta, p2, o->stepSize, saveOld ? (mixed4_s*) o->oldVelm : NULL, random4, randomIndex);
        middle_part3_s(o->numAtoms, (real4_s*) posq, (mixed4_s*) velm, (const mixed4_s*) o->posDelta, (const mixed4_s*) o->oldDelta, o->stepSize, saveOld ? (real4_s*) o->oldPosq : NULL, NULL);
        break;
    case SEEKR2_B200_MIXED:
        middle_part1_m(o->numAtoms, o->paddedNumAtoms, (mixed4_m*) velm, force, o->stepSize);
        middle_part2_m(o->numAtoms, (mixed4_m*) velm, (mixed4_m*) o->posDelta, (mixed4_m*) o->oldDelta, p2, o->stepSize, saveOld ? (mixed4_m*) o->oldVelm : NULL, random4, randomIndex);
        middle_part3_m(o->numAtoms, (real4_m*) posq, (mixed4_m*) velm, (const mixed4_m*) o->posDelta, (const mixed4_m*) o->oldDelta, o->stepSize, saveOld ? (real4_m*) o->oldPosq : NULL, (real4_m*) corr);
        break;
    default:
        middle_part1_d(o->numAtoms, o->paddedNumAtoms, (mixed4_d*) velm, force, o->stepSize);
        middle_part2_d(o->numAtoms, (mixed4_d*) velm, (mixed4_d*) o->posDelta, (mixed4_d*) o->oldDelta, p2, o->stepSize, saveOld ? (mixed4_d*) o->oldVelm : NULL, random4, randomIndex);
        middle_part3_d(o->numAtoms, (real4_d*) posq, (mixed4_d*) velm, (const mixed4_d*) o->posDelta, (const mixed4_d*) o->oldDelta, o->stepSize, saveOld ? (real4_d*) o->oldPosq : NULL, NULL);
        break;
    }
}

static void langevin_two_parts(seekr2_oracle* o, void* posq, void* corr, void* velm, const long long* force,
                               const float* random4, unsigned int randomIndex, int saveOld) {
    if (o->scheme == SEEKR2_B200_SCHEME_MIDDLE) {
        langevin_middle_parts(o, posq, corr, velm, force, random4, randomIndex, saveOld);
        return;
    }
    seekr2_oracle_part1(o, velm, force, o->posDelta, random4, randomIndex);
    /* applyConstraints (CudaSeekr2Kernels.cpp:227): identity, the harness systems have none */
    seekr2_oracle_part2(o, posq, corr, o->posDelta, velm, saveOld ? o->oldPosq : NULL, saveOld ? o->oldVelm : NULL);
    /* computeVirtualSites (:239): none */
}

/* plain Langevin step: the comparator for the "<= 5% over a plain Langevin step" target */
int seekr2_oracle_langevin_step(seekr2_oracle* o, void* posq, void* corr, void* velm, const long long* force,
                                const float* random4, unsigned int randomIndex) {
    langevin_two_parts(o, posq, corr, velm, force, random4, randomIndex, 0);
    o->st.time += o->stepSize;
    o->st.step_count++;
    return SEEKR2_B200_OK;
}

/* CudaIntegrateMmvtLangevinStepKernel::execute (CudaSeekr2Kernels.cpp:175-366); with
   VARIANT_REFERENCE the differences of ReferenceSeekr2Kernels.cpp:195-339 listed in SURVEY.md 3.4. */
int seekr2_oracle_mmvt_step(seekr2_oracle* o, void* posq, void* corr, void* velm, const long long* force,
                            const float* random4, unsigned int randomIndex) {
    const int np = o->paddedNumAtoms;
    if (o->variant == SEEKR2_B200_VARIANT_REFERENCE) {
        /* the Reference platform copies positions, velocities and forces every step (:203-208);
           positions are saved by Part2 below, velocities and forces here */
        memcpy(o->startVelm, velm, mixed4_size(o->precision) * (size_t) np);
        memcpy(o->startForce, force, sizeof(long long) * 3 * (size_t) np);
    }
    if (o->st.step_count <= 0) {                                      /* :205-210 */
        float value0 = (float) seekr2_oracle_group_value(o, posq, corr, 1);
        o->st.last_value_positive = value0 > 0.0f;
        if (value0 > 0.0f) return SEEKR2_B200_ERR_FIRST_STEP;
    }
    if (o->precision == SEEKR2_B200_MIXED && (o->flags & SEEKR2_B200_FLAG_RESTORE_CORRECTION))
        memcpy(o->oldPosqCorrection, corr, real4_size(o->precision) * (size_t) np);
    langevin_two_parts(o, posq, corr, velm, force, random4, randomIndex, 1);   /* :214-239 */

    /* :248 with `float value` (:183); the Reference platform keeps a double (:201), which holds the
       same bit string for any value below 2^24 */
    float value = (float) seekr2_oracle_group_value(o, posq, corr, 1);
    o->st.last_value_positive = value > 0.0f;
    int bounced = 0;
    if (value > 0.0f) {                                                        /* :250 */
        bounced = 1;
        int bitcode = (int) value;                                             /* :262 */
        int num_bounced_surfaces = 0;
        for (int i = 0; i < o->numMilestoneGroups; i++) {                      /* :264-270 */
            if ((bitcode % 2) != 0) num_bounced_surfaces++;
            bitcode = bitcode >> 1;
        }
        bitcode = (int) value;
        for (int i = 0; i < o->numMilestoneGroups; i++) {                      /* :273-310 */
            if ((bitcode % 2) == 0) { bitcode = bitcode >> 1; continue; }
            bitcode = bitcode >> 1;
            push_record(o, SEEKR2_B200_REC_MMVT, i, o->st.bounce_counter, num_bounced_surfaces, o->st.time);  /* :281 */
            if (o->saveStates && num_bounced_surfaces == 1)                    /* :282-293 */
                o->records[o->numRecords - 1].snapshot = take_snapshot(o, posq, corr, velm);
            if (o->variant == SEEKR2_B200_VARIANT_REFERENCE || o->st.previous_milestone != -1)   /* :294-296 vs Reference :290 */
                o->st.N_alpha_beta[i] += 1;
            if (o->st.previous_milestone != i) {                               /* :297-305 */
                if (o->st.previous_milestone != -1) {
                    o->st.Nij_alpha[o->st.previous_milestone * SEEKR2_B200_MAX_MILESTONES + i] += 1;
                    o->st.Ri_alpha[o->st.previous_milestone] += o->st.incubation_time;
                } else {
                    o->st.first_crossing_time = o->st.time;
                }
                o->st.incubation_time = 0.0;
            }
            o->st.T_alpha = o->st.time - o->st.first_crossing_time;            /* :306 */
            o->st.previous_milestone = i;
            o->st.bounce_counter++;
        }
    }
    if (bounced) {                                                             /* :351-358 */
        const void* vsrc = (o->variant == SEEKR2_B200_VARIANT_REFERENCE) ? o->startVelm : o->oldVelm;
        seekr2_oracle_bounce(o, posq, velm, o->oldPosq, vsrc);
        if (o->precision == SEEKR2_B200_MIXED && (o->flags & SEEKR2_B200_FLAG_RESTORE_CORRECTION))
            memcpy(corr, o->oldPosqCorrection, real4_size(o->precision) * (size_t) np);
        o->st.num_bounce_steps++;
    }
    o->st.time += o->stepSize;                                                 /* :362-364 */
    o->st.step_count++;
    o->st.incubation_time += o->stepSize;
    return SEEKR2_B200_OK;
}

/* CudaIntegrateElberLangevinStepKernel::execute (CudaSeekr2Kernels.cpp:465-595). */
int seekr2_oracle_elber_step(seekr2_oracle* o, void* posq, void* corr, void* velm, const long long* force,
                             const float* random4, unsigned int randomIndex) {
    langevin_two_parts(o, posq, corr, velm, force, random4, randomIndex, 0);   /* :493-515 */
    float value = 0.0f, oldvalue = 0.0f;
    int num_bounced_surfaces = 0;
    if (!o->st.end_simulation) {                                               /* :524 */
        int first_record = (int) o->numRecords;
        for (int i = 0; i < o->numSrc; i++) {                                  /* :526-549 */
            value = (float) seekr2_oracle_group_value(o, posq, corr, o->srcGroups[i]);
            if (o->srcValues[i] == -INFINITY) { o->srcValues[i] = value; o->st.elber_sampled |= 1 << i; }
            oldvalue = (float) o->srcValues[i];
            int timeNonZero = o->variant == SEEKR2_B200_VARIANT_REFERENCE ? 1 : (o->st.time != 0.0);  /* :533 vs Reference :472 */
            if (((value - oldvalue) != 0.0f) && timeNonZero) {
                if (o->endOnSrcMilestone) {
                    o->st.end_simulation = 1;
                    num_bounced_surfaces++;
                    push_record(o, SEEKR2_B200_REC_ELBER_SRC, i, o->st.crossing_counter, 0, o->st.time);   /* :540 */
                } else {
                    o->st.crossed_src = 1;
                    o->st.time = 0.0;                                          /* context.setTime(0.0) :545 */
                    o->st.clock_resets++;
                    o->srcValues[i] = value;
                }
            }
        }
        for (int i = 0; i < o->numDest; i++) {                                 /* :551-572 */
            value = (float) seekr2_oracle_group_value(o, posq, corr, o->destGroups[i]);
            if (o->destValues[i] == -INFINITY) { o->destValues[i] = value; o->st.elber_sampled |= 1 << (o->numSrc + i); }
            oldvalue = (float) o->destValues[i];
            int timeNonZero = o->variant == SEEKR2_B200_VARIANT_REFERENCE ? 1 : (o->st.time != 0.0);  /* :558 vs Reference :498 */
            if (((value - oldvalue) != 0.0f) && timeNonZero) {
                o->st.end_simulation = 1;
                num_bounced_surfaces++;
                int star = !(o->st.crossed_src || o->endOnSrcMilestone);       /* :564-568 */
                push_record(o, star ? SEEKR2_B200_REC_ELBER_DEST_STAR : SEEKR2_B200_REC_ELBER_DEST, i,
                            o->st.crossing_counter, 0, o->st.time);
            }
        }
        if (o->st.end_simulation) {                                            /* :573-589 */
            int seq = (o->saveStates && num_bounced_surfaces == 1) ? take_snapshot(o, posq, corr, velm) : 0;   /* :575-587 */
            for (int r = first_record; r < (int) o->numRecords; r++) { o->records[r].num_surfaces = num_bounced_surfaces; o->records[r].snapshot = seq; }
            o->st.crossing_counter++;
        }
    }
    o->st.time += o->stepSize;                                                 /* :592-593 */
    o->st.step_count++;
    return SEEKR2_B200_OK;
}

/* Runs `steps` consecutive steps with random numbers taken at randomIndex0 + s*paddedNumAtoms
   (how prepareRandomNumbers hands out the pool) and forces refreshed by an optional harmonic tether
   F = -k (x - x0) (the harness force model); x0 == NULL keeps `force` fixed.
   Returns the status of the first failing step. */
int seekr2_oracle_run(seekr2_oracle* o, int steps, void* posq, void* corr, void* velm, long long* force,
                      const float* random4, unsigned int randomIndex0, const void* x0, double tetherK) {
    for (int s = 0; s < steps; s++) {
        if (x0) {
            for (int i = 0; i < o->numAtoms; i++) {
                double p[3];
                atom_position(o, posq, corr, i, p);
                double r0[3];
                if (o->precision == SEEKR2_B200_DOUBLE) { const real4_d* q = (const real4_d*) x0; r0[0] = q[i].x; r0[1] = q[i].y; r0[2] = q[i].z; }
                else { const real4_s* q = (const real4_s*) x0; r0[0] = q[i].x; r0[1] = q[i].y; r0[2] = q[i].z; }
                for (int c = 0; c < 3; c++)
                    force[i + c * o->paddedNumAtoms] = llrint((-tetherK * (p[c] - r0[c])) * 4294967296.0);
            }
        }
        unsigned int ri = randomIndex0 + (unsigned int) s * (unsigned int) o->paddedNumAtoms;
        int rc;
        if (o->kind == SEEKR2_B200_KIND_MMVT) rc = seekr2_oracle_mmvt_step(o, posq, corr, velm, force, random4, ri);
        else if (o->kind == SEEKR2_B200_KIND_ELBER) rc = seekr2_oracle_elber_step(o, posq, corr, velm, force, random4, ri);
        else rc = seekr2_oracle_langevin_step(o, posq, corr, velm, force, random4, ri);
        if (rc != SEEKR2_B200_OK) return rc;
    }
    return SEEKR2_B200_OK;
}

/* ---- crossing file / statistics file, formatted as the reference does ----------------------- */

/* CudaSeekr2Kernels.cpp:158-171 (MMVT) and :447-462 (Elber): header only if the file is absent */
int seekr2_oracle_write_header(int kind, const char* path) {
    FILE* probe = fopen(path, "r");
    if (probe) { fclose(probe); return 0; }
    FILE* f = fopen(path, "a");
    if (!f) return SEEKR2_B200_ERR_IO;
    if (kind == SEEKR2_B200_KIND_ELBER) {
        fputs("#\"Crossed boundary ID\",\"crossing counter\",\"total time (ps)\"\n", f);
        fputs("# An asterisk(*) indicates that source milestone was never crossed - asterisked statistics are invalid and should be excluded.\n", f);
    } else {
        fputs("#\"Bounced boundary ID\",\"bounce index\",\"total time (ps)\"\n", f);
    }
    fclose(f);
    return 0;
}

/* MMVT: fixed, precision(3) (:259-260,281).  Elber: default ostream formatting == %g (:540,565-567). */
int seekr2_oracle_append_log(const seekr2_oracle* o, const char* path, const int* labels) {
    FILE* f = fopen(path, "a");
    if (!f) return SEEKR2_B200_ERR_IO;
    for (int64_t r = 0; r < o->numRecords; r++) {
        const seekr2_b200_record* rec = &o->records[r];
        switch (rec->kind) {
        case SEEKR2_B200_REC_MMVT:            fprintf(f, "%d,%d,%.3f\n", labels[rec->index], rec->counter, rec->time); break;
        case SEEKR2_B200_REC_ELBER_SRC:       fprintf(f, "%d,%d,%g\n", labels[rec->index], rec->counter, rec->time); break;
        case SEEKR2_B200_REC_ELBER_DEST:      fprintf(f, "%d,%d,%g\n", labels[o->numSrc + rec->index], rec->counter, rec->time); break;
        default:                              fprintf(f, "%d*,%d,%g\n", labels[o->numSrc + rec->index], rec->counter, rec->time); break;
        }
    }
    fclose(f);
    return 0;
}

/* CudaSeekr2Kernels.cpp:325-343 */
int seekr2_oracle_write_statistics(const seekr2_oracle* o, const char* path, const int* labels) {
    FILE* f = fopen(path, "w");
    if (!f) return SEEKR2_B200_ERR_IO;
    int M = o->numMilestoneGroups;
    for (int i = 0; i < M; i++) fprintf(f, "N_alpha_%d: %d\n", labels[i], o->st.N_alpha_beta[i]);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++)
            fprintf(f, "N_%d_%d_alpha: %d\n", labels[i], labels[j], o->st.Nij_alpha[i * SEEKR2_B200_MAX_MILESTONES + j]);
    for (int i = 0; i < M; i++) fprintf(f, "R_%d_alpha: %.3f\n", labels[i], o->st.Ri_alpha[i]);
    fprintf(f, "T_alpha: %.3f\n", o->st.T_alpha);
    fclose(f);
    return 0;
}
