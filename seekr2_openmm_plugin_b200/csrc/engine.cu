// engine.cu -- host side of the C ABI declared in include/seekr2_b200.h.
//
// This is the B200 counterpart of the host code of the reference's CUDA platform kernels
// (seekr2plugin/platforms/cuda/src/CudaSeekr2Kernels.cpp:63-599): it owns the per-engine device
// state (group slab, per-anchor integrator state, crossing rings), turns T / gamma / dt into the
// kernel parameters, launches the step kernels on the caller's stream and formats the crossing and
// statistics files.  Unlike the reference there is no per-step host work besides the launch itself:
// no energy readback, no host branch for the bounce, no file I/O inside the step.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace s2;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            return fail(SEEKR2_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                  \
                        cudaGetErrorString(err__), __FILE__, __LINE__);                        \
    } while (0)

// OpenMM SimTKOpenMMRealType.h: BOLTZ = RGAS/KILO with RGAS = BOLTZMANN*AVOGADRO (CODATA 2018);
// the constant multiplied by the temperature at CudaSeekr2Kernels.cpp:191.
// auto mode (measured on 23 036-atom anchors, profiles/r02_local_regs_sweep.json; streaming blocks = anchors x ceil(atoms / 256)):
//   up to  4 x SMs blocks (<= 6 anchors)   LOCAL decision, 96 registers, 128 streaming threads per block
//   up to 12 x SMs blocks (<= 19 anchors)  LOCAL decision, 72 registers, 128 streaming threads per block
//   up to 30 x SMs blocks (<= 49 anchors)  LOCAL decision, 72 registers, 256 streaming threads per block
//   beyond                                 decide-block hand-off (HBM-bound: 48 registers, 1280 resident threads per SM)
constexpr int kLocalBlocksPerSm = 30, kLocalLatencyBlocksPerSm = 4, kLocalSmallBlockBlocksPerSm = 12;
constexpr long long kStageRecords = 1 << 16;   // staging buffer of the asynchronous drain (2.6 MB of mapped pinned memory)
constexpr double kBoltzmann = 1.380649e-23;
constexpr double kAvogadro = 6.02214076e23;
constexpr double kBoltz = (kBoltzmann * kAvogadro) / 1000.0;

size_t slab_entry_size(int precision) {
    switch (precision) {
    case SINGLE: return sizeof(SlabEntry<SINGLE>);
    case MIXED: return sizeof(SlabEntry<MIXED>);
    default: return sizeof(SlabEntry<DOUBLE>);
    }
}

}  // namespace

struct seekr2_b200_engine {
    seekr2_b200_config cfg{};
    EngineDev dev{};
    int device = 0;
    double params[3] = {0, 0, 0};   // vscale, fscale, noisescale
    double step_size = 0.0;
    bool params_set = false;
    int phase = 0;                  // fused launches so far mod 4: bit 0 = slab half holding the current group state / flag half,
                                    // the whole value = step_shadow slot the next launch reads
    bool slab_valid = false;
    bool use_pdl = true;            // fused step kernels are launched with programmatic stream serialization (SEEKR2_B200_PDL=0: off)
    bool local_decide = false;      // fused step uses fused_local_kernel (every block decides) instead of the decide-block hand-off
    int local_regs = 96;            // LOCAL mode: register budget of the instantiation that runs (kLocalRegsLatency / kLocalRegsThroughput)
    size_t local_dyn_smem = 0;      // LOCAL mode, MMVT: bytes of the undo log per block (dev.local_undo_tiles tiles)
    int local_grid_x = 0;           // LOCAL mode: blocks per anchor (0 = not sized yet; sized at the first launch from the occupancy of the
                                    // instantiation that runs, so that the whole grid is resident at once)
    int sm_count = 0;
    bool plain_next = true;         // the next fused launch is an ordinary one (no programmatic dependent launch): set whenever the step-index
                                    // shadow was (re)written by something other than the fused launch two steps back
    bool late_step_allowed = true;  // SEEKR2_B200_LATE_STEP=0: off (A/B knob)
    bool plain_by_resync = false;   // ... and that something was seekr2_b200_resync: with a Gaussian pool the fused launch may still overlap it
                                    // (programmatic dependent launch, step index read after the wait: dev.late_step)
    bool shadow_stale = false;      // two-pass steps advanced the clock but not the step-index shadow
    int finish_parity = 0;          // half of decision[2][A] the next finish launch publishes in (the other half is cleared by it)
    std::vector<int> group_atom_of_slot;   // slot -> atom index as given at create (seekr2_b200_set_atom_order maps it through the caller's order)
    long long* h_hint = nullptr;    // mapped pinned [2][A]: ring heads and clock resets as the device last wrote them (seekr2_b200_poll)
    long long drained_total = 0;    // records handed out by drain_end so far
    // device allocations
    int* d_slot_atom = nullptr;
    int* d_slot_group = nullptr;
    double* d_slot_weight = nullptr;
    Boundary* d_boundaries = nullptr;
    AnchorState* d_state = nullptr;
    Record* d_ring = nullptr;
    void* d_slab = nullptr;
    int* d_decision = nullptr;
    int* d_flags = nullptr;
    long long* d_step_shadow = nullptr;
    void* d_snap_posq = nullptr;
    void* d_snap_corr = nullptr;
    void* d_snap_velm = nullptr;
    // host mirrors for drain
    AnchorState* h_state = nullptr;           // pinned [A]
    // asynchronous drain: device-side tails, pack scratch, staging buffer + header in mapped pinned host memory
    long long* d_ring_tail = nullptr;         // [A] records already drained, per anchor
    long long* d_pack_scratch = nullptr;      // [3 A]
    Record* h_stage = nullptr;                // [kStageRecords], written by ring_pack_kernel over PCIe
    long long* h_stage_hdr = nullptr;         // [4]
    long long* h_stage_clock = nullptr;       // [A][2] time bits, step count of every anchor as of the last drain
    cudaEvent_t drain_event = nullptr;
    bool drain_pending = false;
    // graph capture
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool capturing = false;
    int capture_phase0 = 0;
    bool capture_plain0 = true;
    int capture_steps = 0;
    int capture_finishes = 0;
    int capture_finish_parity0 = 0;
    int graph_steps = 0;
};

namespace {

template <int P> Bufs<P> typed_bufs(const seekr2_b200_buffers* b) {
    Bufs<P> t;
    t.posq = (typename Prec<P>::real4*) b->d_posq;
    t.corr = (typename Prec<P>::real4*) b->d_posq_correction;
    t.velm = (typename Prec<P>::mixed4*) b->d_velm;
    t.force = b->d_force;
    t.pos_delta = (typename Prec<P>::mixed4*) b->d_pos_delta;
    t.random = (const float4*) b->d_random;
    t.random_anchor_stride = b->random_anchor_stride;
    t.old_velm = (typename Prec<P>::mixed4*) b->d_old_velm;
    t.old_delta = (typename Prec<P>::mixed4*) b->d_old_delta;
    return t;
}

// kernel parameters in the precision the reference kernels compute in (langevin.cu:13-16,71)
template <int P> StepParams<P> step_params(const seekr2_b200_engine* h) {
    using M = typename Prec<P>::mixed;
    StepParams<P> p;
    p.vscale = (M) h->params[0];
    p.fs = (M) h->params[1] / (M) 4294967296.0;
    p.noisescale = (M) h->params[2];
    p.dt = (M) h->step_size;
    p.inv_dt = 1.0 / (double) p.dt;
    p.fs_mid = p.dt / (M) 4294967296.0;      // langevinMiddle.cu:11
    p.halfdt = 0.5f * p.dt;                  // :37
    p.inv_dt_m = 1 / p.dt;                   // :73 (mixed precision, unlike langevin.cu:71)
    return p;
}

dim3 atom_grid(const seekr2_b200_engine* h) {
    return dim3((unsigned) ((h->cfg.num_atoms + kThreads - 1) / kThreads), (unsigned) h->cfg.num_anchors, 1);
}

int check_bufs(const seekr2_b200_engine* h, const seekr2_b200_buffers* b, bool need_random, bool need_delta) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    if (!b || !b->d_posq || !b->d_velm || !b->d_force) return fail(SEEKR2_B200_ERR_INVALID, "posq, velm and force buffers are required");
    if (h->cfg.precision == MIXED && !b->d_posq_correction) return fail(SEEKR2_B200_ERR_INVALID, "mixed precision needs posqCorrection");
    (void) need_random;   // d_random == NULL selects the in-kernel Philox generator
    if (need_delta && !b->d_pos_delta) return fail(SEEKR2_B200_ERR_INVALID, "posDelta is required for the two-pass path");
    if (!h->params_set) return fail(SEEKR2_B200_ERR_INVALID, "seekr2_b200_set_params has not been called");
    return SEEKR2_B200_OK;
}

template <int P> int launch_gather(seekr2_b200_engine* h, const seekr2_b200_buffers* b, cudaStream_t s) {
    if (h->dev.num_slots > 0) {
        dim3 grid((unsigned) ((h->dev.num_slots + kThreads - 1) / kThreads), (unsigned) h->cfg.num_anchors, 1);
        gather_groups_kernel<P><<<grid, kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), h->phase & 1);
    }
    evaluate_kernel<P><<<h->cfg.num_anchors, kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), step_params<P>(h));
    refresh_shadow_kernel<<<(h->cfg.num_anchors + 127) / 128, 128, 0, s>>>(h->dev, h->phase);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

template <int P, int KIND, int SCHEME, int REGS> int launch_local(seekr2_b200_engine* h, const seekr2_b200_buffers* b, unsigned ri, cudaStream_t s, bool pdl);

template <int P, int KIND, int SCHEME> int launch_fused(seekr2_b200_engine* h, const seekr2_b200_buffers* b, unsigned ri, cudaStream_t s) {
    const int bpa = (h->cfg.num_atoms + kThreads - 1) / kThreads;
    const bool pdl = h->use_pdl && !h->plain_next;
    if (h->local_decide)
    {
        if constexpr (KIND == SEEKR2_B200_KIND_LANGEVIN)
            if (h->local_regs == kLocalRegsPlain) return launch_local<P, KIND, SCHEME, kLocalRegsPlain>(h, b, ri, s, pdl);
        return h->local_regs == kLocalRegsLatency ? launch_local<P, KIND, SCHEME, kLocalRegsLatency>(h, b, ri, s, pdl)
                                                  : launch_local<P, KIND, SCHEME, kLocalRegsThroughput>(h, b, ri, s, pdl);
    }
    const unsigned grid = (unsigned) h->cfg.num_anchors * (unsigned) (bpa + 1);   // A decide blocks first, then the streaming blocks
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid, 1, 1);
    lc.blockDim = dim3(kThreads, 1, 1);
    lc.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    CUDA_TRY(cudaLaunchKernelEx(&lc, fused_step_kernel<P, KIND, SCHEME>, h->dev, typed_bufs<P>(b), step_params<P>(h), h->step_size, ri, h->phase, bpa));
    return SEEKR2_B200_OK;
}

template <int P, int KIND, int SCHEME, int REGS> int launch_local(seekr2_b200_engine* h, const seekr2_b200_buffers* b, unsigned ri, cudaStream_t s, bool pdl) {
    {
        // T streaming threads + the slot warps (none for plain Langevin)
        const int T = h->dev.local_stream_threads;
        const unsigned threads = (unsigned) T + (KIND == SEEKR2_B200_KIND_LANGEVIN ? 0u : (unsigned) h->dev.slot_threads);
        if (h->local_grid_x == 0) {
            // blocks per anchor: as many as are resident at once (each then walks ceil(tiles / G) tiles)
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_local_kernel<P, KIND, SCHEME, REGS>, (int) threads, 0));
            long long resident = (long long) (per_sm > 0 ? per_sm : 1) * h->sm_count;
            if (const char* g = getenv("SEEKR2_B200_LOCAL_RESIDENT")) resident = atoll(g);                          // tuning knob
            long long gx = resident / h->cfg.num_anchors;
            if (gx < 1) gx = 1;
            if (gx > h->dev.local_tiles) gx = h->dev.local_tiles;
            h->local_grid_x = (int) gx;
            // MMVT: the undo log in shared memory (kernels.cuh) -- one entry per streaming thread and tile a block walks, as many
            // tiles as fit without taking a resident block away from the SM
            h->dev.local_undo_tiles = 0;
            h->local_dyn_smem = 0;
            bool want_undo = KIND == SEEKR2_B200_KIND_MMVT && !(h->cfg.flags & SEEKR2_B200_FLAG_RESTORE_CORRECTION);
            if (const char* u = getenv("SEEKR2_B200_LOCAL_UNDO")) want_undo = want_undo && atoi(u) != 0;             // A/B knob
            if (want_undo) {
                const size_t per_tile = (size_t) T * (sizeof(typename Prec<P>::real4) + sizeof(typename Prec<P>::mixed4));
                const int need_per_sm = (int) ((gx * h->cfg.num_anchors + h->sm_count - 1) / h->sm_count);
                // measured (profiles/r02_local_undo_sweep.json): a block that walks ONE tile gains nothing from the log (the slot chain is
                // the step: 2.34 vs 2.40 us at one anchor -- it keeps that tile in registers and waits); two logged tiles are enough
                // for blocks of up to four tiles, one for longer walks (a round of tiles takes as long as the slot chain, and a
                // large log takes the L1 away from the streaming loads: 32 anchors 14.8 us with one logged tile, 16.6 with all)
                const int tpb = (int) ((h->dev.local_tiles + gx - 1) / gx);
                int U = tpb <= 1 ? 0 : tpb <= 4 ? std::min(tpb, 2) : 1;
                if (const char* u = getenv("SEEKR2_B200_LOCAL_UNDO_TILES")) U = std::min(tpb, atoi(u));              // A/B knob
                cudaFuncAttributes fa;
                CUDA_TRY(cudaFuncGetAttributes(&fa, fused_local_kernel<P, KIND, SCHEME, REGS>));
                int optin = 0;
                CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
                const size_t max_dyn = (size_t) optin > fa.sharedSizeBytes ? (size_t) optin - fa.sharedSizeBytes : 0;
                while (U > 0 && (size_t) U * per_tile > max_dyn) U--;
                // the attribute belongs to the function, not to this engine: only ever raise it (another engine of the process may be
                // launching the same instantiation with a larger log)
                if (U > 0 && (size_t) U * per_tile > (size_t) fa.maxDynamicSharedSizeBytes)
                    CUDA_TRY(cudaFuncSetAttribute(fused_local_kernel<P, KIND, SCHEME, REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ((size_t) U * per_tile)));
                for (; U > 0; U--) {
                    int occ = 0;
                    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_local_kernel<P, KIND, SCHEME, REGS>, (int) threads, (size_t) U * per_tile));
                    if (occ >= need_per_sm) break;
                }
                h->dev.local_undo_tiles = U;
                h->local_dyn_smem = (size_t) U * per_tile;
            }
        }
        // programmatic dependent launch: the next step's blocks may start their (static) prologue under this launch's body
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned) h->local_grid_x, (unsigned) h->cfg.num_anchors, 1);
        lc.blockDim = dim3(threads, 1, 1);
        lc.dynamicSmemBytes = h->local_dyn_smem;
        lc.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr;
        lc.numAttrs = pdl ? 1 : 0;
        CUDA_TRY(cudaLaunchKernelEx(&lc, fused_local_kernel<P, KIND, SCHEME, REGS>, h->dev, typed_bufs<P>(b), step_params<P>(h), h->step_size, ri, h->phase));
        return SEEKR2_B200_OK;
    }
}

template <int P> int launch_fused_kind(seekr2_b200_engine* h, const seekr2_b200_buffers* b, unsigned ri, cudaStream_t s) {
    if (h->cfg.scheme == SEEKR2_B200_SCHEME_MIDDLE) {
        switch (h->cfg.kind) {
        case SEEKR2_B200_KIND_MMVT: return launch_fused<P, SEEKR2_B200_KIND_MMVT, SEEKR2_B200_SCHEME_MIDDLE>(h, b, ri, s);
        case SEEKR2_B200_KIND_ELBER: return launch_fused<P, SEEKR2_B200_KIND_ELBER, SEEKR2_B200_SCHEME_MIDDLE>(h, b, ri, s);
        default: return launch_fused<P, SEEKR2_B200_KIND_LANGEVIN, SEEKR2_B200_SCHEME_MIDDLE>(h, b, ri, s);
        }
    }
    switch (h->cfg.kind) {
    case SEEKR2_B200_KIND_MMVT: return launch_fused<P, SEEKR2_B200_KIND_MMVT, SEEKR2_B200_SCHEME_LANGEVIN>(h, b, ri, s);
    case SEEKR2_B200_KIND_ELBER: return launch_fused<P, SEEKR2_B200_KIND_ELBER, SEEKR2_B200_SCHEME_LANGEVIN>(h, b, ri, s);
    default: return launch_fused<P, SEEKR2_B200_KIND_LANGEVIN, SEEKR2_B200_SCHEME_LANGEVIN>(h, b, ri, s);
    }
}

template <int P> int launch_kick(seekr2_b200_engine* h, const seekr2_b200_buffers* b, unsigned ri, cudaStream_t s) {
    if (h->cfg.scheme == SEEKR2_B200_SCHEME_MIDDLE)
        middle_part1_kernel<P><<<atom_grid(h), kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), step_params<P>(h));
    else
        kick_kernel<P><<<atom_grid(h), kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), step_params<P>(h), ri);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

template <int P> int launch_middle_drift(seekr2_b200_engine* h, const seekr2_b200_buffers* b, unsigned ri, cudaStream_t s) {
    middle_part2_kernel<P><<<atom_grid(h), kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), step_params<P>(h), ri);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

template <int P, int KIND> int launch_finish(seekr2_b200_engine* h, const seekr2_b200_buffers* b, cudaStream_t s) {
    // ONE launch: A decide blocks first (the pending position posq + corr + posDelta is the same expression in both schemes),
    // then the streaming blocks, which wait for their anchor's decision word of this launch's parity
    const int bpa = (h->cfg.num_atoms + kThreads - 1) / kThreads;
    const unsigned grid = (unsigned) h->cfg.num_anchors * (unsigned) (bpa + 1);
    if (h->cfg.scheme == SEEKR2_B200_SCHEME_MIDDLE)
        middle_finish_kernel<P, KIND><<<grid, kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), step_params<P>(h), h->step_size, h->finish_parity, bpa);
    else
        finish_kernel<P, KIND><<<grid, kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), step_params<P>(h), h->step_size, h->finish_parity, bpa);
    CUDA_TRY(cudaGetLastError());
    h->finish_parity ^= 1;
    if (h->capturing) h->capture_finishes++;
    return SEEKR2_B200_OK;
}

template <int P> int launch_resync(seekr2_b200_engine* h, const seekr2_b200_buffers* b, cudaStream_t s) {
    const int slots = h->dev.num_slots > 0 ? h->dev.num_slots : 1;
    dim3 grid((unsigned) ((slots + kThreads - 1) / kThreads), (unsigned) h->cfg.num_anchors, 1);
    resync_kernel<P><<<grid, kThreads, 0, s>>>(h->dev, typed_bufs<P>(b), h->phase);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

template <int P> int launch_finish_kind(seekr2_b200_engine* h, const seekr2_b200_buffers* b, cudaStream_t s) {
    switch (h->cfg.kind) {
    case SEEKR2_B200_KIND_MMVT: return launch_finish<P, SEEKR2_B200_KIND_MMVT>(h, b, s);
    case SEEKR2_B200_KIND_ELBER: return launch_finish<P, SEEKR2_B200_KIND_ELBER>(h, b, s);
    default: return launch_finish<P, SEEKR2_B200_KIND_LANGEVIN>(h, b, s);
    }
}

#define DISPATCH_PRECISION(h, expr_single, expr_mixed, expr_double) \
    ((h)->cfg.precision == SINGLE ? (expr_single) : (h)->cfg.precision == MIXED ? (expr_mixed) : (expr_double))

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int seekr2_b200_abi_version(void) { return SEEKR2_B200_ABI_VERSION; }
const char* seekr2_b200_last_error(void) { return g_last_error.c_str(); }

int seekr2_b200_destroy(seekr2_b200_handle h) {
    if (!h) return SEEKR2_B200_OK;
    DeviceGuard guard(h->device);   // the reference's destructors call cu.setAsCurrent() first (:74-80)
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    cudaFree(h->d_slot_atom); cudaFree(h->d_slot_group); cudaFree(h->d_slot_weight);
    cudaFree(h->d_boundaries); cudaFree(h->d_state); cudaFree(h->d_ring); cudaFree(h->d_slab);
    cudaFree(h->d_decision); cudaFree(h->d_flags); cudaFree(h->d_step_shadow);
    cudaFree(h->d_snap_posq); cudaFree(h->d_snap_corr); cudaFree(h->d_snap_velm);
    if (h->h_state) cudaFreeHost(h->h_state);
    cudaFree(h->d_ring_tail); cudaFree(h->d_pack_scratch);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->h_stage_hdr) cudaFreeHost(h->h_stage_hdr);
    if (h->h_stage_clock) cudaFreeHost(h->h_stage_clock);
    if (h->h_hint) cudaFreeHost(h->h_hint);
    if (h->drain_event) cudaEventDestroy(h->drain_event);
    delete h;
    return SEEKR2_B200_OK;
}

int seekr2_b200_create(const seekr2_b200_config* cfg, seekr2_b200_handle* out) {
    if (!cfg || !out) return fail(SEEKR2_B200_ERR_INVALID, "null config or output handle");
    *out = nullptr;
    if (cfg->abi_version != SEEKR2_B200_ABI_VERSION) return fail(SEEKR2_B200_ERR_INVALID, "ABI version mismatch: header %d, caller %d", SEEKR2_B200_ABI_VERSION, cfg->abi_version);
    if (cfg->precision < SINGLE || cfg->precision > DOUBLE) return fail(SEEKR2_B200_ERR_INVALID, "bad precision %d", cfg->precision);
    if (cfg->kind < SEEKR2_B200_KIND_LANGEVIN || cfg->kind > SEEKR2_B200_KIND_ELBER) return fail(SEEKR2_B200_ERR_INVALID, "bad integrator kind %d", cfg->kind);
    if (cfg->scheme != SEEKR2_B200_SCHEME_LANGEVIN && cfg->scheme != SEEKR2_B200_SCHEME_MIDDLE) return fail(SEEKR2_B200_ERR_INVALID, "bad scheme %d", cfg->scheme);
    if (cfg->variant != SEEKR2_B200_VARIANT_CUDA && cfg->variant != SEEKR2_B200_VARIANT_REFERENCE) return fail(SEEKR2_B200_ERR_INVALID, "bad variant %d", cfg->variant);
    if (cfg->num_anchors < 1 || cfg->num_anchors > 65535) return fail(SEEKR2_B200_ERR_INVALID, "num_anchors must be in [1, 65535]");
    if (cfg->num_atoms < 1 || cfg->padded_num_atoms < cfg->num_atoms) return fail(SEEKR2_B200_ERR_INVALID, "bad atom counts %d / %d", cfg->num_atoms, cfg->padded_num_atoms);
    if (cfg->num_groups < 0 || cfg->num_groups > SEEKR2_B200_MAX_GROUPS) return fail(SEEKR2_B200_ERR_INVALID, "at most %d centroid groups", SEEKR2_B200_MAX_GROUPS);
    if (cfg->num_boundaries < 0 || cfg->num_boundaries > SEEKR2_B200_MAX_BOUNDARIES) return fail(SEEKR2_B200_ERR_INVALID, "at most %d boundaries", SEEKR2_B200_MAX_BOUNDARIES);
    if (cfg->num_milestone_groups < 0 || cfg->num_milestone_groups > SEEKR2_B200_MAX_MILESTONES) return fail(SEEKR2_B200_ERR_INVALID, "at most %d milestone groups", SEEKR2_B200_MAX_MILESTONES);
    if (cfg->num_src < 0 || cfg->num_dest < 0 || cfg->num_src + cfg->num_dest > SEEKR2_B200_MAX_ELBER) return fail(SEEKR2_B200_ERR_INVALID, "at most %d Elber milestones", SEEKR2_B200_MAX_ELBER);
    if (cfg->num_groups > 0 && (!cfg->group_offsets || !cfg->group_atoms || !cfg->group_weights)) return fail(SEEKR2_B200_ERR_INVALID, "group arrays missing");
    if (cfg->num_boundaries > 0 && !cfg->boundaries) return fail(SEEKR2_B200_ERR_INVALID, "boundary array missing");

    // centroid groups: OpenMM normalises the weights of a group so that they sum to one
    std::vector<double> norm_weight;
    int real_slots = 0, aligned_slots = 0;
    for (int g = 0; g < cfg->num_groups; g++) {
        const int lo = cfg->group_offsets[g], hi = cfg->group_offsets[g + 1];
        if (hi <= lo) return fail(SEEKR2_B200_ERR_INVALID, "centroid group %d is empty", g);
        double sum = 0.0;
        for (int j = lo; j < hi; j++) sum += cfg->group_weights[j];
        if (!(sum > 0.0)) return fail(SEEKR2_B200_ERR_INVALID, "centroid group %d has non-positive total weight", g);
        for (int j = lo; j < hi; j++) {
            if (cfg->group_atoms[j] < 0 || cfg->group_atoms[j] >= cfg->num_atoms) return fail(SEEKR2_B200_ERR_INVALID, "group %d references atom %d outside [0,%d)", g, cfg->group_atoms[j], cfg->num_atoms);
            norm_weight.push_back(cfg->group_weights[j] / sum);
        }
        real_slots += hi - lo;
        aligned_slots += (hi - lo + 31) & ~31;
    }
    for (int a = 0; a < cfg->num_anchors; a++)
        for (int i = 0; i < cfg->num_boundaries; i++) {
            const seekr2_b200_boundary& bd = cfg->boundaries[(size_t) a * cfg->num_boundaries + i];
            if (bd.type < SEEKR2_B200_CV_SPHERE_PAIR || bd.type > SEEKR2_B200_CV_ANGLE) return fail(SEEKR2_B200_ERR_INVALID, "boundary %d: unknown type %d", i, bd.type);
            if (bd.type == SEEKR2_B200_CV_PLANE && (bd.axis < 0 || bd.axis > 2)) return fail(SEEKR2_B200_ERR_INVALID, "boundary %d: bad axis %d", i, bd.axis);
            if (bd.type == SEEKR2_B200_CV_PAIR_SUM && (bd.axis < 1 || bd.axis > 3)) return fail(SEEKR2_B200_ERR_INVALID, "boundary %d: PAIR_SUM needs 1..3 pairs, got %d", i, bd.axis);
            const int used = bd.type == SEEKR2_B200_CV_SPHERE_PAIR ? 2 : bd.type == SEEKR2_B200_CV_ANGLE ? 3 : bd.type == SEEKR2_B200_CV_PAIR_SUM ? 2 * bd.axis : 1;
            for (int g = 0; g < used; g++)
                if (bd.groups[g] < 0 || bd.groups[g] >= cfg->num_groups) return fail(SEEKR2_B200_ERR_INVALID, "boundary %d: bad group index %d", i, bd.groups[g]);
        }
    if (cfg->flags & SEEKR2_B200_FLAG_CV_FLOAT) {
        if (cfg->precision == DOUBLE) return fail(SEEKR2_B200_ERR_INVALID, "FLAG_CV_FLOAT is for SINGLE and MIXED precision (OpenMM evaluates in double in double precision)");
        for (size_t i = 0; i < (size_t) cfg->num_anchors * cfg->num_boundaries; i++)
            if (cfg->boundaries[i].type > SEEKR2_B200_CV_PLANE) return fail(SEEKR2_B200_ERR_INVALID, "FLAG_CV_FLOAT supports SPHERE_PAIR, SPHERE_FIXED and PLANE boundaries");
    }
    if (cfg->kind == SEEKR2_B200_KIND_ELBER) {
        // every milestone's force group must carry a Force (CudaSeekr2Kernels.cpp:415-436)
        for (int pass = 0; pass < 2; pass++) {
            const int n = pass == 0 ? cfg->num_src : cfg->num_dest;
            const int32_t* groups = pass == 0 ? cfg->src_groups : cfg->dest_groups;
            if (n > 0 && !groups) return fail(SEEKR2_B200_ERR_INVALID, "Elber milestone group array missing");
            for (int i = 0; i < n; i++) {
                bool found = false;
                for (int bnd = 0; bnd < cfg->num_boundaries; bnd++) found |= cfg->boundaries[bnd].force_group == groups[i];
                if (!found) return fail(SEEKR2_B200_ERR_NO_FORCE_GROUP, "System contains no force groups used to detect Elber boundary crossings. Check for mismatches between force group assignments and the groups added to the Elber integrator.");
            }
        }
    }

    seekr2_b200_engine* h = new (std::nothrow) seekr2_b200_engine();
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->cfg.group_offsets = nullptr; h->cfg.group_atoms = nullptr; h->cfg.group_weights = nullptr;
    h->cfg.boundaries = nullptr; h->cfg.bounce_counter0 = nullptr; h->cfg.src_groups = nullptr;
    h->cfg.dest_groups = nullptr; h->cfg.crossing_counter0 = nullptr;
    if (h->cfg.ring_capacity <= 0) h->cfg.ring_capacity = 4096;
    if (cfg->device >= 0) h->device = cfg->device; else cudaGetDevice(&h->device);
    DeviceGuard guard(h->device);

    const int A = cfg->num_anchors;
    EngineDev& d = h->dev;
    auto bail = [&](int code) { seekr2_b200_destroy(h); return code; };
    cudaDeviceProp prop;
    {
        const cudaError_t perr = cudaGetDeviceProperties(&prop, h->device);
        if (perr != cudaSuccess) return bail(fail(SEEKR2_B200_ERR_CUDA, "cudaGetDeviceProperties failed: %s", cudaGetErrorString(perr)));
    }
    h->sm_count = prop.multiProcessorCount;
    const int snap_slots = cfg->save_state_slots > 0 ? cfg->save_state_slots : 0;

    // Decision mode of the fused step.  LOCAL (every block evaluates the boundaries itself in its own slot warps, optimistic
    // stores, no hand-off through global memory) is for launches whose state lives in L2 -- one to a few tens of anchors per
    // GPU, bound by latency and instruction issue; the decide-block hand-off is for batched, HBM-bound launches, where the
    // redundant slot warps and the larger register budget would cost bandwidth.
    bool local;
    int local_threads = kThreads;
    {
        const long long stream_blocks = (long long) A * ((cfg->num_atoms + kThreads - 1) / kThreads);
        const bool eligible = snap_slots == 0 && aligned_slots <= kThreads && !(cfg->flags & SEEKR2_B200_FLAG_CV_FLOAT);
        long long limit = (long long) prop.multiProcessorCount * kLocalBlocksPerSm;
        if (const char* lm = getenv("SEEKR2_B200_LOCAL_LIMIT")) limit = atoll(lm);                                   // tuning knob
        local = eligible && stream_blocks <= limit;
        if (const char* m = getenv("SEEKR2_B200_DECIDE")) {                                                         // A/B knob
            if (m[0] == 'l') local = eligible;
            if (m[0] == 'h') local = false;
        }
        if (cfg->flags & SEEKR2_B200_FLAG_DECIDE_HANDOFF) local = false;
        if (cfg->flags & SEEKR2_B200_FLAG_DECIDE_LOCAL) {
            if (!eligible) return bail(fail(SEEKR2_B200_ERR_INVALID, "FLAG_DECIDE_LOCAL needs save_state_slots == 0 and at most %d group slots (have %d)", kThreads, aligned_slots));
            local = true;
        }
        h->local_decide = local;
        // register budget and block size of the LOCAL kernel: latency regime (one round of tiles) vs throughput regime
        const bool latency = stream_blocks <= (long long) prop.multiProcessorCount * kLocalLatencyBlocksPerSm;
        h->local_regs = latency ? kLocalRegsLatency : cfg->kind == SEEKR2_B200_KIND_LANGEVIN ? kLocalRegsPlain : kLocalRegsThroughput;
        if (stream_blocks <= (long long) prop.multiProcessorCount * kLocalSmallBlockBlocksPerSm) local_threads = 128;   // also the latency regime: host-guest size 2.51 vs 2.70 us (plain 1.48 vs 1.69), one trypsin anchor 2.27 vs 2.27 (plain 1.73 vs 1.85)
        if (cfg->kind == SEEKR2_B200_KIND_LANGEVIN && !latency) {
            // plain Langevin, multi-tile regime: 48 registers (more resident blocks) or the 64 ptxas takes under the 72 cap -- whichever
            // wastes fewer block slots in the last round of tiles (rounds x resident blocks; 12 anchors: 2 x 1480 against 2 x 1184,
            // measured 5.1 against 4.7 us; 8 anchors: 1 x 1480 against 2 x 1184, 3.2 against 3.9)
            const long long tiles = (long long) A * ((cfg->num_atoms + local_threads - 1) / local_threads);
            auto padded_work = [&](int regs) {
                const long long per_sm = std::min<long long>(65536 / ((long long) regs * local_threads), 2048 / local_threads);
                const long long resident = per_sm * prop.multiProcessorCount;
                return ((tiles + resident - 1) / resident) * resident;
            };
            h->local_regs = padded_work(kLocalRegsPlain) <= padded_work(64) ? kLocalRegsPlain : kLocalRegsThroughput;
        }
        {   // one or two anchors of a large system: 192-thread tiles fill the GPU with one or two blocks per SM -- fewer, evenly spread
            // blocks, and fewer slot warps asking L2 for the same lines (23 036 atoms: one anchor 1.94 vs 2.07 us, two 2.35 vs 2.61;
            // profiles/r02_local_threads_sweep.json)
            const long long b192 = (long long) A * ((cfg->num_atoms + 191) / 192);
            if (2 * b192 >= prop.multiProcessorCount && b192 <= 2LL * prop.multiProcessorCount) local_threads = 192;
        }
        if (const char* lr = getenv("SEEKR2_B200_LOCAL_REGS")) h->local_regs = atoi(lr) == kLocalRegsThroughput ? kLocalRegsThroughput : atoi(lr) == kLocalRegsPlain ? kLocalRegsPlain : kLocalRegsLatency;   // A/B knob (48: plain Langevin only)
        if (const char* pd = getenv("SEEKR2_B200_PDL")) if (pd[0] == '0') h->use_pdl = false;                       // A/B knob
        if (const char* lt = getenv("SEEKR2_B200_LATE_STEP")) if (lt[0] == '0') h->late_step_allowed = false;       // A/B knob
    }

    // Group slots.  ALIGNED layout (hand-off kernel; LOCAL with larger groups): every group's slot range is rounded up to a
    // multiple of 32 with weight-0 copies of its last atom, so that each warp holds slots of ONE group only and the warp-level
    // reduction of the centroid sums never mixes groups.  PACKED layout (LOCAL mode, all group atoms together fit one warp):
    // the groups sit side by side in a single warp and are summed with masked reductions in registers.
    bool packed = local && real_slots <= 32 && cfg->num_groups <= kMaxPackedGroups;
    if (const char* pk = getenv("SEEKR2_B200_PACKED")) if (pk[0] == '0') packed = false;                            // A/B knob
    std::vector<int> slot_atom, slot_group;
    std::vector<double> slot_weight;
    for (int g = 0, nw = 0; g < cfg->num_groups; g++) {
        const int lo = cfg->group_offsets[g], hi = cfg->group_offsets[g + 1];
        const size_t first = slot_atom.size();
        for (int j = lo; j < hi; j++) {
            slot_atom.push_back(cfg->group_atoms[j]);
            slot_group.push_back(g);
            slot_weight.push_back(norm_weight[nw++]);
        }
        if (packed) continue;
        while (slot_atom.size() % 32 != 0) {
            slot_atom.push_back(cfg->group_atoms[hi - 1]);
            slot_group.push_back(g);
            slot_weight.push_back(0.0);
        }
        // a group that fits into one warp's 32 slots is summed by that warp alone: its leader stores the sums
        // instead of adding them atomically (cv_accumulate)
        if (slot_atom.size() - first == 32)
            for (size_t j = first; j < slot_atom.size(); j++) slot_group[j] |= kGroupSingleWarp;
    }
    const int num_slots = (int) slot_atom.size();
    for (int g = 0; g < SEEKR2_B200_MAX_GROUPS; g++) { d.group_slot_begin[g] = 0; d.group_slot_count[g] = 0; }
    for (int s = 0, g = -1; s < num_slots; s++) {
        const int gs = slot_group[s] & (kGroupSingleWarp - 1);
        if (gs != g) { g = gs; d.group_slot_begin[g] = s; d.group_slot_count[g] = cfg->group_offsets[g + 1] - cfg->group_offsets[g]; }
    }
    d.slot_layout = (packed || num_slots == 0) ? kLayoutPacked : kLayoutAligned;
    d.slot_threads = d.slot_layout == kLayoutPacked ? 32 : num_slots;
    {
        int T = local_threads;
        if (const char* lt = getenv("SEEKR2_B200_LOCAL_THREADS")) T = atoi(lt);                                      // tuning knob
        if (T < 32 || T > 512 || (T & 31)) return bail(fail(SEEKR2_B200_ERR_INVALID, "SEEKR2_B200_LOCAL_THREADS must be a multiple of 32 in [32, 512]"));
        d.local_stream_threads = T;
        d.local_tiles = (cfg->num_atoms + T - 1) / T;
    }
    d.tune = 0;
    d.spin_limit = 40000000u;                      // ~2.5 s of 64 ns polls
    if (const char* sl = getenv("SEEKR2_B200_SPIN_LIMIT")) { const long long v = atoll(sl); d.spin_limit = v <= 0 ? 0xffffffffu : (unsigned) (v > 0xfffffffeLL ? 0xfffffffeLL : v); }
    if (const char* tn = getenv("SEEKR2_B200_TUNE")) d.tune = atoi(tn);                                             // tuning knob
    d.num_atoms = cfg->num_atoms; d.padded_num_atoms = cfg->padded_num_atoms; d.num_anchors = A;
    d.num_slots = num_slots; d.num_groups = cfg->num_groups; d.num_boundaries = cfg->num_boundaries;
    d.num_milestone_groups = cfg->num_milestone_groups;
    d.variant = cfg->variant; d.flags_bits = cfg->flags;
    d.num_src = cfg->num_src; d.num_dest = cfg->num_dest; d.end_on_src = cfg->end_on_src_milestone ? 1 : 0;
    for (int i = 0; i < cfg->num_src; i++) d.src_groups[i] = cfg->src_groups[i];
    for (int i = 0; i < cfg->num_dest; i++) d.dest_groups[i] = cfg->dest_groups[i];
    d.ring_capacity = h->cfg.ring_capacity;

#define CREATE_TRY(call)                                                                            \
    do {                                                                                            \
        cudaError_t err__ = (call);                                                                 \
        if (err__ != cudaSuccess) {                                                                 \
            fail(SEEKR2_B200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(err__));          \
            return bail(SEEKR2_B200_ERR_CUDA);                                                      \
        }                                                                                           \
    } while (0)

    if (num_slots > 0) {
        CREATE_TRY(cudaMalloc(&h->d_slot_atom, sizeof(int) * num_slots));
        CREATE_TRY(cudaMalloc(&h->d_slot_group, sizeof(int) * num_slots));
        CREATE_TRY(cudaMalloc(&h->d_slot_weight, sizeof(double) * num_slots));
        CREATE_TRY(cudaMemcpy(h->d_slot_atom, slot_atom.data(), sizeof(int) * num_slots, cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(h->d_slot_group, slot_group.data(), sizeof(int) * num_slots, cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMemcpy(h->d_slot_weight, slot_weight.data(), sizeof(double) * num_slots, cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMalloc(&h->d_slab, slab_entry_size(cfg->precision) * 2 * (size_t) A * num_slots));
        CREATE_TRY(cudaMemset(h->d_slab, 0, slab_entry_size(cfg->precision) * 2 * (size_t) A * num_slots));
    }
    if (cfg->num_boundaries > 0) {
        const size_t nb = (size_t) A * cfg->num_boundaries;
        std::vector<Boundary> bounds(cfg->boundaries, cfg->boundaries + nb);
        for (Boundary& bd : bounds)
            if (bd.type == SEEKR2_B200_CV_ANGLE) bd.center[0] = cos(bd.threshold);   // compared through the cosine (no acos on the device)
        CREATE_TRY(cudaMalloc(&h->d_boundaries, sizeof(Boundary) * nb));
        CREATE_TRY(cudaMemcpy(h->d_boundaries, bounds.data(), sizeof(Boundary) * nb, cudaMemcpyHostToDevice));
    }
    d.mmvt_any_term = 1;
    d.has_rare_terms = 0;
    for (size_t i = 0; i < (size_t) A * cfg->num_boundaries; i++) {
        const Boundary& bd = cfg->boundaries[i];
        if (bd.force_group == 1 && !(bd.style == SEEKR2_B200_STYLE_STEP && bd.bitcode >= 1.0 && bd.bitcode <= 2147483648.0)) d.mmvt_any_term = 0;
        if (bd.type > SEEKR2_B200_CV_PLANE) d.has_rare_terms = 1;
    }
    // the common MMVT anchor -- inner and outer sphere around one pair of groups: the LOCAL kernel's slot warps then sum the
    // difference of the two centroids directly (kernels.cuh; both slot layouts)
    d.pair_fast = 0;
    if (local && cfg->num_boundaries > 0 && !(cfg->flags & SEEKR2_B200_FLAG_CV_FLOAT)) {
        const int ga = cfg->boundaries[0].groups[0], gb = cfg->boundaries[0].groups[1];
        bool ok = ga != gb, cmp = true;
        for (size_t i = 0; ok && i < (size_t) A * cfg->num_boundaries; i++) {
            const Boundary& bd = cfg->boundaries[i];
            ok = bd.type == SEEKR2_B200_CV_SPHERE_PAIR && ((bd.groups[0] == ga && bd.groups[1] == gb) || (bd.groups[0] == gb && bd.groups[1] == ga));
            // 2: the term's activity is read off a comparison of the squared distance with the squared radius -- exact when k (d^2 - r^2)
            // cannot underflow or overflow (d^2 - r^2 is 0 or at least an ulp of r^2 in magnitude) and nothing is NaN
            const double ak = fabs(bd.k), ar = fabs(bd.threshold);
            cmp = cmp && bd.style == SEEKR2_B200_STYLE_STEP && ak >= 0x1p-100 && ak <= 0x1p100 && ar >= 0x1p-100 && ar <= 0x1p100;
        }
        if (const char* pf = getenv("SEEKR2_B200_PAIR_FAST")) {                                                       // A/B knob
            if (pf[0] == '0') ok = false;
            if (pf[0] == '1') cmp = false;
        }
        d.pair_fast = ok ? (cmp ? 2 : 1) : 0;
        d.pair_ga = ga; d.pair_gb = gb;
    }
    CREATE_TRY(cudaMalloc(&h->d_state, sizeof(AnchorState) * A));
    CREATE_TRY(cudaMalloc(&h->d_ring, sizeof(Record) * (size_t) A * d.ring_capacity));
    CREATE_TRY(cudaMalloc(&h->d_decision, sizeof(int) * 2 * A));         // [2][A]: finish launches alternate halves
    CREATE_TRY(cudaMemset(h->d_decision, 0, sizeof(int) * 2 * A));
    CREATE_TRY(cudaMalloc(&h->d_flags, sizeof(int) * (2 * A + 2)));      // [2][A] flags + [2] head-start counters
    CREATE_TRY(cudaMemset(h->d_flags, 0, sizeof(int) * (2 * A + 2)));
    CREATE_TRY(cudaMalloc(&h->d_step_shadow, sizeof(long long) * 4 * A));
    {   // launch 0 reads slot 0 (step 0), launch 1 slot 1 (step 1); slots 2 and 3 are written by launches 0 and 1
        std::vector<long long> shadow(4 * (size_t) A, 0);
        for (int a = 0; a < A; a++) shadow[(size_t) A + a] = 1;
        CREATE_TRY(cudaMemcpy(h->d_step_shadow, shadow.data(), sizeof(long long) * 4 * A, cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaHostAlloc(&h->h_state, sizeof(AnchorState) * A, cudaHostAllocDefault));
    // initialize(): zero statistics, seed counters (CudaSeekr2Kernels.cpp:118-133,151-154,403-404)
    memset(h->h_state, 0, sizeof(AnchorState) * A);
    for (int a = 0; a < A; a++) {
        h->h_state[a].previous_milestone = -1;
        h->h_state[a].bounce_counter = cfg->bounce_counter0 ? cfg->bounce_counter0[a] : 0;
        h->h_state[a].crossing_counter = cfg->crossing_counter0 ? cfg->crossing_counter0[a] : 0;
    }
    CREATE_TRY(cudaMemcpy(h->d_state, h->h_state, sizeof(AnchorState) * A, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc(&h->d_ring_tail, sizeof(long long) * A));
    CREATE_TRY(cudaMemset(h->d_ring_tail, 0, sizeof(long long) * A));
    CREATE_TRY(cudaMalloc(&h->d_pack_scratch, sizeof(long long) * 3 * A));
    CREATE_TRY(cudaHostAlloc(&h->h_stage, sizeof(Record) * kStageRecords, cudaHostAllocMapped));
    CREATE_TRY(cudaHostAlloc(&h->h_stage_hdr, sizeof(long long) * 4, cudaHostAllocMapped));
    memset(h->h_stage_hdr, 0, sizeof(long long) * 4);
    CREATE_TRY(cudaHostAlloc(&h->h_stage_clock, sizeof(long long) * 2 * A, cudaHostAllocMapped));
    memset(h->h_stage_clock, 0, sizeof(long long) * 2 * A);
    CREATE_TRY(cudaEventCreateWithFlags(&h->drain_event, cudaEventDisableTiming));
    CREATE_TRY(cudaHostAlloc(&h->h_hint, sizeof(long long) * 2 * A, cudaHostAllocMapped));
    memset(h->h_hint, 0, sizeof(long long) * 2 * A);
    CREATE_TRY(cudaHostGetDevicePointer((void**) &d.hint, h->h_hint, 0));
    h->group_atom_of_slot = slot_atom;

    d.slot_atom = h->d_slot_atom; d.slot_group = h->d_slot_group; d.slot_weight = h->d_slot_weight;
    d.boundaries = h->d_boundaries; d.state = h->d_state; d.ring = h->d_ring; d.slab = h->d_slab;
    d.decision = h->d_decision; d.flags = h->d_flags; d.step_shadow = h->d_step_shadow;
    d.rng_seed = 0; d.rng_anchor_base = 0; d.rng_anchor_stride = 1; d.probe = nullptr;
    d.snap_slots = snap_slots;
    if (d.snap_slots > 0) {
        if (d.snap_slots >= (1 << 20)) return bail(fail(SEEKR2_B200_ERR_INVALID, "too many save_state_slots"));
        const size_t per = (size_t) A * d.snap_slots * cfg->padded_num_atoms;
        const size_t real4_b = cfg->precision == DOUBLE ? 32 : 16, mixed4_b = cfg->precision == SINGLE ? 16 : 32;
        CREATE_TRY(cudaMalloc(&h->d_snap_posq, per * real4_b));
        CREATE_TRY(cudaMalloc(&h->d_snap_velm, per * mixed4_b));
        if (cfg->precision == MIXED) CREATE_TRY(cudaMalloc(&h->d_snap_corr, per * real4_b));
    }
    d.snap_posq = h->d_snap_posq; d.snap_corr = h->d_snap_corr; d.snap_velm = h->d_snap_velm;
    d.head_start = h->d_flags + 2 * A;
    {
        // streaming blocks resident when the kernel starts: SMs x blocks per SM, minus the decide blocks
        d.first_wave_blocks = prop.multiProcessorCount * kFusedMinBlocks - A;
        if (d.first_wave_blocks < 0) d.first_wave_blocks = 0;
        if (const char* hs = getenv("SEEKR2_B200_HEAD_START")) if (hs[0] == '0') d.first_wave_blocks = 0;   // tuning knob
        const int slot_warps = ((num_slots < kThreads ? num_slots : kThreads) + 31) / 32;
        d.head_start_target = A * slot_warps;
        d.prefetch_ahead = d.first_wave_blocks;
        if (const char* pf = getenv("SEEKR2_B200_PREFETCH")) if (pf[0] == '0') d.prefetch_ahead = 0;               // tuning knob
    }
    *out = h;
    return SEEKR2_B200_OK;
#undef CREATE_TRY
}

int seekr2_b200_set_params(seekr2_b200_handle h, double temperature, double friction, double step_size) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "parameters cannot change inside a graph capture");
    // CudaSeekr2Kernels.cpp:191-194
    const double kT = kBoltz * temperature;
    const double vscale = exp(-step_size * friction);
    const double fscale = (friction == 0 ? step_size : (1 - vscale) / friction);
    const double noisescale = sqrt(kT * (1 - vscale * vscale));
    h->params[0] = vscale; h->params[1] = fscale; h->params[2] = noisescale;
    h->step_size = step_size;
    h->params_set = true;
    return SEEKR2_B200_OK;
}

int seekr2_b200_set_rng(seekr2_b200_handle h, uint64_t seed, uint32_t anchor_base, uint32_t anchor_id_stride) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "the RNG cannot change inside a graph capture");
    h->dev.rng_seed = seed;
    h->dev.rng_anchor_base = anchor_base;
    h->dev.rng_anchor_stride = anchor_id_stride ? anchor_id_stride : 1;
    return SEEKR2_B200_OK;
}

int seekr2_b200_debug_set_probe(seekr2_b200_handle h, void* d_probe) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    h->dev.probe = (long long*) d_probe;
    return SEEKR2_B200_OK;
}

int seekr2_b200_decision_mode(seekr2_b200_handle h) { return h && h->local_decide ? 1 : 0; }

int seekr2_b200_get_params(seekr2_b200_handle h, double out[3]) {
    if (!h || !out) return fail(SEEKR2_B200_ERR_INVALID, "null argument");
    memcpy(out, h->params, sizeof(double) * 3);
    return SEEKR2_B200_OK;
}

int seekr2_b200_sync_groups(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, void* stream) {
    if (int rc = check_bufs(h, bufs, false, false)) return rc;
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "sync_groups cannot be captured");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    int rc = DISPATCH_PRECISION(h, launch_gather<SINGLE>(h, bufs, s), launch_gather<MIXED>(h, bufs, s), launch_gather<DOUBLE>(h, bufs, s));
    if (rc == SEEKR2_B200_OK) {
        h->slab_valid = true;
        h->shadow_stale = false;     // launch_gather ends with refresh_shadow_kernel
        h->plain_next = true;        // ... an ordinary launch: the next fused launch must not read the shadow before it has completed
        h->plain_by_resync = false;
    }
    return rc;
}

int seekr2_b200_resync(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, void* stream) {
    if (int rc = check_bufs(h, bufs, false, false)) return rc;
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "resync cannot be captured");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    int rc = DISPATCH_PRECISION(h, launch_resync<SINGLE>(h, bufs, s), launch_resync<MIXED>(h, bufs, s), launch_resync<DOUBLE>(h, bufs, s));
    if (rc == SEEKR2_B200_OK) {
        h->slab_valid = true;
        h->shadow_stale = false;
        h->plain_next = true;        // the shadow was written by an ordinary launch: the next fused launch must not read it early ...
        h->plain_by_resync = true;   // ... unless it reads it late (pool mode)
    }
    return rc;
}

int seekr2_b200_set_atom_order(seekr2_b200_handle h, const int32_t* index_of_atom, void* stream) {
    if (!h || !index_of_atom) return fail(SEEKR2_B200_ERR_INVALID, "null argument");
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "set_atom_order cannot be captured");
    DeviceGuard guard(h->device);
    const size_t n = h->group_atom_of_slot.size();
    if (n == 0) return SEEKR2_B200_OK;
    std::vector<int> mapped(n);
    for (size_t s = 0; s < n; s++) {
        const int at = index_of_atom[h->group_atom_of_slot[s]];
        if (at < 0 || at >= h->cfg.num_atoms) return fail(SEEKR2_B200_ERR_INVALID, "atom order maps atom %d to index %d outside [0,%d)", h->group_atom_of_slot[s], at, h->cfg.num_atoms);
        mapped[s] = at;
    }
    // pageable source: the copy is staged before the call returns, and it is ordered behind the steps already queued on `stream`
    CUDA_TRY(cudaMemcpyAsync(h->d_slot_atom, mapped.data(), sizeof(int) * n, cudaMemcpyHostToDevice, (cudaStream_t) stream));
    h->slab_valid = false;           // the slab was gathered (and is indexed) in the old order
    return SEEKR2_B200_OK;
}

int seekr2_b200_poll(seekr2_b200_handle h, int64_t* records_pending, int64_t* clock_resets) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    const int A = h->cfg.num_anchors;
    const volatile long long* hint = h->h_hint;
    long long appended = 0, resets = 0;
    for (int a = 0; a < A; a++) { appended += hint[a]; resets += hint[A + a]; }
    if (records_pending) *records_pending = appended - h->drained_total;
    if (clock_resets) *clock_resets = resets;
    return SEEKR2_B200_OK;
}

int seekr2_b200_check_first_step(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, int32_t* out_flags, void* stream) {
    if (int rc = seekr2_b200_sync_groups(h, bufs, stream)) return rc;
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    const int A = h->cfg.num_anchors;
    CUDA_TRY(cudaMemcpyAsync(h->h_state, h->d_state, sizeof(AnchorState) * A, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    bool any = false;
    for (int a = 0; a < A; a++) {
        const int flag = h->h_state[a].last_value_positive;
        if (out_flags) out_flags[a] = flag;
        any |= flag != 0;
    }
    if (any && h->cfg.kind == SEEKR2_B200_KIND_MMVT)
        return fail(SEEKR2_B200_ERR_FIRST_STEP, "MMVT simulation bouncing on first step: the system will be trapped behind a boundary. Check and revise MMVT boundary definitions and/or atomic positions.");
    return SEEKR2_B200_OK;
}

int seekr2_b200_step(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, uint32_t random_index, void* stream) {
    if (int rc = check_bufs(h, bufs, true, false)) return rc;
    if (h->cfg.kind != SEEKR2_B200_KIND_LANGEVIN && h->dev.num_slots > 0 && !h->slab_valid)
        return fail(SEEKR2_B200_ERR_INVALID, "seekr2_b200_sync_groups must be called before the first fused step");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    if (h->shadow_stale) {
        // two-pass steps advanced the clock but not the step-index shadow the fused kernels read (every kind, with or without slots)
        if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "a fused step after two-pass steps cannot be the first thing captured: call seekr2_b200_sync_groups before seekr2_b200_graph_begin");
        refresh_shadow_kernel<<<(h->cfg.num_anchors + 127) / 128, 128, 0, s>>>(h->dev, h->phase);
        CUDA_TRY(cudaGetLastError());
        h->shadow_stale = false;
        h->plain_next = true;
        h->plain_by_resync = false;
    }
    // directly behind seekr2_b200_resync, with a Gaussian pool: overlap the resync kernel (it releases its dependents at once)
    h->dev.late_step = h->use_pdl && h->late_step_allowed && h->plain_next && h->plain_by_resync && bufs->d_random != nullptr && !h->capturing ? 1 : 0;
    if (h->dev.late_step) h->plain_next = false;
    int rc = DISPATCH_PRECISION(h, launch_fused_kind<SINGLE>(h, bufs, random_index, s),
                                launch_fused_kind<MIXED>(h, bufs, random_index, s),
                                launch_fused_kind<DOUBLE>(h, bufs, random_index, s));
    if (rc != SEEKR2_B200_OK) return rc;
    h->dev.late_step = 0;
    h->phase = (h->phase + 1) & 3;   // slab half and decision flag alternate every launch; the step-index slots cycle through four
    h->plain_next = false;
    h->plain_by_resync = false;
    if (h->capturing) h->capture_steps++;
    return SEEKR2_B200_OK;
}

int seekr2_b200_kick(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, uint32_t random_index, void* stream) {
    if (int rc = check_bufs(h, bufs, true, true)) return rc;
    const bool middle = h->cfg.scheme == SEEKR2_B200_SCHEME_MIDDLE;
    if (h->cfg.kind == SEEKR2_B200_KIND_MMVT && (h->cfg.variant == SEEKR2_B200_VARIANT_REFERENCE || middle) && !bufs->d_old_velm)
        return fail(SEEKR2_B200_ERR_INVALID, "multi-pass MMVT needs d_old_velm (VARIANT_REFERENCE and the MIDDLE scheme)");
    if (middle && !bufs->d_old_delta) return fail(SEEKR2_B200_ERR_INVALID, "the MIDDLE multi-pass path needs d_old_delta");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    return DISPATCH_PRECISION(h, launch_kick<SINGLE>(h, bufs, random_index, s), launch_kick<MIXED>(h, bufs, random_index, s),
                              launch_kick<DOUBLE>(h, bufs, random_index, s));
}

int seekr2_b200_middle_drift(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, uint32_t random_index, void* stream) {
    if (int rc = check_bufs(h, bufs, true, true)) return rc;
    if (h->cfg.scheme != SEEKR2_B200_SCHEME_MIDDLE) return fail(SEEKR2_B200_ERR_INVALID, "seekr2_b200_middle_drift is for the MIDDLE scheme only");
    if (!bufs->d_old_delta) return fail(SEEKR2_B200_ERR_INVALID, "the MIDDLE multi-pass path needs d_old_delta");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    return DISPATCH_PRECISION(h, launch_middle_drift<SINGLE>(h, bufs, random_index, s), launch_middle_drift<MIXED>(h, bufs, random_index, s),
                              launch_middle_drift<DOUBLE>(h, bufs, random_index, s));
}

int seekr2_b200_finish(seekr2_b200_handle h, const seekr2_b200_buffers* bufs, void* stream) {
    if (int rc = check_bufs(h, bufs, false, true)) return rc;
    if (h->cfg.scheme == SEEKR2_B200_SCHEME_MIDDLE && (!bufs->d_old_delta || (h->cfg.kind == SEEKR2_B200_KIND_MMVT && !bufs->d_old_velm)))
        return fail(SEEKR2_B200_ERR_INVALID, "the MIDDLE multi-pass path needs d_old_delta (and d_old_velm for MMVT)");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    h->slab_valid = false;   // the two-pass path does not maintain the slab ...
    h->shadow_stale = true;  // ... nor the step-index shadow of the fused kernels
    return DISPATCH_PRECISION(h, launch_finish_kind<SINGLE>(h, bufs, s), launch_finish_kind<MIXED>(h, bufs, s),
                              launch_finish_kind<DOUBLE>(h, bufs, s));
}

int seekr2_b200_graph_begin(seekr2_b200_handle h, void* stream) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "already capturing");
    DeviceGuard guard(h->device);
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
    // relaxed mode: only this library's launches go into the capture, and a cudaFree / cudaMalloc issued meanwhile by the same thread
    // (a garbage-collected engine being destroyed in a Python host, an allocator) must not invalidate it
    CUDA_TRY(cudaStreamBeginCapture((cudaStream_t) stream, cudaStreamCaptureModeRelaxed));
    h->capturing = true;
    h->capture_phase0 = h->phase;
    h->capture_steps = 0;
    h->capture_finishes = 0;
    h->capture_finish_parity0 = h->finish_parity;
    h->capture_plain0 = h->plain_next;
    h->plain_next = true;            // the graph's first step has no kernel of this graph in front of it
    h->plain_by_resync = false;
    return SEEKR2_B200_OK;
}

int seekr2_b200_graph_end(seekr2_b200_handle h, void* stream) {
    if (!h || !h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "not capturing");
    DeviceGuard guard(h->device);
    h->capturing = false;
    cudaGraph_t g = nullptr;
    CUDA_TRY(cudaStreamEndCapture((cudaStream_t) stream, &g));
    h->graph = g;
    h->phase = h->capture_phase0;        // the captured steps have not run yet
    h->plain_next = h->capture_plain0;
    h->graph_steps = h->capture_steps;
    h->finish_parity = h->capture_finish_parity0;
    if (h->capture_finishes & 1)
        return fail(SEEKR2_B200_ERR_INVALID, "a step graph must contain an even number of two-pass steps (the decision half is baked into each finish launch); got %d", h->capture_finishes);
    if (h->graph_steps & 3)
        return fail(SEEKR2_B200_ERR_INVALID, "a step graph must contain a multiple of 4 fused steps (the launch phase is baked into each launch); got %d", h->graph_steps);
    CUDA_TRY(cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
    return SEEKR2_B200_OK;
}

int seekr2_b200_graph_launch(seekr2_b200_handle h, void* stream) {
    if (!h || !h->graph_exec) return fail(SEEKR2_B200_ERR_INVALID, "no instantiated step graph");
    if (h->phase != h->capture_phase0) return fail(SEEKR2_B200_ERR_INVALID, "launch phase differs from the one the graph was captured with");
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaGraphLaunch(h->graph_exec, (cudaStream_t) stream));
    h->plain_next = false;
    return SEEKR2_B200_OK;
}

// Asynchronous drain (north star: "a device ring buffer drained asynchronously").  begin() only enqueues the pack
// kernel and an event on `stream` and returns; the caller may queue further steps behind it.  end() waits for that
// event alone -- not for anything queued later -- and hands out the records, already in host memory.
int seekr2_b200_drain_begin(seekr2_b200_handle h, int64_t max_records, void* stream) {
    if (!h) return fail(SEEKR2_B200_ERR_INVALID, "null engine handle");
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "drain cannot be captured");
    if (h->drain_pending) return fail(SEEKR2_B200_ERR_INVALID, "a drain is already in flight: call seekr2_b200_drain_end first");
    if (max_records < 0) return fail(SEEKR2_B200_ERR_INVALID, "negative max_records");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    Record* d_stage = nullptr;
    long long* d_hdr = nullptr;
    long long* d_clock = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer((void**) &d_stage, h->h_stage, 0));
    CUDA_TRY(cudaHostGetDevicePointer((void**) &d_hdr, h->h_stage_hdr, 0));
    CUDA_TRY(cudaHostGetDevicePointer((void**) &d_clock, h->h_stage_clock, 0));
    const long long limit = max_records < kStageRecords ? max_records : kStageRecords;
    ring_pack_kernel<<<1, 1024, 0, s>>>(h->dev, h->d_ring_tail, h->d_pack_scratch, d_stage, d_hdr, d_clock, limit);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->drain_event, s));
    h->drain_pending = true;
    return SEEKR2_B200_OK;
}

int seekr2_b200_drain_end(seekr2_b200_handle h, seekr2_b200_record* out, int64_t max_records, int64_t* out_n, int64_t* out_waiting) {
    if (!h || !out_n) return fail(SEEKR2_B200_ERR_INVALID, "null argument");
    if (!h->drain_pending) return fail(SEEKR2_B200_ERR_INVALID, "no drain in flight: call seekr2_b200_drain_begin first");
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaEventSynchronize(h->drain_event));
    h->drain_pending = false;
    const long long n = h->h_stage_hdr[0];
    *out_n = 0;
    if (out_waiting) *out_waiting = h->h_stage_hdr[2];
    if (n > max_records) return fail(SEEKR2_B200_ERR_INVALID, "record buffer smaller than the max_records given to seekr2_b200_drain_begin");
    if (n > 0) {
        if (!out) return fail(SEEKR2_B200_ERR_INVALID, "null record buffer");
        memcpy(out, h->h_stage, sizeof(Record) * (size_t) n);
    }
    *out_n = n;
    h->drained_total += n;
    if (h->h_stage_hdr[1]) return fail(SEEKR2_B200_ERR_RING_OVERFLOW, "crossing ring overflowed: records were lost; drain more often or raise ring_capacity");
    return SEEKR2_B200_OK;
}

int seekr2_b200_drain_clock(seekr2_b200_handle h, int32_t anchor, double* time, int64_t* step_count) {
    if (!h || anchor < 0 || anchor >= h->cfg.num_anchors) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    if (h->drain_pending) return fail(SEEKR2_B200_ERR_INVALID, "a drain is in flight: call seekr2_b200_drain_end first");
    if (time) memcpy(time, &h->h_stage_clock[2 * anchor], sizeof(double));
    if (step_count) *step_count = h->h_stage_clock[2 * anchor + 1];
    return SEEKR2_B200_OK;
}

int seekr2_b200_drain_ready(seekr2_b200_handle h) {
    if (!h) return -1;
    if (!h->drain_pending) return 1;
    DeviceGuard guard(h->device);
    const cudaError_t e = cudaEventQuery(h->drain_event);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    fail(SEEKR2_B200_ERR_CUDA, "cudaEventQuery failed: %s", cudaGetErrorString(e));
    return -1;
}

int seekr2_b200_drain(seekr2_b200_handle h, seekr2_b200_record* out, int64_t max_records, int64_t* out_n, void* stream) {
    if (!h || !out_n) return fail(SEEKR2_B200_ERR_INVALID, "null argument");
    *out_n = 0;
    const int rc = seekr2_b200_drain_begin(h, max_records, stream);
    if (rc != SEEKR2_B200_OK) return rc;
    return seekr2_b200_drain_end(h, out, max_records, out_n, nullptr);
}

int seekr2_b200_get_state(seekr2_b200_handle h, int32_t anchor, seekr2_b200_anchor_state* out, void* stream) {
    if (!h || !out || anchor < 0 || anchor >= h->cfg.num_anchors) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    CUDA_TRY(cudaMemcpyAsync(out, h->d_state + anchor, sizeof(AnchorState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SEEKR2_B200_OK;
}

int seekr2_b200_set_time(seekr2_b200_handle h, int32_t anchor, double time, int64_t step_count, void* stream) {
    if (!h || anchor < 0 || anchor >= h->cfg.num_anchors) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    if (h->capturing) return fail(SEEKR2_B200_ERR_INVALID, "set_time cannot be captured");
    DeviceGuard guard(h->device);
    // one tiny launch carrying the values as kernel parameters: stream-ordered, no host synchronisation, no staging buffer
    set_time_kernel<<<1, 1, 0, (cudaStream_t) stream>>>(h->dev, anchor, time, (long long) step_count, h->phase);
    CUDA_TRY(cudaGetLastError());
    h->plain_next = true;            // the step-index shadow was written by an ordinary launch
    h->plain_by_resync = false;
    return SEEKR2_B200_OK;
}

int seekr2_b200_fetch_snapshot(seekr2_b200_handle h, int32_t anchor, int32_t snapshot, double* positions, double* velocities, void* stream) {
    if (!h || anchor < 0 || anchor >= h->cfg.num_anchors || snapshot < 1 || (!positions && !velocities)) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    if (h->dev.snap_slots <= 0) return fail(SEEKR2_B200_ERR_INVALID, "the engine was created with save_state_slots = 0");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t) stream;
    AnchorState st;
    CUDA_TRY(cudaMemcpyAsync(&st, h->d_state + anchor, sizeof(AnchorState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (snapshot > st.num_snapshots) return fail(SEEKR2_B200_ERR_INVALID, "snapshot %d has not been taken yet", snapshot);
    if (st.num_snapshots - snapshot >= h->dev.snap_slots)
        return fail(SEEKR2_B200_ERR_SNAPSHOT_LOST, "crossing state %d of anchor %d was overwritten (%d taken since, %d slots)", snapshot, anchor, st.num_snapshots - snapshot, h->dev.snap_slots);
    const int N = h->cfg.num_atoms, Np = h->cfg.padded_num_atoms;
    const size_t off = ((size_t) anchor * h->dev.snap_slots + (snapshot - 1) % h->dev.snap_slots) * Np;
    const size_t real_b = h->cfg.precision == DOUBLE ? 8 : 4, mixed_b = h->cfg.precision == SINGLE ? 4 : 8;
    std::vector<char> x(N * 4 * real_b), c(h->cfg.precision == MIXED ? N * 4 * real_b : 0), v(N * 4 * mixed_b);
    CUDA_TRY(cudaMemcpyAsync(x.data(), (char*) h->d_snap_posq + off * 4 * real_b, x.size(), cudaMemcpyDeviceToHost, s));
    if (!c.empty()) CUDA_TRY(cudaMemcpyAsync(c.data(), (char*) h->d_snap_corr + off * 4 * real_b, c.size(), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(v.data(), (char*) h->d_snap_velm + off * 4 * mixed_b, v.size(), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int i = 0; i < N; i++)
        for (int k = 0; k < 3; k++) {
            if (positions) {
                double p = real_b == 8 ? ((double*) x.data())[4 * i + k] : (double) ((float*) x.data())[4 * i + k];
                if (!c.empty()) p += (double) ((float*) c.data())[4 * i + k];     // what getState(Positions) returns in mixed precision
                positions[3 * i + k] = p;
            }
            if (velocities) velocities[3 * i + k] = mixed_b == 8 ? ((double*) v.data())[4 * i + k] : (double) ((float*) v.data())[4 * i + k];
        }
    return SEEKR2_B200_OK;
}

int seekr2_b200_write_state_xml(const char* path, int32_t num_atoms, const double* positions, const double* velocities, double time, const double* box) {
    if (!path || num_atoms < 0 || (!positions && !velocities)) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    FILE* f = fopen(path, "w");                       // std::ios_base::trunc (:290)
    if (!f) return fail(SEEKR2_B200_ERR_IO, "cannot open %s", path);
    fprintf(f, "<?xml version=\"1.0\" ?>\n<State time=\"%.17g\" type=\"State\" version=\"1\">\n", time);
    static const double zero_box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double* bx = box ? box : zero_box;
    fprintf(f, "\t<PeriodicBoxVectors>\n");
    for (int r = 0; r < 3; r++) fprintf(f, "\t\t<%c x=\"%.17g\" y=\"%.17g\" z=\"%.17g\"/>\n", "ABC"[r], bx[3 * r], bx[3 * r + 1], bx[3 * r + 2]);
    fprintf(f, "\t</PeriodicBoxVectors>\n");
    if (positions) {
        fprintf(f, "\t<Positions>\n");
        for (int i = 0; i < num_atoms; i++) fprintf(f, "\t\t<Position x=\"%.17g\" y=\"%.17g\" z=\"%.17g\"/>\n", positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
        fprintf(f, "\t</Positions>\n");
    }
    if (velocities) {
        fprintf(f, "\t<Velocities>\n");
        for (int i = 0; i < num_atoms; i++) fprintf(f, "\t\t<Velocity x=\"%.17g\" y=\"%.17g\" z=\"%.17g\"/>\n", velocities[3 * i], velocities[3 * i + 1], velocities[3 * i + 2]);
        fprintf(f, "\t</Velocities>\n");
    }
    fprintf(f, "</State>\n");
    fclose(f);
    return SEEKR2_B200_OK;
}

// ---- files ------------------------------------------------------------------------------------

int seekr2_b200_write_header(int32_t kind, const char* path) {
    if (!path) return fail(SEEKR2_B200_ERR_INVALID, "null path");
    // "see whether the file already exists" (CudaSeekr2Kernels.cpp:158-171, :447-462)
    if (FILE* probe = fopen(path, "r")) { fclose(probe); return SEEKR2_B200_OK; }
    FILE* f = fopen(path, "a");
    if (!f) return fail(SEEKR2_B200_ERR_IO, "cannot open %s", path);
    if (kind == SEEKR2_B200_KIND_ELBER) {
        fputs("#\"Crossed boundary ID\",\"crossing counter\",\"total time (ps)\"\n", f);
        fputs("# An asterisk(*) indicates that source milestone was never crossed - asterisked statistics are invalid and should be excluded.\n", f);
    } else {
        fputs("#\"Bounced boundary ID\",\"bounce index\",\"total time (ps)\"\n", f);
    }
    fclose(f);
    return SEEKR2_B200_OK;
}

int seekr2_b200_append_records(int32_t kind, const char* path, const seekr2_b200_record* recs, int64_t n,
                               int32_t anchor, const int32_t* labels, int32_t num_labels, int32_t num_src) {
    if (!path || (n > 0 && (!recs || !labels))) return fail(SEEKR2_B200_ERR_INVALID, "null argument");
    (void) kind;
    FILE* f = fopen(path, "a");                       // std::ios_base::app (:258, :539, :563)
    if (!f) return fail(SEEKR2_B200_ERR_IO, "cannot open %s", path);
    for (int64_t r = 0; r < n; r++) {
        const seekr2_b200_record& rec = recs[r];
        if (rec.anchor != anchor) continue;
        const int li = (rec.kind == SEEKR2_B200_REC_ELBER_DEST || rec.kind == SEEKR2_B200_REC_ELBER_DEST_STAR) ? num_src + rec.index : rec.index;
        if (li < 0 || li >= num_labels) { fclose(f); return fail(SEEKR2_B200_ERR_INVALID, "record index %d has no label", li); }
        switch (rec.kind) {
        case SEEKR2_B200_REC_MMVT:   // fixed, precision(3) (:259-260,281)
            fprintf(f, "%d,%d,%.3f\n", labels[li], rec.counter, rec.time); break;
        case SEEKR2_B200_REC_ELBER_DEST_STAR:   // default ostream formatting == %g (:567)
            fprintf(f, "%d*,%d,%g\n", labels[li], rec.counter, rec.time); break;
        default:
            fprintf(f, "%d,%d,%g\n", labels[li], rec.counter, rec.time); break;
        }
    }
    fclose(f);
    return SEEKR2_B200_OK;
}

int seekr2_b200_write_statistics(const char* path, const seekr2_b200_anchor_state* st, const int32_t* labels, int32_t num_labels) {
    if (!path || !st || (num_labels > 0 && !labels) || num_labels > SEEKR2_B200_MAX_MILESTONES) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    FILE* f = fopen(path, "w");                       // ios_base::trunc (:326)
    if (!f) return fail(SEEKR2_B200_ERR_IO, "cannot open %s", path);
    for (int i = 0; i < num_labels; i++) fprintf(f, "N_alpha_%d: %d\n", labels[i], st->N_alpha_beta[i]);
    for (int i = 0; i < num_labels; i++)
        for (int j = 0; j < num_labels; j++)
            fprintf(f, "N_%d_%d_alpha: %d\n", labels[i], labels[j], st->Nij_alpha[i * SEEKR2_B200_MAX_MILESTONES + j]);
    for (int i = 0; i < num_labels; i++) fprintf(f, "R_%d_alpha: %.3f\n", labels[i], st->Ri_alpha[i]);
    fprintf(f, "T_alpha: %.3f\n", st->T_alpha);
    fclose(f);
    return SEEKR2_B200_OK;
}

// ---- harness ------------------------------------------------------------------------------------

int seekr2_b200_harness_tether_force(int32_t precision, int32_t num_anchors, int32_t num_atoms, int32_t padded,
                                     const void* d_posq, const void* d_posq_correction, const void* d_x0, double k,
                                     long long* d_force, void* stream) {
    (void) d_posq_correction;   // forces see posq only, as OpenMM's force kernels do
    if (!d_posq || !d_x0 || !d_force) return fail(SEEKR2_B200_ERR_INVALID, "null buffer");
    dim3 grid((unsigned) ((num_atoms + kThreads - 1) / kThreads), (unsigned) num_anchors, 1);
    cudaStream_t s = (cudaStream_t) stream;
    if (precision == DOUBLE) tether_force_kernel<DOUBLE><<<grid, kThreads, 0, s>>>(num_atoms, padded, (const double4*) d_posq, (const double4*) d_x0, k, d_force);
    else tether_force_kernel<SINGLE><<<grid, kThreads, 0, s>>>(num_atoms, padded, (const float4*) d_posq, (const float4*) d_x0, k, d_force);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

int seekr2_b200_harness_bond_force(int32_t precision, int32_t num_atoms, int32_t padded, const void* d_posq,
                                   const void* d_posq_correction, int32_t num_bonds, const int32_t* d_bond_atoms,
                                   const double* d_bond_params, long long* d_force, void* stream) {
    (void) d_posq_correction; (void) num_atoms;
    if (!d_posq || !d_force || (num_bonds > 0 && (!d_bond_atoms || !d_bond_params))) return fail(SEEKR2_B200_ERR_INVALID, "null buffer");
    cudaStream_t s = (cudaStream_t) stream;
    CUDA_TRY(cudaMemsetAsync(d_force, 0, sizeof(long long) * 3 * (size_t) padded, s));
    if (num_bonds > 0) {
        const int grid = (num_bonds + 127) / 128;
        if (precision == DOUBLE) bond_force_kernel<DOUBLE><<<grid, 128, 0, s>>>(padded, (const double4*) d_posq, num_bonds, d_bond_atoms, d_bond_params, d_force);
        else bond_force_kernel<SINGLE><<<grid, 128, 0, s>>>(padded, (const float4*) d_posq, num_bonds, d_bond_atoms, d_bond_params, d_force);
    }
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

int seekr2_b200_harness_fill_pool(void* d_random, int32_t num_anchors, int32_t padded, int64_t anchor_stride,
                                  int32_t num_steps, uint64_t seed, uint32_t anchor_base, uint32_t anchor_id_stride, uint64_t step0, void* stream) {
    if (!d_random || num_anchors < 1 || padded < 1 || num_steps < 0 || anchor_stride < (int64_t) num_steps * padded)
        return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    if (num_steps == 0) return SEEKR2_B200_OK;
    dim3 grid((unsigned) ((padded + kThreads - 1) / kThreads), (unsigned) num_anchors, (unsigned) (num_steps < 32 ? num_steps : 32));
    fill_pool_kernel<<<grid, kThreads, 0, (cudaStream_t) stream>>>((float4*) d_random, padded, anchor_stride, num_steps, seed, anchor_base, anchor_id_stride ? anchor_id_stride : 1, step0);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

int seekr2_b200_harness_check_sqrt(unsigned long long* d_mismatches, void* stream) {
    if (!d_mismatches) return fail(SEEKR2_B200_ERR_INVALID, "bad argument");
    static_assert(kThreads == 256, "2^32 patterns = 65536 blocks x 256 rounds x 256 threads");
    CUDA_TRY(cudaMemsetAsync(d_mismatches, 0, sizeof(unsigned long long), (cudaStream_t) stream));
    check_sqrt_kernel<<<65536, kThreads, 0, (cudaStream_t) stream>>>(d_mismatches);
    CUDA_TRY(cudaGetLastError());
    return SEEKR2_B200_OK;
}

}  // extern "C"
