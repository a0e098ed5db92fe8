/* OpenMM KernelImpls over the B200 C ABI: the replacement of
 * seekr2plugin/platforms/cuda/src/CudaSeekr2Kernels.cpp:63-1177.  OpenMM is absent here: compiled and run
 * against tests/openmm_shim only (oracle/build_shim.sh, tests/test_gpu_openmm_shim.py).
 *
 * Per step the reference does Part1 -> applyConstraints -> Part2 -> computeVirtualSites ->
 * calcForcesAndEnergy(false,true,mask) [blocking readback] -> host crossing block -> mmvtBounce.
 * Here, per execute():
 *   System without constraints and virtual sites (posDelta untouched between the passes):
 *       seekr2_b200_resync (tiny: boundary-group atoms -> slab) + seekr2_b200_step -- ONE pass over the atoms, the whole
 *       step (kick, drift, CV, crossing test, bounce, log, clock) in one launch;
 *   otherwise: seekr2_b200_kick -> applyConstraints -> seekr2_b200_finish (one launch: decision + Part2 + bounce).
 * execute() never waits for the device: the crossing records stay in the device ring, the step kernels report "something
 * was appended" through mapped host memory (seekr2_b200_poll, no CUDA call), and only then is a drain queued behind the
 * current step and collected by a later execute() once it has landed.  Blocking points -- where the files must be
 * complete and Context time exact -- are computeKineticEnergy() and the destructor.  Elber's device-side setTime(0)
 * (:545) reaches cu.setTime() with the drain that reports it.  SEEKR2_B200_ASYNC_DRAIN=0 restores the reference's
 * behaviour (a blocking drain inside every execute()); crossing states (saveStateFileName) always use it.
 * The first-step test (:205-210) and the Elber force-group check (:415-436) keep their exceptions and messages. */
#include "B200Seekr2Kernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaIntegrationUtilities.h"
#include <cstdlib>
#include <fstream>

using namespace OpenMM;
namespace Seekr2Plugin {

B200StepKernelBase::~B200StepKernelBase() {
    cu.setAsCurrent();                                            /* :74-80 */
    if (handle && asyncDrain) {
        try { flushDrain(); } catch (...) {}                      /* the crossings file is complete when the kernel goes away */
    }
    seekr2_b200_destroy(handle);
}

void B200StepKernelBase::check(int status) const {
    if (status != SEEKR2_B200_OK) throw OpenMMException(seekr2_b200_last_error());
}

void B200StepKernelBase::create(const System& system, seekr2_b200_config& cfg, int seed) {
    cu.getPlatformData().initializeContexts(system);               /* :106 */
    cu.setAsCurrent();
    cu.getIntegrationUtilities().initRandomNumberGenerator(seed);  /* :109 -- OpenMM's pool is used */
    boundarySet = recogniseBoundaries(system);
    slotIndex = boundarySet.groupAtoms;                             /* identity order until OpenMM reorders */
    /* one launch for the whole step needs posDelta untouched between the passes: no constraints, no virtual sites */
    fused = system.getNumConstraints() == 0;
    for (int i = 0; i < system.getNumParticles() && fused; i++) fused = !system.isVirtualSite(i);
    if (const char* f = getenv("SEEKR2_B200_ADAPTER_FUSED")) if (f[0] == '0') fused = false;    /* A/B knob */
    cfg.abi_version = SEEKR2_B200_ABI_VERSION;
    cfg.precision = cu.getUseDoublePrecision() ? SEEKR2_B200_DOUBLE : cu.getUseMixedPrecision() ? SEEKR2_B200_MIXED : SEEKR2_B200_SINGLE;
    cfg.variant = SEEKR2_B200_VARIANT_CUDA;
    /* OpenMM evaluates the boundary expressions in float in single / mixed precision; the engine's default is an
       order-independent double / fixed-point evaluation (1-2.5 decisions in 10^5 steps of a constantly bouncing run differ,
       DESIGN.md section 3).  SEEKR2_B200_CV_FLOAT=1 selects the engine's float restatement of OpenMM's evaluation instead. */
    if (const char* f = getenv("SEEKR2_B200_CV_FLOAT")) if (f[0] == '1' && !cu.getUseDoublePrecision()) cfg.flags |= SEEKR2_B200_FLAG_CV_FLOAT;
    cfg.device = cu.getDeviceIndex();
    cfg.num_anchors = 1;
    cfg.num_atoms = cu.getNumAtoms();
    cfg.padded_num_atoms = cu.getPaddedNumAtoms();
    cfg.num_groups = (int) boundarySet.groupOffsets.size() - 1;
    cfg.group_offsets = boundarySet.groupOffsets.data();
    cfg.group_atoms = boundarySet.groupAtoms.data();
    cfg.group_weights = boundarySet.groupWeights.data();
    cfg.num_boundaries = (int) boundarySet.boundaries.size();
    cfg.boundaries = boundarySet.boundaries.data();
    /* asynchronous drain unless crossing states are saved (their atom order is the one of the crossing step, and OpenMM
       may reorder atoms before a late collection) */
    asyncDrain = saveStateFileName.empty();
    if (const char* a = getenv("SEEKR2_B200_ASYNC_DRAIN")) asyncDrain = asyncDrain && a[0] != '0';
    kind = cfg.kind;
    /* the ring is drained in the step that crossed, so two slots are plenty */
    cfg.save_state_slots = saveStateFileName.empty() ? 0 : 2;
    check(seekr2_b200_create(&cfg, &handle));
}

seekr2_b200_buffers B200StepKernelBase::buffers() {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    seekr2_b200_buffers b{};
    b.d_posq = (void*) cu.getPosq().getDevicePointer();
    b.d_posq_correction = cu.getUseMixedPrecision() ? (void*) cu.getPosqCorrection().getDevicePointer() : nullptr;   /* :231 */
    b.d_velm = (void*) cu.getVelm().getDevicePointer();
    b.d_force = (const long long*) cu.getForce().getDevicePointer();
    b.d_pos_delta = (void*) integration.getPosDelta().getDevicePointer();
    b.d_random = (const void*) integration.getRandom().getDevicePointer();
    b.random_anchor_stride = 0;
    return b;
}

/* CudaSeekr2Kernels.cpp:187-203 (Middle :733-748): step size for OpenMM's constraint kernels, integration
   parameters for ours; runs before the first-step test, which needs them */
void B200StepKernelBase::updateParams(double temperature, double friction, double stepSize) {
    cu.getIntegrationUtilities().setNextStepSize(stepSize);
    if (temperature != prevTemp || friction != prevFriction || stepSize != prevStepSize) {
        check(seekr2_b200_set_params(handle, temperature, friction, stepSize));
        prevTemp = temperature; prevFriction = friction; prevStepSize = stepSize;
    }
}

/* cu.getAtomIndex()[i] = the particle held at buffer index i (CudaContext reorders atoms for its neighbour list,
   cu.reorderAtoms() at :365, also from setPositions): the boundary groups name particles, the engine wants buffer
   indices.  A handful of comparisons per step; a full remap only when the order really changed. */
void B200StepKernelBase::followAtomOrder() {
    const std::vector<int>& order = cu.getAtomIndex();
    bool same = true;
    for (size_t s = 0; s < slotIndex.size() && same; s++) same = order[slotIndex[s]] == boundarySet.groupAtoms[s];
    if (same) return;
    std::vector<int32_t> indexOf(order.size());
    for (size_t i = 0; i < order.size(); i++) indexOf[order[i]] = (int32_t) i;
    check(seekr2_b200_set_atom_order(handle, indexOf.data(), (void*) cu.getCurrentStream()));
    for (size_t s = 0; s < slotIndex.size(); s++) slotIndex[s] = indexOf[boundarySet.groupAtoms[s]];
}

void B200StepKernelBase::stepTwoPass(double temperature, double friction, double stepSize, double tolerance) {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    updateParams(temperature, friction, stepSize);
    followAtomOrder();
    seekr2_b200_buffers b = buffers();
    void* stream = (void*) cu.getCurrentStream();
    const int randomIndex = integration.prepareRandomNumbers(cu.getPaddedNumAtoms());       /* :214 */
    if (fused) {
        /* Part1 + Part2 + CV + crossing + bounce + log + clock in one launch (:214-365); the slab of boundary-group
           atoms is re-read first because anything may have touched posq / velm since the last step */
        check(seekr2_b200_resync(handle, &b, stream));
        check(seekr2_b200_step(handle, &b, (uint32_t) randomIndex, stream));
        return;
    }
    check(seekr2_b200_kick(handle, &b, (uint32_t) randomIndex, stream));                    /* Part1 :215-223 */
    integration.applyConstraints(tolerance);                                                /* :227 */
    check(seekr2_b200_finish(handle, &b, stream));                                          /* Part2 + CV + crossing + bounce :231-358 */
    integration.computeVirtualSites();                                                      /* :239 (boundary groups with virtual sites are refused at initialize) */
}

void B200StepKernelBase::allocateScratch(bool needOldVelm, bool needOldDelta) {
    const size_t elem = (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) ? 4 * sizeof(double) : 4 * sizeof(float);
    if (needOldVelm && !oldVelm.isInitialized()) oldVelm.initialize(cu, cu.getPaddedNumAtoms(), elem, "b200OldVelm");
    if (needOldDelta && !oldDelta.isInitialized()) oldDelta.initialize(cu, cu.getPaddedNumAtoms(), elem, "b200OldDelta");
}

/* CudaIntegrateMmvtLangevinMiddleStepKernel::execute, integration part (CudaSeekr2Kernels.cpp:756-802) */
void B200StepKernelBase::stepMiddle(double temperature, double friction, double stepSize, double tolerance) {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    updateParams(temperature, friction, stepSize);
    followAtomOrder();
    seekr2_b200_buffers b = buffers();
    if (fused) {                                                    /* no solver between the parts: oldDelta == delta */
        void* stream = (void*) cu.getCurrentStream();
        const int randomIndex = integration.prepareRandomNumbers(cu.getPaddedNumAtoms());   /* :762 */
        check(seekr2_b200_resync(handle, &b, stream));
        check(seekr2_b200_step(handle, &b, (uint32_t) randomIndex, stream));
        return;
    }
    b.d_old_velm = (void*) oldVelm.getDevicePointer();
    b.d_old_delta = (void*) oldDelta.getDevicePointer();
    void* stream = (void*) cu.getCurrentStream();
    const int randomIndex = integration.prepareRandomNumbers(cu.getPaddedNumAtoms());       /* :762 */
    check(seekr2_b200_kick(handle, &b, (uint32_t) randomIndex, stream));                    /* Part1 :763-767 */
    integration.applyVelocityConstraints(tolerance);                                        /* :770 */
    check(seekr2_b200_middle_drift(handle, &b, (uint32_t) randomIndex, stream));            /* Part2 :774-783 */
    integration.applyConstraints(tolerance);                                                /* :787 */
    check(seekr2_b200_finish(handle, &b, stream));                                          /* Part3 + CV + crossing + bounce :791-908 */
    integration.computeVirtualSites();                                                      /* :799 */
}

void B200StepKernelBase::drainTo(const std::string& file, int kind, const std::vector<int32_t>& labels, int numSrc) {
    drainFile = file; drainKind = kind; drainLabels = labels; drainNumSrc = numSrc;
    collected = false;
    if (asyncDrain) {
        /* collect the drain begun by an earlier step if it has landed (never waits) ... */
        const int64_t n = drainPending ? collectPending(false) : -1;
        if (n < 0) records.clear();                                /* nothing new this time */
        /* ... and queue one behind this step only if the device has reported something to fetch: records appended, or
           an Elber clock reset.  A plain read of mapped host memory -- on all but crossing steps execute() issues
           nothing besides the step itself. */
        if (!drainPending) {
            int64_t pending = 0, resets = 0;
            check(seekr2_b200_poll(handle, &pending, &resets));
            if (pending > 0 || resets != seenResets) {
                seenResets = resets;
                check(seekr2_b200_drain_begin(handle, 4096, (void*) cu.getCurrentStream()));
                drainPending = true;
                pendingClockValid = true;
            }
        }
        return;
    }
    records.resize(4096);
    int64_t n = 0;
    check(seekr2_b200_drain(handle, records.data(), (int64_t) records.size(), &n, (void*) cu.getCurrentStream()));
    noteClock(true);
    writeRecords(n);
}

int64_t B200StepKernelBase::collectPending(bool block) {
    if (!drainPending) return 0;
    if (!block) {
        const int ready = seekr2_b200_drain_ready(handle);
        if (ready < 0) check(SEEKR2_B200_ERR_CUDA);
        if (ready == 0) return -1;
    }
    records.resize(4096);
    int64_t n = 0, waiting = 0;
    check(seekr2_b200_drain_end(handle, records.data(), (int64_t) records.size(), &n, &waiting));
    drainPending = false;
    noteClock(pendingClockValid);
    writeRecords(n);
    return n;
}

void B200StepKernelBase::flushDrain() {
    if (!asyncDrain || drainFile.empty()) return;
    int64_t total = collectPending(true);
    for (;;) {                                                     /* whatever was appended after the last drain_begin */
        records.resize(4096);
        int64_t n = 0;
        check(seekr2_b200_drain(handle, records.data(), (int64_t) records.size(), &n, (void*) cu.getCurrentStream()));
        noteClock(true);
        writeRecords(n);
        total += n;
        if (n < (int64_t) 4096) break;
    }
    int64_t resets = 0;
    check(seekr2_b200_poll(handle, nullptr, &resets));
    seenResets = resets;
    if (total > 0 && kind == SEEKR2_B200_KIND_MMVT) writeStatistics();
}

/* Context time as the device holds it (Elber sets it back to 0 on the device, :545): the clock that came with the latest
   drain, advanced by the steps issued since (the same repeated addition the device does, so the bits agree).  Returns
   false when no usable clock has been collected. */
bool B200StepKernelBase::deviceTime(long long stepsIssued, double stepSize, double& time) {
    if (!collected || !clockValid) return false;
    time = clockTime;
    for (long long k = clockStep; k < stepsIssued; k++) time += stepSize;
    return true;
}

/* a drain has just been collected: keep its clock (the staging area belongs to the next drain from now on) */
void B200StepKernelBase::noteClock(bool valid) {
    collected = true;
    clockValid = valid;
    int64_t at = 0;
    check(seekr2_b200_drain_clock(handle, 0, &clockTime, &at));
    clockStep = at;
}

/* CudaSeekr2Kernels.cpp:313-345: rewritten whenever crossings have been logged */
void B200StepKernelBase::writeStatistics() {
    if (statisticsFile.empty()) return;
    seekr2_b200_anchor_state st;
    check(seekr2_b200_get_state(handle, 0, &st, (void*) cu.getCurrentStream()));
    check(seekr2_b200_write_statistics(statisticsFile.c_str(), &st, drainLabels.data(), (int32_t) drainLabels.size()));
}

void B200StepKernelBase::writeRecords(int64_t n) {
    const std::string& file = drainFile;
    const int kind = drainKind, numSrc = drainNumSrc;
    const std::vector<int32_t>& labels = drainLabels;
    if (n > 0) check(seekr2_b200_append_records(kind, file.c_str(), records.data(), n, 0, labels.data(), (int32_t) labels.size(), numSrc));
    records.resize((size_t) n);
    /* saveStateFileName: the crossed configuration, kept on the device before the bounce, goes to
       "<prefix>_<counter>_<milestone group>" (CudaSeekr2Kernels.cpp:282-293, Elber :575-587) */
    int32_t written = 0;
    for (const seekr2_b200_record& r : records) {
        if (r.snapshot <= 0 || r.snapshot == written) continue;     /* Elber: one state per ending step */
        written = r.snapshot;
        const int n3 = 3 * cu.getNumAtoms();
        std::vector<double> x(n3), v(n3), xo(n3), vo(n3);
        check(seekr2_b200_fetch_snapshot(handle, 0, r.snapshot, x.data(), v.data(), (void*) cu.getCurrentStream()));
        const std::vector<int>& order = cu.getAtomIndex();             /* undo OpenMM's atom reordering, as getState does */
        for (int i = 0; i < cu.getNumAtoms(); i++)
            for (int k = 0; k < 3; k++) { xo[3 * order[i] + k] = x[3 * i + k]; vo[3 * order[i] + k] = v[3 * i + k]; }
        const int label = labels[(r.kind == SEEKR2_B200_REC_ELBER_DEST || r.kind == SEEKR2_B200_REC_ELBER_DEST_STAR) ? numSrc + r.index : r.index];
        Vec3 a, b, c;
        cu.getPeriodicBoxVectors(a, b, c);
        const double box[9] = {a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]};
        /* "_<counter>_<group>" (:285-287, :578-580, :838-840); the reference's ElberLangevinMiddle kernel leaves
           the counter out (:1157) and so do we */
        const std::string name = saveStateFileName + (stateNameHasCounter ? "_" + std::to_string(r.counter) : std::string()) + "_" + std::to_string(label);
        check(seekr2_b200_write_state_xml(name.c_str(), cu.getNumAtoms(), xo.data(), vo.data(), r.time, box));
    }
}

/* ---- MMVT (CudaSeekr2Kernels.cpp:104-370; Middle :646-919) ------------------------------------------- */

template <class K, class I, bool MIDDLE>
void B200MmvtStepKernel<K, I, MIDDLE>::initialize(const System& system, const I& integrator) {
    seekr2_b200_config cfg{};
    cfg.kind = SEEKR2_B200_KIND_MMVT;
    cfg.scheme = MIDDLE ? SEEKR2_B200_SCHEME_MIDDLE : SEEKR2_B200_SCHEME_LANGEVIN;
    cfg.num_milestone_groups = integrator.getNumMilestoneGroups();
    const int32_t counter0 = integrator.getBounceCounter();        /* :151 */
    cfg.bounce_counter0 = &counter0;
    saveStateFileName = integrator.getSaveStateFileName();          /* :135-139 */
    requireTypedMilestoneForces(system, {1});                       /* the hard-wired mask 2 (:206,:248) */
    create(system, cfg, integrator.getRandomNumberSeed());
    if (MIDDLE) allocateScratch(true, true);
    outputFileName = integrator.getOutputFileName();
    saveStatisticsFileName = integrator.getSaveStatisticsFileName();
    statisticsFile = saveStatisticsFileName;
    for (int i = 0; i < integrator.getNumMilestoneGroups(); i++) milestoneGroups.push_back(integrator.getMilestoneGroup(i));
    check(seekr2_b200_write_header(SEEKR2_B200_KIND_MMVT, outputFileName.c_str()));          /* :158-171 */
}

template <class K, class I, bool MIDDLE>
void B200MmvtStepKernel<K, I, MIDDLE>::execute(ContextImpl& context, const I& integrator) {
    cu.setAsCurrent();
    void* stream = (void*) cu.getCurrentStream();
    if (cu.getTime() != expectedTime)                              /* someone called context.setTime(): follow it */
        check(seekr2_b200_set_time(handle, 0, cu.getTime(), cu.getStepCount(), stream));
    updateParams(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize());
    if (cu.getStepCount() <= 0) {                                  /* :205-210 */
        followAtomOrder();
        seekr2_b200_buffers b = buffers();
        check(seekr2_b200_check_first_step(handle, &b, nullptr, stream));
    }
    if (MIDDLE) stepMiddle(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    else stepTwoPass(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    drainTo(outputFileName, SEEKR2_B200_KIND_MMVT, milestoneGroups, 0);
    /* (asynchronous drain: `records` holds what an earlier step's drain delivered; the statistics only change on
       crossings, so the file written now equals the one the reference wrote at the latest collected crossing) */
    if (!records.empty()) writeStatistics();                       /* :313-345 */
    cu.setTime(cu.getTime() + integrator.getStepSize());           /* :362-365 */
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
    expectedTime = cu.getTime();
}

template <class K, class I, bool MIDDLE>
double B200MmvtStepKernel<K, I, MIDDLE>::computeKineticEnergy(ContextImpl&, const I& integrator) {
    flushDrain();                                                  /* a natural synchronisation point: bring the files up to date */
    return cu.getIntegrationUtilities().computeKineticEnergy(0.5 * integrator.getStepSize());   /* :368-370, :917-919 */
}

/* ---- Elber (CudaSeekr2Kernels.cpp:392-599; Middle :947-1177) ---------------------------------------- */

template <class K, class I, bool MIDDLE>
void B200ElberStepKernel<K, I, MIDDLE>::initialize(const System& system, const I& integrator) {
    seekr2_b200_config cfg{};
    cfg.kind = SEEKR2_B200_KIND_ELBER;
    cfg.scheme = MIDDLE ? SEEKR2_B200_SCHEME_MIDDLE : SEEKR2_B200_SCHEME_LANGEVIN;
    std::vector<int32_t> src, dest;
    for (int i = 0; i < integrator.getNumSrcMilestoneGroups(); i++) src.push_back(integrator.getSrcMilestoneGroup(i));
    for (int i = 0; i < integrator.getNumDestMilestoneGroups(); i++) dest.push_back(integrator.getDestMilestoneGroup(i));
    /* :415-436 -- the reference's own check and message, on every force of the System (not only the typed ones) */
    for (const std::vector<int32_t>* list : {&src, &dest})
        for (int32_t group : *list) {
            bool found = false;
            for (int j = 0; j < system.getNumForces(); j++) found = found || system.getForce(j).getForceGroup() == group;
            if (!found) throw OpenMMException("System contains no force groups used to detect Elber boundary crossings. Check for mismatches between force group assignments and the groups added to the Elber integrator.");
        }
    {
        std::vector<int> used(src.begin(), src.end());
        used.insert(used.end(), dest.begin(), dest.end());
        requireTypedMilestoneForces(system, used);
    }
    cfg.num_src = (int) src.size(); cfg.src_groups = src.data();
    cfg.num_dest = (int) dest.size(); cfg.dest_groups = dest.data();
    cfg.end_on_src_milestone = integrator.getEndOnSrcMilestone();  /* :414 */
    const int32_t counter0 = integrator.getCrossingCounter();      /* :403 */
    cfg.crossing_counter0 = &counter0;
    saveStateFileName = integrator.getSaveStateFileName();          /* :405-409 */
    stateNameHasCounter = !MIDDLE;                                  /* :1157 vs :578-580 */
    create(system, cfg, integrator.getRandomNumberSeed());
    if (MIDDLE) allocateScratch(false, true);
    outputFileName = integrator.getOutputFileName();
    labels = src; labels.insert(labels.end(), dest.begin(), dest.end());
    numSrc = (int) src.size();
    check(seekr2_b200_write_header(SEEKR2_B200_KIND_ELBER, outputFileName.c_str()));         /* :447-462 */
}

template <class K, class I, bool MIDDLE>
void B200ElberStepKernel<K, I, MIDDLE>::execute(ContextImpl& context, const I& integrator) {
    cu.setAsCurrent();
    void* stream = (void*) cu.getCurrentStream();
    /* Elber's crossing test looks at the context time (:533,:558) and may reset it (:545): the engine keeps that clock on
       the device.  OpenMM's copy follows it -- predicted as time + stepSize, corrected from the clock every drain carries
       -- and is pushed to the device only when somebody else changed it (context.setTime, a new Context). */
    if (first || cu.getTime() != expectedTime) {
        check(seekr2_b200_set_time(handle, 0, cu.getTime(), cu.getStepCount(), stream));
        pendingClockValid = false;                                  /* a drain still in flight carries the old clock */
        first = false;
    }
    if (MIDDLE) stepMiddle(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    else stepTwoPass(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    drainTo(outputFileName, SEEKR2_B200_KIND_ELBER, labels, numSrc);
    double t = cu.getTime() + integrator.getStepSize();             /* :592, unless the device reset the clock */
    deviceTime(cu.getStepCount() + 1, integrator.getStepSize(), t);
    cu.setTime(t);
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
    expectedTime = t;
}

template <class K, class I, bool MIDDLE>
double B200ElberStepKernel<K, I, MIDDLE>::computeKineticEnergy(ContextImpl&, const I& integrator) {
    /* a natural synchronisation point: bring the file and Context time up to date */
    flushDrain();
    double t = 0;
    if (deviceTime(cu.getStepCount(), integrator.getStepSize(), t)) { cu.setTime(t); expectedTime = t; }
    return cu.getIntegrationUtilities().computeKineticEnergy(0.5 * integrator.getStepSize());   /* :597-599, :1175-1177 */
}

template class B200MmvtStepKernel<IntegrateMmvtLangevinStepKernel, MmvtLangevinIntegrator, false>;
template class B200MmvtStepKernel<IntegrateMmvtLangevinMiddleStepKernel, MmvtLangevinMiddleIntegrator, true>;
template class B200ElberStepKernel<IntegrateElberLangevinStepKernel, ElberLangevinIntegrator, false>;
template class B200ElberStepKernel<IntegrateElberLangevinMiddleStepKernel, ElberLangevinMiddleIntegrator, true>;

}  // namespace Seekr2Plugin
