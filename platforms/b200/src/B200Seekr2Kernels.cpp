/* OpenMM KernelImpls over the B200 C ABI: the replacement of
 * seekr2plugin/platforms/cuda/src/CudaSeekr2Kernels.cpp:63-599.  NOT COMPILED here (no OpenMM).
 *
 * Per step the reference does Part1 -> applyConstraints -> Part2 -> computeVirtualSites ->
 * calcForcesAndEnergy(false,true,mask) [blocking readback] -> host crossing block -> mmvtBounce.
 * Here: seekr2_b200_kick -> applyConstraints -> seekr2_b200_finish (decision + bounce on the device);
 * the crossing records are drained when the ring is polled (every step by default, to keep OpenMM's
 * Context time in step for Elber's setTime(0); `drainInterval` steps otherwise).  The first-step test
 * (:205-210) and the Elber force-group check (:415-436) keep their exceptions and messages. */
#include "B200Seekr2Kernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaIntegrationUtilities.h"
#include <fstream>

using namespace OpenMM;
namespace Seekr2Plugin {

B200StepKernelBase::~B200StepKernelBase() {
    cu.setAsCurrent();                                            /* :74-80 */
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
    cfg.abi_version = SEEKR2_B200_ABI_VERSION;
    cfg.precision = cu.getUseDoublePrecision() ? SEEKR2_B200_DOUBLE : cu.getUseMixedPrecision() ? SEEKR2_B200_MIXED : SEEKR2_B200_SINGLE;
    cfg.variant = SEEKR2_B200_VARIANT_CUDA;
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

void B200StepKernelBase::stepTwoPass(double temperature, double friction, double stepSize, double tolerance) {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    integration.setNextStepSize(stepSize);                         /* :187 */
    if (temperature != prevTemp || friction != prevFriction || stepSize != prevStepSize) {   /* :188-203 */
        check(seekr2_b200_set_params(handle, temperature, friction, stepSize));
        prevTemp = temperature; prevFriction = friction; prevStepSize = stepSize;
    }
    seekr2_b200_buffers b = buffers();
    void* stream = (void*) cu.getCurrentStream();
    const int randomIndex = integration.prepareRandomNumbers(cu.getPaddedNumAtoms());       /* :214 */
    check(seekr2_b200_kick(handle, &b, (uint32_t) randomIndex, stream));                    /* Part1 :215-223 */
    integration.applyConstraints(tolerance);                                                /* :227 */
    check(seekr2_b200_finish(handle, &b, stream));                                          /* Part2 + CV + crossing + bounce :231-358 */
    integration.computeVirtualSites();                                                      /* :239 (groups with virtual sites are refused) */
}

void B200StepKernelBase::allocateScratch(bool needOldVelm, bool needOldDelta) {
    const size_t elem = (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) ? 4 * sizeof(double) : 4 * sizeof(float);
    if (needOldVelm && !oldVelm.isInitialized()) oldVelm.initialize(cu, cu.getPaddedNumAtoms(), elem, "b200OldVelm");
    if (needOldDelta && !oldDelta.isInitialized()) oldDelta.initialize(cu, cu.getPaddedNumAtoms(), elem, "b200OldDelta");
}

/* CudaIntegrateMmvtLangevinMiddleStepKernel::execute, integration part (CudaSeekr2Kernels.cpp:756-802) */
void B200StepKernelBase::stepMiddle(double temperature, double friction, double stepSize, double tolerance) {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    integration.setNextStepSize(stepSize);
    if (temperature != prevTemp || friction != prevFriction || stepSize != prevStepSize) {   /* :734-748 */
        check(seekr2_b200_set_params(handle, temperature, friction, stepSize));
        prevTemp = temperature; prevFriction = friction; prevStepSize = stepSize;
    }
    seekr2_b200_buffers b = buffers();
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
    records.resize(4096);
    int64_t n = 0;
    check(seekr2_b200_drain(handle, records.data(), (int64_t) records.size(), &n, (void*) cu.getCurrentStream()));
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
        const std::string name = saveStateFileName + "_" + std::to_string(r.counter) + "_" + std::to_string(label);
        check(seekr2_b200_write_state_xml(name.c_str(), cu.getNumAtoms(), xo.data(), vo.data(), r.time, box));
    }
}

/* ---- MMVT ------------------------------------------------------------------------------------------- */

void B200IntegrateMmvtLangevinStepKernel::initialize(const System& system, const MmvtLangevinIntegrator& integrator) {
    seekr2_b200_config cfg{};
    cfg.kind = SEEKR2_B200_KIND_MMVT;
    cfg.num_milestone_groups = integrator.getNumMilestoneGroups();
    const int32_t counter0 = integrator.getBounceCounter();        /* :151 */
    cfg.bounce_counter0 = &counter0;
    saveStateFileName = integrator.getSaveStateFileName();          /* :135-139 */
    create(system, cfg, integrator.getRandomNumberSeed());
    outputFileName = integrator.getOutputFileName();
    saveStatisticsFileName = integrator.getSaveStatisticsFileName();
    for (int i = 0; i < integrator.getNumMilestoneGroups(); i++) milestoneGroups.push_back(integrator.getMilestoneGroup(i));
    check(seekr2_b200_write_header(SEEKR2_B200_KIND_MMVT, outputFileName.c_str()));          /* :158-171 */
}

void B200IntegrateMmvtLangevinStepKernel::execute(ContextImpl& context, const MmvtLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    if (cu.getStepCount() <= 0) {                                  /* :205-210 */
        seekr2_b200_buffers b = buffers();
        check(seekr2_b200_check_first_step(handle, &b, nullptr, (void*) cu.getCurrentStream()));
    }
    stepTwoPass(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    drainTo(outputFileName, SEEKR2_B200_KIND_MMVT, milestoneGroups, 0);
    if (!saveStatisticsFileName.empty() && !records.empty()) {     /* :313-345 */
        seekr2_b200_anchor_state st;
        check(seekr2_b200_get_state(handle, 0, &st, (void*) cu.getCurrentStream()));
        check(seekr2_b200_write_statistics(saveStatisticsFileName.c_str(), &st, milestoneGroups.data(), (int32_t) milestoneGroups.size()));
    }
    cu.setTime(cu.getTime() + integrator.getStepSize());           /* :362-365 */
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
}

double B200IntegrateMmvtLangevinStepKernel::computeKineticEnergy(ContextImpl&, const MmvtLangevinIntegrator& integrator) {
    return cu.getIntegrationUtilities().computeKineticEnergy(0.5 * integrator.getStepSize());   /* :368-370 */
}

/* ---- Elber ------------------------------------------------------------------------------------------ */

void B200IntegrateElberLangevinStepKernel::initialize(const System& system, const ElberLangevinIntegrator& integrator) {
    seekr2_b200_config cfg{};
    cfg.kind = SEEKR2_B200_KIND_ELBER;
    std::vector<int32_t> src, dest;
    for (int i = 0; i < integrator.getNumSrcMilestoneGroups(); i++) src.push_back(integrator.getSrcMilestoneGroup(i));
    for (int i = 0; i < integrator.getNumDestMilestoneGroups(); i++) dest.push_back(integrator.getDestMilestoneGroup(i));
    cfg.num_src = (int) src.size(); cfg.src_groups = src.data();
    cfg.num_dest = (int) dest.size(); cfg.dest_groups = dest.data();
    cfg.end_on_src_milestone = integrator.getEndOnSrcMilestone();  /* :414 */
    const int32_t counter0 = integrator.getCrossingCounter();      /* :403 */
    cfg.crossing_counter0 = &counter0;
    saveStateFileName = integrator.getSaveStateFileName();          /* :405-409 */
    create(system, cfg, integrator.getRandomNumberSeed());          /* missing force group -> the reference's message (:421-422) */
    outputFileName = integrator.getOutputFileName();
    labels = src; labels.insert(labels.end(), dest.begin(), dest.end());
    numSrc = (int) src.size();
    check(seekr2_b200_write_header(SEEKR2_B200_KIND_ELBER, outputFileName.c_str()));         /* :447-462 */
}

void B200IntegrateElberLangevinStepKernel::execute(ContextImpl& context, const ElberLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    /* Elber's crossing test looks at the context time (:533,:558) and may reset it (:545): the engine
       keeps that clock on the device; mirror OpenMM's copy into it before and out of it after */
    check(seekr2_b200_set_time(handle, 0, cu.getTime(), cu.getStepCount(), (void*) cu.getCurrentStream()));
    stepTwoPass(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    drainTo(outputFileName, SEEKR2_B200_KIND_ELBER, labels, numSrc);
    seekr2_b200_anchor_state st;
    check(seekr2_b200_get_state(handle, 0, &st, (void*) cu.getCurrentStream()));
    cu.setTime(st.time);                                           /* already advanced by stepSize (:592) */
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
}

double B200IntegrateElberLangevinStepKernel::computeKineticEnergy(ContextImpl&, const ElberLangevinIntegrator& integrator) {
    return cu.getIntegrationUtilities().computeKineticEnergy(0.5 * integrator.getStepSize());   /* :597-599 */
}

/* ---- MMVT, LangevinMiddle ---------------------------------------------------------------------------- */

void B200IntegrateMmvtLangevinMiddleStepKernel::initialize(const System& system, const MmvtLangevinMiddleIntegrator& integrator) {
    seekr2_b200_config cfg{};
    cfg.kind = SEEKR2_B200_KIND_MMVT;
    cfg.scheme = SEEKR2_B200_SCHEME_MIDDLE;
    cfg.num_milestone_groups = integrator.getNumMilestoneGroups();
    const int32_t counter0 = integrator.getBounceCounter();
    cfg.bounce_counter0 = &counter0;
    saveStateFileName = integrator.getSaveStateFileName();
    create(system, cfg, integrator.getRandomNumberSeed());
    allocateScratch(true, true);
    mmvtOutput = integrator.getOutputFileName();
    mmvtStats = integrator.getSaveStatisticsFileName();
    for (int i = 0; i < integrator.getNumMilestoneGroups(); i++) mmvtGroups.push_back(integrator.getMilestoneGroup(i));
    check(seekr2_b200_write_header(SEEKR2_B200_KIND_MMVT, mmvtOutput.c_str()));
}

void B200IntegrateMmvtLangevinMiddleStepKernel::execute(ContextImpl& context, const MmvtLangevinMiddleIntegrator& integrator) {
    cu.setAsCurrent();
    if (cu.getStepCount() <= 0) {                                  /* :750-755 */
        seekr2_b200_buffers b = buffers();
        check(seekr2_b200_check_first_step(handle, &b, nullptr, (void*) cu.getCurrentStream()));
    }
    stepMiddle(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), integrator.getConstraintTolerance());
    drainTo(mmvtOutput, SEEKR2_B200_KIND_MMVT, mmvtGroups, 0);
    if (!mmvtStats.empty() && !records.empty()) {
        seekr2_b200_anchor_state st;
        check(seekr2_b200_get_state(handle, 0, &st, (void*) cu.getCurrentStream()));
        check(seekr2_b200_write_statistics(mmvtStats.c_str(), &st, mmvtGroups.data(), (int32_t) mmvtGroups.size()));
    }
    cu.setTime(cu.getTime() + integrator.getStepSize());           /* :911-914 */
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
}

double B200IntegrateMmvtLangevinMiddleStepKernel::computeKineticEnergy(ContextImpl&, const MmvtLangevinMiddleIntegrator& integrator) {
    return cu.getIntegrationUtilities().computeKineticEnergy(0.5 * integrator.getStepSize());
}

/* ElberLangevinMiddle: identical to B200IntegrateElberLangevinStepKernel with cfg.scheme = MIDDLE and
   stepMiddle() in place of stepTwoPass() (CudaSeekr2Kernels.cpp:1032-1173); left to the maintainer who
   can compile against OpenMM -- every C-ABI entry point it needs is exercised by tests/test_gpu_middle.py. */

}  // namespace Seekr2Plugin
