#ifndef B200_SEEKR2_KERNELS_H_
#define B200_SEEKR2_KERNELS_H_
/* OpenMM KernelImpls over the B200 C ABI.  Drop-in counterparts of
 * CudaIntegrateMmvtLangevinStepKernel / CudaIntegrateElberLangevinStepKernel
 * (seekr2plugin/platforms/cuda/src/CudaSeekr2Kernels.h:46-160).  OpenMM is absent from this environment: the
 * adapter is compiled and exercised against tests/openmm_shim only -- see platforms/b200/README.md. */
#include "Seekr2Kernels.h"
#include "openmm/cuda/CudaContext.h"
#include "seekr2_b200.h"
#include <string>
#include <vector>

namespace Seekr2Plugin {

/* typed boundary description extracted from the System (BoundaryRecogniser.cpp) */
struct B200BoundarySet {
    std::vector<int32_t> groupOffsets{0};
    std::vector<int32_t> groupAtoms;
    std::vector<double> groupWeights;
    std::vector<seekr2_b200_boundary> boundaries;
};
/* throws OpenMM::OpenMMException for any CustomCentroidBondForce it cannot type */
B200BoundarySet recogniseBoundaries(const OpenMM::System& system);
/* throws if one of these force groups holds a Force that is not a CustomCentroidBondForce */
void requireTypedMilestoneForces(const OpenMM::System& system, const std::vector<int>& forceGroups);

class B200StepKernelBase {
protected:
    explicit B200StepKernelBase(OpenMM::CudaContext& cu) : cu(cu) {}
    ~B200StepKernelBase();
    void create(const OpenMM::System& system, seekr2_b200_config& cfg, int seed);
    std::string saveStateFileName;               /* empty = crossing states are not kept */
    bool stateNameHasCounter = true;
    seekr2_b200_buffers buffers();
    void check(int status) const;                 /* non-zero status -> OpenMMException(last_error) */
    void updateParams(double temperature, double friction, double stepSize);
    void stepTwoPass(double temperature, double friction, double stepSize, double tolerance);
    void stepMiddle(double temperature, double friction, double stepSize, double tolerance);
    void allocateScratch(bool oldVelm, bool oldDelta);
    OpenMM::CudaArray oldVelm, oldDelta;          /* multi-pass scratch (CudaSeekr2Kernels.cpp:618-643) */
    void drainTo(const std::string& file, int kind, const std::vector<int32_t>& labels, int numSrc);
    /* asynchronous drain (default; SEEKR2_B200_ASYNC_DRAIN=0 or a saveStateFileName turn it off): execute() never waits
       for the device -- a drain is begun behind a step only when the device has reported something to fetch
       (seekr2_b200_poll) and collected by a later execute() once it has landed; flushDrain() (blocking) runs where the
       reference's files must be complete: computeKineticEnergy, destruction */
    bool asyncDrain = false, drainPending = false;
    bool collected = false;          /* this execute() collected a drain (its clock is in the engine's staging area) */
    bool clockValid = false, pendingClockValid = true;   /* ... and no set_time was issued after that drain was queued */
    int64_t seenResets = 0;
    int kind = 0;
    bool fused = false;              /* one-launch step (no constraints, no virtual sites) */
    std::vector<int32_t> slotIndex;  /* current buffer index of every boundary-group atom (cu.getAtomIndex()) */
    void followAtomOrder();
    bool deviceTime(long long stepsIssued, double stepSize, double& time);
    void noteClock(bool valid);
    double clockTime = 0.0;          /* the anchor clock that came with the latest collected drain ... */
    long long clockStep = 0;         /* ... and the step count it belongs to */
    int64_t collectPending(bool block);           /* records handed to writeRecords(); -1 = not ready */
    void flushDrain();
    void writeRecords(int64_t n);
    void writeStatistics();
    std::string drainFile, statisticsFile;
    int drainKind = 0, drainNumSrc = 0;
    std::vector<int32_t> drainLabels;
    OpenMM::CudaContext& cu;
    seekr2_b200_handle handle = nullptr;
    B200BoundarySet boundarySet;
    double prevTemp = -1, prevFriction = -1, prevStepSize = -1;
    std::vector<seekr2_b200_record> records;
};

/* One implementation per method, instantiated for the plain and the LangevinMiddle integrator
 * (Seekr2Kernels.h:53-172): MIDDLE selects scheme = SEEKR2_B200_SCHEME_MIDDLE and the three-pass call sequence
 * kick / middle_drift / finish of CudaSeekr2Kernels.cpp:762-802 instead of kick / finish. */
template <class KERNEL, class INTEGRATOR, bool MIDDLE>
class B200MmvtStepKernel : public KERNEL, private B200StepKernelBase {
public:
    B200MmvtStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu) : KERNEL(name, platform), B200StepKernelBase(cu) {}
    void initialize(const OpenMM::System& system, const INTEGRATOR& integrator) override;
    void execute(OpenMM::ContextImpl& context, const INTEGRATOR& integrator) override;
    double computeKineticEnergy(OpenMM::ContextImpl& context, const INTEGRATOR& integrator) override;
private:
    std::string outputFileName, saveStatisticsFileName;
    std::vector<int32_t> milestoneGroups;
    double expectedTime = 0.0;       /* the device clock, as long as nobody called context.setTime() */
};

template <class KERNEL, class INTEGRATOR, bool MIDDLE>
class B200ElberStepKernel : public KERNEL, private B200StepKernelBase {
public:
    B200ElberStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu) : KERNEL(name, platform), B200StepKernelBase(cu) {}
    void initialize(const OpenMM::System& system, const INTEGRATOR& integrator) override;
    void execute(OpenMM::ContextImpl& context, const INTEGRATOR& integrator) override;
    double computeKineticEnergy(OpenMM::ContextImpl& context, const INTEGRATOR& integrator) override;
private:
    std::string outputFileName;
    std::vector<int32_t> labels;     /* source groups then destination groups */
    int numSrc = 0;
    bool first = true;
    double expectedTime = 0.0;       /* what cu.getTime() holds unless somebody called context.setTime() */
};

typedef B200MmvtStepKernel<IntegrateMmvtLangevinStepKernel, MmvtLangevinIntegrator, false> B200IntegrateMmvtLangevinStepKernel;
typedef B200MmvtStepKernel<IntegrateMmvtLangevinMiddleStepKernel, MmvtLangevinMiddleIntegrator, true> B200IntegrateMmvtLangevinMiddleStepKernel;
typedef B200ElberStepKernel<IntegrateElberLangevinStepKernel, ElberLangevinIntegrator, false> B200IntegrateElberLangevinStepKernel;
typedef B200ElberStepKernel<IntegrateElberLangevinMiddleStepKernel, ElberLangevinMiddleIntegrator, true> B200IntegrateElberLangevinMiddleStepKernel;

}  // namespace Seekr2Plugin
#endif
