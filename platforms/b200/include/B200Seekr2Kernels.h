#ifndef B200_SEEKR2_KERNELS_H_
#define B200_SEEKR2_KERNELS_H_
/* OpenMM KernelImpls over the B200 C ABI.  Drop-in counterparts of
 * CudaIntegrateMmvtLangevinStepKernel / CudaIntegrateElberLangevinStepKernel
 * (seekr2plugin/platforms/cuda/src/CudaSeekr2Kernels.h:46-160).  NOT COMPILED in this environment
 * (OpenMM absent) -- see platforms/b200/README.md. */
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

class B200StepKernelBase {
protected:
    explicit B200StepKernelBase(OpenMM::CudaContext& cu) : cu(cu) {}
    ~B200StepKernelBase();
    void create(const OpenMM::System& system, seekr2_b200_config& cfg, int seed);
    std::string mmvtOutput, mmvtStats;
    std::string saveStateFileName;               /* empty = crossing states are not kept */
    std::vector<int32_t> mmvtGroups;
    seekr2_b200_buffers buffers();
    void check(int status) const;                 /* non-zero status -> OpenMMException(last_error) */
    void stepTwoPass(double temperature, double friction, double stepSize, double tolerance);
    void stepMiddle(double temperature, double friction, double stepSize, double tolerance);
    void allocateScratch(bool oldVelm, bool oldDelta);
    OpenMM::CudaArray oldVelm, oldDelta;          /* multi-pass scratch (CudaSeekr2Kernels.cpp:618-643) */
    void drainTo(const std::string& file, int kind, const std::vector<int32_t>& labels, int numSrc);
    OpenMM::CudaContext& cu;
    seekr2_b200_handle handle = nullptr;
    B200BoundarySet boundarySet;
    double prevTemp = -1, prevFriction = -1, prevStepSize = -1;
    std::vector<seekr2_b200_record> records;
};

class B200IntegrateMmvtLangevinStepKernel : public IntegrateMmvtLangevinStepKernel, private B200StepKernelBase {
public:
    B200IntegrateMmvtLangevinStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu)
        : IntegrateMmvtLangevinStepKernel(name, platform), B200StepKernelBase(cu) {}
    void initialize(const OpenMM::System& system, const MmvtLangevinIntegrator& integrator) override;
    void execute(OpenMM::ContextImpl& context, const MmvtLangevinIntegrator& integrator) override;
    double computeKineticEnergy(OpenMM::ContextImpl& context, const MmvtLangevinIntegrator& integrator) override;
private:
    std::string outputFileName, saveStatisticsFileName;
    std::vector<int32_t> milestoneGroups;
};

class B200IntegrateElberLangevinStepKernel : public IntegrateElberLangevinStepKernel, private B200StepKernelBase {
public:
    B200IntegrateElberLangevinStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu)
        : IntegrateElberLangevinStepKernel(name, platform), B200StepKernelBase(cu) {}
    void initialize(const OpenMM::System& system, const ElberLangevinIntegrator& integrator) override;
    void execute(OpenMM::ContextImpl& context, const ElberLangevinIntegrator& integrator) override;
    double computeKineticEnergy(OpenMM::ContextImpl& context, const ElberLangevinIntegrator& integrator) override;
private:
    std::string outputFileName;
    std::vector<int32_t> labels;     /* source groups then destination groups */
    int numSrc = 0;
};

/* LangevinMiddle twins (Seekr2Kernels.h:115-172; CudaSeekr2Kernels.cpp:604-1177): same engine with
 * scheme = SEEKR2_B200_SCHEME_MIDDLE and the three-pass call sequence kick / middle_drift / finish. */
class B200IntegrateMmvtLangevinMiddleStepKernel : public IntegrateMmvtLangevinMiddleStepKernel, private B200StepKernelBase {
public:
    B200IntegrateMmvtLangevinMiddleStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu)
        : IntegrateMmvtLangevinMiddleStepKernel(name, platform), B200StepKernelBase(cu) {}
    void initialize(const OpenMM::System& system, const MmvtLangevinMiddleIntegrator& integrator) override;
    void execute(OpenMM::ContextImpl& context, const MmvtLangevinMiddleIntegrator& integrator) override;
    double computeKineticEnergy(OpenMM::ContextImpl& context, const MmvtLangevinMiddleIntegrator& integrator) override;
};

}  // namespace Seekr2Plugin
#endif
