/* Plugin entry points: same symbols and behaviour as the reference's CUDA plugin
 * (seekr2plugin/platforms/cuda/src/CudaSeekr2KernelFactory.cpp:46-84).  OpenMM is absent here:
 * compiled and run against tests/openmm_shim only. */
#include "B200Seekr2Kernels.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaPlatform.h"

using namespace OpenMM;
namespace Seekr2Plugin {
class B200Seekr2KernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const override {
        CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];   /* :74 */
        if (name == IntegrateMmvtLangevinStepKernel::Name()) return new B200IntegrateMmvtLangevinStepKernel(name, platform, cu);
        if (name == IntegrateElberLangevinStepKernel::Name()) return new B200IntegrateElberLangevinStepKernel(name, platform, cu);
        if (name == IntegrateMmvtLangevinMiddleStepKernel::Name()) return new B200IntegrateMmvtLangevinMiddleStepKernel(name, platform, cu);
        if (name == IntegrateElberLangevinMiddleStepKernel::Name()) return new B200IntegrateElberLangevinMiddleStepKernel(name, platform, cu);
        throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());  /* :83 */
    }
};
}  // namespace Seekr2Plugin

extern "C" void registerPlatforms() {}

extern "C" void registerKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        auto* factory = new Seekr2Plugin::B200Seekr2KernelFactory();     /* never freed, like the reference (:52) */
        platform.registerKernelFactory(Seekr2Plugin::IntegrateMmvtLangevinStepKernel::Name(), factory);
        platform.registerKernelFactory(Seekr2Plugin::IntegrateElberLangevinStepKernel::Name(), factory);
        platform.registerKernelFactory(Seekr2Plugin::IntegrateMmvtLangevinMiddleStepKernel::Name(), factory);
        platform.registerKernelFactory(Seekr2Plugin::IntegrateElberLangevinMiddleStepKernel::Name(), factory);
    } catch (std::exception&) {
        /* the CUDA platform is not available: ignore, like the reference (:58-60) */
    }
}

/* for statically linked tests (:63-71).  The library replaces the reference's CUDA plugin, so it answers to the
   reference's symbol as well: the reference's own platform tests link against it unchanged. */
extern "C" void registerSeekr2B200KernelFactories() {
    try { Platform::getPlatformByName("CUDA"); } catch (...) { Platform::registerPlatform(new CudaPlatform()); }
    registerKernelFactories();
}
extern "C" void registerSeekr2CudaKernelFactories() { registerSeekr2B200KernelFactories(); }
