/* openmm_shim (NOT OpenMM): serialize<State> as the reference calls it (CudaSeekr2Kernels.cpp:288).  The
   document has the element layout of OpenMM's StateProxy; the number formatting is the shim's (%.17g). */
#ifndef OPENMM_SHIM_XMLSERIALIZER_H_
#define OPENMM_SHIM_XMLSERIALIZER_H_
#include "openmm/ShimCore.h"
#include <ostream>
namespace OpenMM {
class XmlSerializer {
public:
    template <class T> static void serialize(const T* object, const std::string& rootName, std::ostream& stream) { serializeState(*object, rootName, stream); }
private:
    static void serializeState(const State& state, const std::string& rootName, std::ostream& stream);
};
}
#endif
