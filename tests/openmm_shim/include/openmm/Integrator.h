/* openmm_shim (NOT OpenMM): see ShimCore.h */
#include "openmm/ShimCore.h"
