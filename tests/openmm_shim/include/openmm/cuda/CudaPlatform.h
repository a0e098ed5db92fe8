/* openmm_shim (NOT OpenMM): see ShimCuda.h */
#include "openmm/cuda/ShimCuda.h"
