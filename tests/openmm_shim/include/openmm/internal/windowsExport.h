/* openmm_shim (NOT OpenMM) */
#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif
