/* openmm_shim (NOT OpenMM, NOT SFMT): the two calls the reference's tests make, backed by std::mt19937_64. */
#ifndef OPENMM_SHIM_SFMT_H_
#define OPENMM_SHIM_SFMT_H_
#include <cstdint>
#include <random>
namespace OpenMM_SFMT {
class SFMT { public: std::mt19937_64 engine; };
inline void init_gen_rand(uint32_t seed, SFMT& sfmt) { sfmt.engine.seed(seed); }
inline double genrand_real2(SFMT& sfmt) { return (sfmt.engine() >> 11) * (1.0 / 9007199254740992.0); }   /* [0,1) */
}
#endif
