/* openmm_shim (NOT OpenMM): the assertion macros the reference's sources and tests use. */
#ifndef OPENMM_SHIM_ASSERTIONUTILITIES_H_
#define OPENMM_SHIM_ASSERTIONUTILITIES_H_
#include "openmm/ShimCore.h"
#include <cmath>
#include <sstream>
#include <string>

namespace OpenMM {
void throwException(const char* file, int line, const std::string& details);
}

#define ASSERT(cond) {if (!(cond)) OpenMM::throwException(__FILE__, __LINE__, "");};
#define ASSERT_VALID_INDEX(index, vector) {if (index < 0 || index >= (int) vector.size()) OpenMM::throwException(__FILE__, __LINE__, "Index out of range");};
#define ASSERT_EQUAL(expected, found) {if (!((expected) == (found))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_TOL(expected, found, tol) {double _scale_ = std::abs(expected) > 1.0 ? std::abs(expected) : 1.0; if (!(std::abs((expected)-(found))/_scale_ <= (tol))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_USUALLY_EQUAL_TOL(expected, found, tol) {double _scale_ = std::abs(expected) > 1.0 ? std::abs(expected) : 1.0; if (!(std::abs((expected)-(found))/_scale_ <= (tol))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found)<<" (This test is stochastic and may occasionally fail)"; OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_VEC(expected, found, tol) {double _norm_ = std::sqrt((expected).dot(expected)); double _scale_ = _norm_ > 1.0 ? _norm_ : 1.0; if ((std::abs(((expected)[0])-((found)[0]))/_scale_ > (tol)) || (std::abs(((expected)[1])-((found)[1]))/_scale_ > (tol)) || (std::abs(((expected)[2])-((found)[2]))/_scale_ > (tol))) {std::stringstream details; details << " Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#endif
