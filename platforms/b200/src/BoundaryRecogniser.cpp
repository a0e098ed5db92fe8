/* Walks the System's CustomCentroidBondForce objects and emits the typed boundary descriptors of
 * include/seekr2_b200.h.  Same recognition rules as the Python harness
 * (seekr2_openmm_plugin_b200/system.py::_recognise), which is tested on its own; this file is
 * compiled and run against tests/openmm_shim only (OpenMM is absent here). */
#include "B200Seekr2Kernels.h"
#include "openmm/CustomCentroidBondForce.h"
#include "openmm/OpenMMException.h"
#include <algorithm>
#include <cstdlib>
#include <map>
#include <regex>

using namespace OpenMM;
namespace Seekr2Plugin {

namespace {
struct Typed { int type, style, axis; double kSign; std::string thr; std::string cx, cy, cz; };

Typed recognise(std::string e, int numGroups) {
    e.erase(std::remove_if(e.begin(), e.end(), ::isspace), e.end());
    Typed t{0, SEEKR2_B200_STYLE_RAW, 0, 1.0, "radius", "", "", ""};
    std::smatch m;
    if (std::regex_match(e, m, std::regex(R"(bitcode\*step\((.*)\))"))) { t.style = SEEKR2_B200_STYLE_STEP; e = m[1]; }
    if (!e.empty() && e[0] == '-') { t.kSign = -1.0; e = e.substr(1); }
    if (!std::regex_match(e, m, std::regex(R"(k\*\((.*)\))"))) throw OpenMMException("B200 SEEKR2 plugin: cannot type boundary expression");
    std::string u = m[1];
    if (numGroups == 2 && std::regex_match(u, m, std::regex(R"(distance\(g1,g2\)\^2-(\w+)\^2)"))) { t.type = SEEKR2_B200_CV_SPHERE_PAIR; t.thr = m[1]; return t; }
    if (numGroups == 1 && std::regex_match(u, m, std::regex(R"(\(x1-(\w+)\)\^2\+\(y1-(\w+)\)\^2\+\(z1-(\w+)\)\^2-(\w+)\^2)"))) {
        t.type = SEEKR2_B200_CV_SPHERE_FIXED; t.cx = m[1]; t.cy = m[2]; t.cz = m[3]; t.thr = m[4]; return t; }
    if (numGroups == 1 && std::regex_match(u, m, std::regex(R"(([xyz])1-(\w+))"))) {
        t.type = SEEKR2_B200_CV_PLANE; t.axis = std::string("xyz").find(m[1].str()[0]); t.thr = m[2]; return t; }
    if (numGroups == 4 && std::regex_match(u, m, std::regex(R"(distance\(g1,g2\)\^2\+distance\(g3,g4\)\^2-(\w+)\^2)"))) {
        t.type = SEEKR2_B200_CV_PAIR_SUM; t.axis = 2; t.thr = m[1]; return t; }
    if (numGroups == 6 && std::regex_match(u, m, std::regex(R"(distance\(g1,g2\)\^2\+distance\(g3,g4\)\^2\+distance\(g5,g6\)\^2-(\w+)\^2)"))) {
        t.type = SEEKR2_B200_CV_PAIR_SUM; t.axis = 3; t.thr = m[1]; return t; }     /* demo4:27 */
    if (numGroups == 3 && std::regex_match(u, m, std::regex(R"(angle\(g1,g2,g3\)-(\w+))"))) {
        t.type = SEEKR2_B200_CV_ANGLE; t.thr = m[1]; return t; }                    /* demo5:24 */
    throw OpenMMException("B200 SEEKR2 plugin: cannot type boundary expression");
}
}  // namespace

/* The reference evaluates WHATEVER sits in a milestone's force group (calcForcesAndEnergy(false, true, mask),
   CudaSeekr2Kernels.cpp:248,527,552); this implementation evaluates typed CustomCentroidBondForce terms only, so any
   other Force in such a group must be refused rather than ignored. */
void requireTypedMilestoneForces(const System& system, const std::vector<int>& forceGroups) {
    for (int f = 0; f < system.getNumForces(); f++) {
        const Force& force = system.getForce(f);
        if (std::find(forceGroups.begin(), forceGroups.end(), force.getForceGroup()) == forceGroups.end()) continue;
        if (dynamic_cast<const CustomCentroidBondForce*>(&force) == nullptr)
            throw OpenMMException("B200 SEEKR2 plugin: force group " + std::to_string(force.getForceGroup()) + " is used to detect boundary crossings but holds a Force that is not a CustomCentroidBondForce");
    }
}

B200BoundarySet recogniseBoundaries(const System& system) {
    B200BoundarySet out;
    /* identical groups (same particles, same weights) are lowered once: the reference's examples give every milestone
       Force its own copies of the same two groups (demo3:37-45), and the engine takes at most SEEKR2_B200_MAX_GROUPS */
    std::map<std::pair<std::vector<int>, std::vector<double>>, int> seen;
    for (int f = 0; f < system.getNumForces(); f++) {
        const auto* force = dynamic_cast<const CustomCentroidBondForce*>(&system.getForce(f));
        if (force == nullptr) continue;
        if (force->usesPeriodicBoundaryConditions()) throw OpenMMException("B200 SEEKR2 plugin: periodic centroid boundaries are not supported");
        const Typed t = recognise(force->getEnergyFunction(), force->getNumGroupsPerBond());
        std::vector<int> remap;
        for (int g = 0; g < force->getNumGroups(); g++) {
            std::vector<int> particles; std::vector<double> weights;
            force->getGroupParameters(g, particles, weights);
            /* without explicit weights a centroid group is mass weighted */
            std::vector<double> w(particles.size());
            for (size_t i = 0; i < particles.size(); i++) w[i] = weights.empty() ? system.getParticleMass(particles[i]) : weights[i];
            const auto key = std::make_pair(particles, w);
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(key, (int) out.groupOffsets.size() - 1).first;
                for (size_t i = 0; i < particles.size(); i++) { out.groupAtoms.push_back(particles[i]); out.groupWeights.push_back(w[i]); }
                out.groupOffsets.push_back((int32_t) out.groupAtoms.size());
            }
            remap.push_back(it->second);
        }
        std::map<std::string, int> index;
        for (int p = 0; p < force->getNumPerBondParameters(); p++) index[force->getPerBondParameterName(p)] = p;
        /* every name in the expression must be a per-bond parameter (or a numeric literal): a global parameter, or a
           name the Force does not define, would silently give a different surface than the reference evaluates */
        auto value = [&](const std::vector<double>& params, const std::string& name) {
            auto it = index.find(name);
            if (it != index.end()) return params[it->second];
            char* end = nullptr;
            const double literal = std::strtod(name.c_str(), &end);
            if (end != name.c_str() && *end == '\0') return literal;
            throw OpenMMException("B200 SEEKR2 plugin: '" + name + "' in boundary expression '" + force->getEnergyFunction() + "' is not a per-bond parameter of the Force");
        };
        for (int g = 0; g < force->getNumGroups(); g++) {
            std::vector<int> particles; std::vector<double> weights;
            force->getGroupParameters(g, particles, weights);
            for (int particle : particles)
                if (system.isVirtualSite(particle)) throw OpenMMException("B200 SEEKR2 plugin: boundary groups containing virtual sites are not supported");
        }
        for (int b = 0; b < force->getNumBonds(); b++) {
            std::vector<int> groups; std::vector<double> params;
            force->getBondParameters(b, groups, params);
            seekr2_b200_boundary bd{};
            bd.type = t.type; bd.style = t.style; bd.axis = t.axis; bd.force_group = force->getForceGroup();
            for (size_t g = 0; g < groups.size() && g < 6; g++) bd.groups[g] = remap[groups[g]];
            bd.bitcode = t.style == SEEKR2_B200_STYLE_STEP ? value(params, "bitcode") : 1.0;
            bd.k = t.kSign * value(params, "k");
            bd.threshold = value(params, t.thr);
            if (t.type == SEEKR2_B200_CV_SPHERE_FIXED) { bd.center[0] = value(params, t.cx); bd.center[1] = value(params, t.cy); bd.center[2] = value(params, t.cz); }
            out.boundaries.push_back(bd);
        }
    }
    return out;
}
}  // namespace Seekr2Plugin
