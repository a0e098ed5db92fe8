/* tests/openmm_shim -- TEST INFRASTRUCTURE ONLY.  NOT OpenMM (see include/openmm/ShimCore.h).
 * Host side: System, forces (incl. the expression evaluator of CustomCentroidBondForce), Platform registry,
 * Context / ContextImpl, State serializer. */
#include "openmm/ShimCore.h"
#include "openmm/internal/AssertionUtilities.h"
#include "openmm/reference/SimTKOpenMMRealType.h"
#include "openmm/serialization/XmlSerializer.h"
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <ostream>
#include <random>
#include <sstream>

namespace OpenMM {

std::ostream& operator<<(std::ostream& o, const Vec3& v) { return o << "[" << v[0] << ", " << v[1] << ", " << v[2] << "]"; }

void throwException(const char* file, int line, const std::string& details) {
    std::string fn(file);
    std::string::size_type pos = fn.find_last_of("/\\");
    if (pos + 1 >= fn.size()) pos = 0;
    std::stringstream message;
    message << "Assertion failure at " << fn.substr(pos + 1) << ":" << line;
    if (details.size() > 0) message << ".  " << details;
    throw OpenMMException(message.str());
}

/* ---- System -------------------------------------------------------------------------------------------- */

System::System() { box[0] = Vec3(2, 0, 0); box[1] = Vec3(0, 2, 0); box[2] = Vec3(0, 0, 2); }
System::~System() { for (Force* f : forces) delete f; }
int System::addConstraint(int particle1, int particle2, double distance) {
    constraints.push_back(Constraint{particle1, particle2, distance});
    return (int) constraints.size() - 1;
}
void System::getConstraintParameters(int index, int& particle1, int& particle2, double& distance) const {
    const Constraint& c = constraints.at(index);
    particle1 = c.p1; particle2 = c.p2; distance = c.distance;
}

/* ---- simple forces ------------------------------------------------------------------------------------ */

int HarmonicBondForce::addBond(int particle1, int particle2, double length, double k) {
    bonds.push_back(Bond{particle1, particle2, length, k});
    return (int) bonds.size() - 1;
}
void HarmonicBondForce::getBondParameters(int index, int& particle1, int& particle2, double& length, double& k) const {
    const Bond& b = bonds.at(index);
    particle1 = b.p1; particle2 = b.p2; length = b.length; k = b.k;
}
double HarmonicBondForce::shimEvaluate(const std::vector<Vec3>& x, const std::vector<double>&, std::vector<Vec3>* f) const {
    double energy = 0;
    for (const Bond& b : bonds) {
        const Vec3 d = x[b.p2] - x[b.p1];
        const double r = std::sqrt(d.dot(d)), dr = r - b.length;
        energy += 0.5 * b.k * dr * dr;
        if (f && r > 0) {
            const Vec3 g = d * (b.k * dr / r);     /* dE/dx2 */
            (*f)[b.p1] += g;
            (*f)[b.p2] -= g;
        }
    }
    return energy;
}

int NonbondedForce::addParticle(double charge, double sigma, double epsilon) {
    particles.push_back(Particle{charge, sigma, epsilon});
    return (int) particles.size() - 1;
}
void NonbondedForce::getParticleParameters(int index, double& charge, double& sigma, double& epsilon) const {
    const Particle& p = particles.at(index);
    charge = p.charge; sigma = p.sigma; epsilon = p.epsilon;
}
/* NoCutoff: Coulomb + Lennard-Jones with Lorentz-Berthelot combination, every pair */
double NonbondedForce::shimEvaluate(const std::vector<Vec3>& x, const std::vector<double>&, std::vector<Vec3>* f) const {
    double energy = 0;
    const int n = (int) particles.size();
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            const Vec3 d = x[j] - x[i];
            const double r2 = d.dot(d), r = std::sqrt(r2), inv = 1.0 / r;
            const double sig = 0.5 * (particles[i].sigma + particles[j].sigma);
            const double eps = std::sqrt(particles[i].epsilon * particles[j].epsilon);
            const double s2 = sig * sig / r2, s6 = s2 * s2 * s2;
            const double qq = ONE_4PI_EPS0 * particles[i].charge * particles[j].charge;
            energy += qq * inv + 4 * eps * (s6 * s6 - s6);
            if (f) {
                const double dEdr = -qq * inv * inv - 4 * eps * (12 * s6 * s6 - 6 * s6) * inv;
                const Vec3 g = d * (dEdr * inv);   /* dE/dxj */
                (*f)[j] -= g;
                (*f)[i] += g;
            }
        }
    return energy;
}

/* ---- CustomCentroidBondForce -------------------------------------------------------------------------- */

int CustomCentroidBondForce::addGroup(const std::vector<int>& particles, const std::vector<double>& weights) {
    if (!weights.empty() && weights.size() != particles.size())
        throw OpenMMException("CustomCentroidBondForce: wrong number of weights specified for a group.");
    groups.push_back(Group{particles, weights});
    return (int) groups.size() - 1;
}
void CustomCentroidBondForce::getGroupParameters(int index, std::vector<int>& particles, std::vector<double>& weights) const {
    ASSERT_VALID_INDEX(index, groups);
    particles = groups[index].particles; weights = groups[index].weights;
}
int CustomCentroidBondForce::addBond(const std::vector<int>& bondGroups, const std::vector<double>& parameters) {
    if ((int) bondGroups.size() != groupsPerBond) throw OpenMMException("CustomCentroidBondForce: wrong number of groups specified for a bond.");
    bonds.push_back(Bond{bondGroups, parameters});
    return (int) bonds.size() - 1;
}
void CustomCentroidBondForce::getBondParameters(int index, std::vector<int>& bondGroups, std::vector<double>& parameters) const {
    ASSERT_VALID_INDEX(index, bonds);
    bondGroups = bonds[index].groups; parameters = bonds[index].parameters;
}
void CustomCentroidBondForce::setBondParameters(int index, const std::vector<int>& bondGroups, const std::vector<double>& parameters) {
    ASSERT_VALID_INDEX(index, bonds);
    bonds[index].groups = bondGroups; bonds[index].parameters = parameters;
}

namespace {
/* Recursive-descent evaluator: expr := term (('+'|'-') term)*; term := unary (('*'|'/') unary)*;
   unary := '-' unary | power; power := atom ('^' unary)?; atom := number | name | name '(' args ')' | '(' expr ')' */
struct Evaluator {
    const std::string& text;
    size_t pos;
    const std::map<std::string, double>& variables;
    const std::vector<Vec3>& centers;     /* of the bond's groups, g1.. */
    Evaluator(const std::string& text, const std::map<std::string, double>& variables, const std::vector<Vec3>& centers)
        : text(text), pos(0), variables(variables), centers(centers) {}
    void skip() { while (pos < text.size() && std::isspace((unsigned char) text[pos])) pos++; }
    bool eat(char c) { skip(); if (pos < text.size() && text[pos] == c) { pos++; return true; } return false; }
    [[noreturn]] void fail(const std::string& why) { throw OpenMMException("openmm_shim: cannot evaluate '" + text + "': " + why); }
    double expr() {
        double v = term();
        for (;;) { if (eat('+')) v += term(); else if (eat('-')) v -= term(); else return v; }
    }
    double term() {
        double v = unary();
        for (;;) { if (eat('*')) v *= unary(); else if (eat('/')) v /= unary(); else return v; }
    }
    double unary() { if (eat('-')) return -unary(); if (eat('+')) return unary(); return power(); }
    double power() {
        const double base = atom();
        if (!eat('^')) return base;
        const double e = unary();
        if (e == 2.0) return base * base;      /* Lepton turns x^2 into a Square operation */
        if (e == 3.0) return base * base * base;
        return std::pow(base, e);
    }
    int groupIndex(const std::string& name) {
        if (name.size() < 2 || name[0] != 'g') fail("expected a group name, got " + name);
        const int g = std::atoi(name.c_str() + 1);
        if (g < 1 || g > (int) centers.size()) fail("no group " + name);
        return g - 1;
    }
    std::string name() {
        skip();
        const size_t start = pos;
        while (pos < text.size() && (std::isalnum((unsigned char) text[pos]) || text[pos] == '_')) pos++;
        if (start == pos) fail("unexpected character at " + std::to_string(pos));
        return text.substr(start, pos - start);
    }
    double atom() {
        skip();
        if (eat('(')) { const double v = expr(); if (!eat(')')) fail("missing )"); return v; }
        if (pos < text.size() && (std::isdigit((unsigned char) text[pos]) || text[pos] == '.')) {
            char* end = 0;
            const double v = std::strtod(text.c_str() + pos, &end);
            pos = (size_t) (end - text.c_str());
            return v;
        }
        const std::string id = name();
        if (eat('(')) {
            if (id == "distance") {
                const int a = groupIndex(name()); if (!eat(',')) fail("distance needs two groups");
                const int b = groupIndex(name()); if (!eat(')')) fail("missing )");
                const Vec3 d = centers[a] - centers[b];
                return std::sqrt(d.dot(d));
            }
            if (id == "angle") {
                const int a = groupIndex(name()); if (!eat(',')) fail("angle needs three groups");
                const int b = groupIndex(name()); if (!eat(',')) fail("angle needs three groups");
                const int c = groupIndex(name()); if (!eat(')')) fail("missing )");
                const Vec3 u = centers[a] - centers[b], w = centers[c] - centers[b];
                double cosine = u.dot(w) / std::sqrt(u.dot(u) * w.dot(w));
                cosine = cosine > 1 ? 1 : cosine < -1 ? -1 : cosine;
                return std::acos(cosine);
            }
            const double arg = expr();
            if (!eat(')')) fail("missing )");
            if (id == "step") return arg >= 0 ? 1.0 : 0.0;
            if (id == "sqrt") return std::sqrt(arg);
            if (id == "abs") return std::fabs(arg);
            if (id == "exp") return std::exp(arg);
            if (id == "cos") return std::cos(arg);
            if (id == "sin") return std::sin(arg);
            fail("unknown function " + id);
        }
        if (id.size() >= 2 && (id[0] == 'x' || id[0] == 'y' || id[0] == 'z') && std::isdigit((unsigned char) id[1])) {
            bool digits = true;
            for (size_t i = 1; i < id.size(); i++) digits = digits && std::isdigit((unsigned char) id[i]);
            if (digits) {
                const int g = std::atoi(id.c_str() + 1);
                if (g < 1 || g > (int) centers.size()) fail("no group for " + id);
                return centers[g - 1][id[0] - 'x'];
            }
        }
        std::map<std::string, double>::const_iterator it = variables.find(id);
        if (it == variables.end()) fail("unknown name " + id);
        return it->second;
    }
};
}  // namespace

/* energy only (the plugin evaluates boundary groups with includeForces = false; as integration forces a
   step() term is flat almost everywhere, so these forces contribute nothing to the force pass) */
double CustomCentroidBondForce::shimEvaluate(const std::vector<Vec3>& x, const std::vector<double>& masses, std::vector<Vec3>*) const {
    std::vector<Vec3> allCenters(groups.size());
    for (size_t g = 0; g < groups.size(); g++) {
        const Group& group = groups[g];
        double total = 0;
        for (size_t i = 0; i < group.particles.size(); i++) total += group.weights.empty() ? masses[group.particles[i]] : group.weights[i];
        Vec3 c;
        for (size_t i = 0; i < group.particles.size(); i++) {
            const double w = (group.weights.empty() ? masses[group.particles[i]] : group.weights[i]) / total;   /* normalised weights */
            c += x[group.particles[i]] * w;
        }
        allCenters[g] = c;
    }
    double energy = 0;
    for (const Bond& bond : bonds) {
        std::vector<Vec3> centers;
        for (int g : bond.groups) centers.push_back(allCenters.at(g));
        std::map<std::string, double> variables;
        for (size_t p = 0; p < bondParameterNames.size(); p++) variables[bondParameterNames[p]] = bond.parameters.at(p);
        Evaluator ev(energyExpression, variables, centers);
        const double value = ev.expr();
        ev.skip();
        if (ev.pos != energyExpression.size()) ev.fail("trailing characters");
        energy += value;
    }
    return energy;
}

/* ---- State ---------------------------------------------------------------------------------------------- */

const std::vector<Vec3>& State::getPositions() const {
    if (!(types & Positions)) throw OpenMMException("Invoked getPositions() on a State which does not contain positions.");
    return positions;
}
const std::vector<Vec3>& State::getVelocities() const {
    if (!(types & Velocities)) throw OpenMMException("Invoked getVelocities() on a State which does not contain velocities.");
    return velocities;
}
const std::vector<Vec3>& State::getForces() const {
    if (!(types & Forces)) throw OpenMMException("Invoked getForces() on a State which does not contain forces.");
    return forces;
}
double State::getKineticEnergy() const {
    if (!(types & Energy)) throw OpenMMException("Invoked getKineticEnergy() on a State which does not contain energies.");
    return ke;
}
double State::getPotentialEnergy() const {
    if (!(types & Energy)) throw OpenMMException("Invoked getPotentialEnergy() on a State which does not contain energies.");
    return pe;
}

void XmlSerializer::serializeState(const State& state, const std::string& rootName, std::ostream& stream) {
    char buf[256];
    stream << "<?xml version=\"1.0\" ?>\n";
    std::snprintf(buf, sizeof buf, "<%s openmmVersion=\"shim\" stepCount=\"%lld\" time=\"%.17g\" type=\"State\" version=\"1\">\n", rootName.c_str(), state.stepCount, state.time);
    stream << buf;
    stream << "\t<PeriodicBoxVectors>\n";
    for (int r = 0; r < 3; r++) {
        std::snprintf(buf, sizeof buf, "\t\t<%c x=\"%.17g\" y=\"%.17g\" z=\"%.17g\"/>\n", "ABC"[r], state.box[r][0], state.box[r][1], state.box[r][2]);
        stream << buf;
    }
    stream << "\t</PeriodicBoxVectors>\n";
    if (state.types & State::Positions) {
        stream << "\t<Positions>\n";
        for (const Vec3& p : state.positions) { std::snprintf(buf, sizeof buf, "\t\t<Position x=\"%.17g\" y=\"%.17g\" z=\"%.17g\"/>\n", p[0], p[1], p[2]); stream << buf; }
        stream << "\t</Positions>\n";
    }
    if (state.types & State::Velocities) {
        stream << "\t<Velocities>\n";
        for (const Vec3& v : state.velocities) { std::snprintf(buf, sizeof buf, "\t\t<Velocity x=\"%.17g\" y=\"%.17g\" z=\"%.17g\"/>\n", v[0], v[1], v[2]); stream << buf; }
        stream << "\t</Velocities>\n";
    }
    stream << "</" << rootName << ">\n";
}

/* ---- Platform ------------------------------------------------------------------------------------------ */

namespace {
std::vector<Platform*>& platforms() { static std::vector<Platform*> list; return list; }
std::string canonicalProperty(const std::string& property) {
    if (property == "CudaPrecision") return "Precision";            /* deprecated aliases OpenMM still accepts */
    if (property == "CudaDeviceIndex") return "DeviceIndex";
    return property;
}
}  // namespace

const std::string& Platform::getPropertyDefaultValue(const std::string& property) const {
    std::map<std::string, std::string>::const_iterator it = defaultProperties.find(canonicalProperty(property));
    if (it == defaultProperties.end()) throw OpenMMException("getPropertyDefaultValue: Illegal property name");
    return it->second;
}
void Platform::setPropertyDefaultValue(const std::string& property, const std::string& value) {
    const std::string key = canonicalProperty(property);
    if (defaultProperties.find(key) == defaultProperties.end()) throw OpenMMException("setPropertyDefaultValue: Illegal property name");
    defaultProperties[key] = value;
}
void Platform::registerKernelFactory(const std::string& name, KernelFactory* factory) { kernelFactories[name] = factory; }
bool Platform::supportsKernels(const std::vector<std::string>& kernelNames) const {
    for (const std::string& n : kernelNames) if (kernelFactories.find(n) == kernelFactories.end()) return false;
    return true;
}
Kernel Platform::createKernel(const std::string& name, ContextImpl& context) const {
    std::map<std::string, KernelFactory*>::const_iterator it = kernelFactories.find(name);
    if (it == kernelFactories.end()) throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
    return Kernel(it->second->createKernelImpl(name, *this, context));
}
void Platform::registerPlatform(Platform* platform) { platforms().push_back(platform); }
int Platform::getNumPlatforms() { return (int) platforms().size(); }
Platform& Platform::getPlatform(int index) { return *platforms().at(index); }
Platform& Platform::getPlatformByName(const std::string& name) {
    for (Platform* p : platforms()) if (p->getName() == name) return *p;
    throw OpenMMException("There is no registered Platform called \"" + name + "\"");
}

/* ---- Context / ContextImpl ----------------------------------------------------------------------------- */

ContextImpl::ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform, const std::map<std::string, std::string>& properties)
    : shimContext(0), shimForceEvaluations(0), shimEnergyEvaluations(0), shimForceNs(0), shimEnergyNs(0), owner(owner), system(system), integrator(integrator), platform(platform), platformData(0), lastForceGroups(-1) {
    if (system.getNumParticles() == 0) throw OpenMMException("Cannot create a Context for a System with no particles");
    for (int i = 0; i < system.getNumConstraints(); i++) {
        int p1, p2; double d;
        system.getConstraintParameters(i, p1, p2, d);
        const double m1 = system.getParticleMass(p1), m2 = system.getParticleMass(p2);
        if ((m1 == 0.0 || m2 == 0.0) && (m1 != 0.0 || m2 != 0.0))
            throw OpenMMException("A constraint cannot involve a massless particle");
    }
    if (!platform->supportsKernels(integrator.getKernelNames())) throw OpenMMException("Specified a Platform for a Context which does not support all required kernels");
    platform->contextCreated(*this, properties);
    integrator.initialize(*this);
}
ContextImpl::~ContextImpl() { platform->contextDestroyed(*this); }
double ContextImpl::getTime() const { return shimContext->getTime(); }
void ContextImpl::setTime(double t) { shimContext->setTime(t); }
long long ContextImpl::getStepCount() const { return shimContext->getStepCount(); }

double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    lastForceGroups = groups;
    /* timing runs (bench.py's adapter leg): OPENMM_SHIM_FREEZE_FORCES=1 leaves the force buffer as the first evaluation
       wrote it, so that integrator.step(n) times the integrator kernels and not this stand-in's host-side force pass;
       energy-only passes (what the reference's kernels ask for every step) are always evaluated, and timed */
    static const bool freeze = [] { const char* f = std::getenv("OPENMM_SHIM_FREEZE_FORCES"); return f && f[0] == '1'; }();
    if (freeze && includeForces && !includeEnergy && shimForceEvaluations > 0) { shimForceEvaluations++; return 0.0; }
    const auto t0 = std::chrono::steady_clock::now();
    struct Timer {
        const std::chrono::steady_clock::time_point t0; long long& into;
        ~Timer() { into += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
    } timer{t0, includeForces ? shimForceNs : shimEnergyNs};
    /* Energy-only pass of a few force groups (what the reference's kernels ask for every step, CudaSeekr2Kernels.cpp:248,
       :527,:552): when every Force in the mask says which particles its energy depends on, only those positions are fetched
       from the device (one small gather kernel + one small blocking copy -- the shape of what OpenMM does: a few small kernels
       and a blocking read-back of the energy) instead of the whole position array. */
    std::vector<Vec3>& x = shimPositions;
    if (shimMasses.size() != (size_t) system.getNumParticles()) {
        shimMasses.resize(system.getNumParticles());
        for (int i = 0; i < system.getNumParticles(); i++) shimMasses[i] = system.getParticleMass(i);
    }
    const std::vector<double>& masses = shimMasses;
    bool fetched = false;
    if (!includeForces) {
        auto it = shimEnergySubsets.find(groups);
        if (it == shimEnergySubsets.end()) {
            std::vector<int> subset;
            bool all = false;
            for (int i = 0; i < system.getNumForces() && !all; i++) {
                const Force& force = system.getForce(i);
                if (((groups >> force.getForceGroup()) & 1) && !force.shimEnergyParticles(subset)) all = true;
            }
            std::sort(subset.begin(), subset.end());
            subset.erase(std::unique(subset.begin(), subset.end()), subset.end());
            if (all) subset.assign(1, -1);
            it = shimEnergySubsets.emplace(groups, subset).first;
        }
        if (it->second.empty() || it->second[0] >= 0) {
            x.resize(system.getNumParticles());
            fetched = it->second.empty() || shimContext->getForcePositionsOf(it->second, x);
        }
    }
    if (!fetched) shimContext->getForcePositions(x);
    std::vector<Vec3> f(includeForces ? x.size() : 0);
    double energy = 0;
    for (int i = 0; i < system.getNumForces(); i++) {
        const Force& force = system.getForce(i);
        if (!((groups >> force.getForceGroup()) & 1)) continue;
        energy += force.shimEvaluate(x, masses, includeForces ? &f : 0);
    }
    if (includeForces) { shimContext->setForces(f); shimForceEvaluations++; }
    if (includeEnergy) shimEnergyEvaluations++;
    return includeEnergy ? energy : 0.0;
}

Context::Context(const System& system, Integrator& integrator) : system(system), integrator(integrator), platform(&Platform::getPlatformByName("CUDA")), impl(0) { create(); }
Context::Context(const System& system, Integrator& integrator, Platform& platform) : system(system), integrator(integrator), platform(&platform), impl(0) { create(); }
Context::Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties)
    : system(system), integrator(integrator), platform(&platform), properties(properties), impl(0) { create(); }
void Context::create() {
    if (integrator.owner != 0 && integrator.owner != this) throw OpenMMException("This Integrator is already bound to a context");
    integrator.owner = this;
    try {
        impl = new ContextImpl(*this, system, integrator, platform, properties);
    } catch (...) {
        integrator.owner = 0;
        integrator.context = 0;
        throw;
    }
}
Context::~Context() {
    integrator.cleanup();
    delete impl;
    integrator.owner = 0;
    integrator.context = 0;
}
void Context::reinitialize(bool preserveState) {
    if (preserveState) throw OpenMMException("openmm_shim: reinitialize(preserveState=true) is not implemented");
    integrator.cleanup();
    delete impl;
    impl = 0;
    impl = new ContextImpl(*this, system, integrator, platform, properties);
}
double Context::getTime() const { return impl->getTime(); }
void Context::setTime(double time) { impl->setTime(time); }
void Context::setPositions(const std::vector<Vec3>& positions) {
    if ((int) positions.size() != system.getNumParticles()) throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
    impl->shimContext->setPositions(positions);
}
void Context::setVelocities(const std::vector<Vec3>& velocities) {
    if ((int) velocities.size() != system.getNumParticles()) throw OpenMMException("Called setVelocities() on a Context with the wrong number of velocities");
    impl->shimContext->setVelocities(velocities);
}
void Context::setVelocitiesToTemperature(double temperature, int randomSeed) {
    std::mt19937_64 engine((unsigned long long) randomSeed + 0x9E3779B97F4A7C15ull);
    std::normal_distribution<double> normal(0.0, 1.0);
    std::vector<Vec3> v(system.getNumParticles());
    for (int i = 0; i < system.getNumParticles(); i++) {
        const double m = system.getParticleMass(i);
        if (m != 0) {
            const double s = std::sqrt(BOLTZ * temperature / m);
            v[i] = Vec3(s * normal(engine), s * normal(engine), s * normal(engine));
        }
    }
    impl->shimContext->setVelocities(v);
    impl->shimContext->applyVelocityConstraints(1e-5);
}
void Context::setPeriodicBoxVectors(const Vec3&, const Vec3&, const Vec3&) {}
void Context::applyConstraints(double tol) { impl->shimContext->applyConstraints(tol); }
void Context::applyVelocityConstraints(double tol) { impl->shimContext->applyVelocityConstraints(tol); }

State Context::getState(int types, bool, int groups) const {
    State s;
    s.types = types;
    s.time = impl->getTime();
    s.stepCount = impl->getStepCount();
    system.getDefaultPeriodicBoxVectors(s.box[0], s.box[1], s.box[2]);
    if (types & State::Energy) {
        /* OpenMM: forces are needed first when the integrator's kinetic energy is time-shifted */
        s.pe = impl->calcForcesAndEnergy(integrator.kineticEnergyRequiresForce() || (types & State::Forces), true, groups);
        s.ke = integrator.computeKineticEnergy();
    } else if (types & State::Forces) {
        impl->calcForcesAndEnergy(true, false, groups);
    }
    if (types & State::Forces) impl->shimContext->getForces(s.forces);
    if (types & State::Positions) impl->shimContext->getPositions(s.positions);
    if (types & State::Velocities) impl->shimContext->getVelocities(s.velocities);
    return s;
}

}  // namespace OpenMM
