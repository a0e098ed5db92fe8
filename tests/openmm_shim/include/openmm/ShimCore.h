/* tests/openmm_shim -- TEST INFRASTRUCTURE ONLY.  NOT OpenMM.
 *
 * A minimal stand-in for the part of the OpenMM C++ API that the SEEKR2 plugin touches, written from the
 * way the reference uses that API (seekr2plugin/openmmapi/src, platforms/cuda/src/CudaSeekr2Kernels.cpp,
 * platforms/cuda/tests) because OpenMM itself is absent from this environment.  It exists so that
 *   (1) the reference's own, unmodified integrator classes and CUDA-platform host code, and
 *   (2) this repository's platforms/b200 adapter
 * can be compiled and run on the B200 against the same "OpenMM", and compared.  Class and method names
 * follow OpenMM; everything behind them (force evaluation, constraint solver, random pool, serializer) is a
 * small host-side implementation that makes no attempt at OpenMM's performance or generality. */
#ifndef OPENMM_SHIM_CORE_H_
#define OPENMM_SHIM_CORE_H_

#include <cmath>
#include <exception>
#include <iosfwd>
#include <map>
#include <string>
#include <utility>
#include <vector>

#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif

namespace OpenMM {

class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    bool operator==(const Vec3& o) const { return data[0] == o[0] && data[1] == o[1] && data[2] == o[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    Vec3 operator+(const Vec3& o) const { return Vec3(data[0] + o[0], data[1] + o[1], data[2] + o[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(data[0] - o[0], data[1] - o[1], data[2] - o[2]); }
    Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }
    Vec3 operator*(double s) const { return Vec3(data[0] * s, data[1] * s, data[2] * s); }
    Vec3 operator/(double s) const { return Vec3(data[0] / s, data[1] / s, data[2] / s); }
    Vec3& operator+=(const Vec3& o) { data[0] += o[0]; data[1] += o[1]; data[2] += o[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { data[0] -= o[0]; data[1] -= o[1]; data[2] -= o[2]; return *this; }
    Vec3& operator*=(double s) { data[0] *= s; data[1] *= s; data[2] *= s; return *this; }
    double dot(const Vec3& o) const { return data[0] * o[0] + data[1] * o[1] + data[2] * o[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(data[1] * o[2] - data[2] * o[1], data[2] * o[0] - data[0] * o[2], data[0] * o[1] - data[1] * o[0]);
    }
private:
    double data[3];
};
inline Vec3 operator*(double s, const Vec3& v) { return v * s; }
std::ostream& operator<<(std::ostream& o, const Vec3& v);

class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string& message) : message(message) {}
    ~OpenMMException() throw() {}
    const char* what() const throw() { return message.c_str(); }
private:
    std::string message;
};

/* ---- System and forces ------------------------------------------------------------------------------ */

class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int group) { forceGroup = group; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
    /* shim: energy of this force at `positions`; adds the forces to `forces` when it is not NULL */
    virtual double shimEvaluate(const std::vector<Vec3>& positions, const std::vector<double>& masses, std::vector<Vec3>* forces) const = 0;
    /* shim: true if the ENERGY depends on the listed particles only (appended to `particles`); lets an energy-only pass of
       a force group fetch those positions instead of all of them */
    virtual bool shimEnergyParticles(std::vector<int>& particles) const { return false; }
private:
    int forceGroup;
};

class HarmonicBondForce : public Force {
public:
    int addBond(int particle1, int particle2, double length, double k);
    int getNumBonds() const { return (int) bonds.size(); }
    void getBondParameters(int index, int& particle1, int& particle2, double& length, double& k) const;
    double shimEvaluate(const std::vector<Vec3>&, const std::vector<double>&, std::vector<Vec3>*) const;
private:
    struct Bond { int p1, p2; double length, k; };
    std::vector<Bond> bonds;
};

class NonbondedForce : public Force {
public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2, Ewald = 3, PME = 4, LJPME = 5 };
    int addParticle(double charge, double sigma, double epsilon);
    int getNumParticles() const { return (int) particles.size(); }
    void getParticleParameters(int index, double& charge, double& sigma, double& epsilon) const;
    NonbondedMethod getNonbondedMethod() const { return NoCutoff; }
    double shimEvaluate(const std::vector<Vec3>&, const std::vector<double>&, std::vector<Vec3>*) const;
private:
    struct Particle { double charge, sigma, epsilon; };
    std::vector<Particle> particles;
};

/* CustomCentroidBondForce: groups of particles, bonds between group centroids, one algebraic energy
   expression.  The shim evaluates the expression itself (numbers, + - * / ^, parentheses, unary minus,
   step(), sqrt(), abs(), distance(gi,gj), angle(gi,gj,gk), xN/yN/zN, per-bond parameters). */
class CustomCentroidBondForce : public Force {
public:
    CustomCentroidBondForce(int numGroups, const std::string& energy) : groupsPerBond(numGroups), energyExpression(energy) {}
    int getNumGroupsPerBond() const { return groupsPerBond; }
    int getNumGroups() const { return (int) groups.size(); }
    int getNumBonds() const { return (int) bonds.size(); }
    int getNumPerBondParameters() const { return (int) bondParameterNames.size(); }
    int getNumGlobalParameters() const { return 0; }
    const std::string& getEnergyFunction() const { return energyExpression; }
    void setEnergyFunction(const std::string& energy) { energyExpression = energy; }
    int addPerBondParameter(const std::string& name) { bondParameterNames.push_back(name); return (int) bondParameterNames.size() - 1; }
    const std::string& getPerBondParameterName(int index) const { return bondParameterNames.at(index); }
    int addGroup(const std::vector<int>& particles, const std::vector<double>& weights = std::vector<double>());
    void getGroupParameters(int index, std::vector<int>& particles, std::vector<double>& weights) const;
    int addBond(const std::vector<int>& groups, const std::vector<double>& parameters = std::vector<double>());
    void getBondParameters(int index, std::vector<int>& groups, std::vector<double>& parameters) const;
    void setBondParameters(int index, const std::vector<int>& groups, const std::vector<double>& parameters = std::vector<double>());
    void setUsesPeriodicBoundaryConditions(bool periodic) { usePeriodic = periodic; }
    bool usesPeriodicBoundaryConditions() const { return usePeriodic; }
    double shimEvaluate(const std::vector<Vec3>&, const std::vector<double>&, std::vector<Vec3>*) const;
    bool shimEnergyParticles(std::vector<int>& particles) const {
        for (const Group& g : groups) particles.insert(particles.end(), g.particles.begin(), g.particles.end());
        return true;
    }
private:
    struct Group { std::vector<int> particles; std::vector<double> weights; };
    struct Bond { std::vector<int> groups; std::vector<double> parameters; };
    int groupsPerBond;
    std::string energyExpression;
    std::vector<std::string> bondParameterNames;
    std::vector<Group> groups;
    std::vector<Bond> bonds;
    bool usePeriodic = false;
};

class System {
public:
    System();
    ~System();
    int getNumParticles() const { return (int) masses.size(); }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size() - 1; }
    double getParticleMass(int index) const { return masses.at(index); }
    void setParticleMass(int index, double mass) { masses.at(index) = mass; }
    bool isVirtualSite(int) const { return false; }
    int getNumConstraints() const { return (int) constraints.size(); }
    int addConstraint(int particle1, int particle2, double distance);
    void getConstraintParameters(int index, int& particle1, int& particle2, double& distance) const;
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size() - 1; }   /* takes ownership */
    int getNumForces() const { return (int) forces.size(); }
    const Force& getForce(int index) const { return *forces.at(index); }
    Force& getForce(int index) { return *forces.at(index); }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
    bool usesPeriodicBoundaryConditions() const { return false; }
private:
    System(const System&);
    System& operator=(const System&);
    struct Constraint { int p1, p2; double distance; };
    std::vector<double> masses;
    std::vector<Constraint> constraints;
    std::vector<Force*> forces;
    Vec3 box[3];
};

/* ---- State ------------------------------------------------------------------------------------------- */

class State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16, ParameterDerivatives = 32, IntegratorParameters = 64 };
    State() : time(0), stepCount(0), ke(0), pe(0), types(0) {}
    double getTime() const { return time; }
    long long getStepCount() const { return stepCount; }
    const std::vector<Vec3>& getPositions() const;
    const std::vector<Vec3>& getVelocities() const;
    const std::vector<Vec3>& getForces() const;
    double getKineticEnergy() const;
    double getPotentialEnergy() const;
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    int getDataTypes() const { return types; }
    /* shim: filled by Context::getState */
    double time; long long stepCount; double ke, pe; int types;
    std::vector<Vec3> positions, velocities, forces;
    Vec3 box[3];
};

/* ---- kernels and platforms ------------------------------------------------------------------------- */

class Platform;
class ContextImpl;
class Context;

class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), referenceCount(0) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int referenceCount;
};

class Kernel {
public:
    Kernel() : impl(0) {}
    Kernel(KernelImpl* impl) : impl(impl) { if (impl) impl->referenceCount++; }
    Kernel(const Kernel& copy) : impl(copy.impl) { if (impl) impl->referenceCount++; }
    ~Kernel() { release(); }
    Kernel& operator=(const Kernel& copy) {
        if (copy.impl) copy.impl->referenceCount++;
        release();
        impl = copy.impl;
        return *this;
    }
    std::string getName() const { return impl->getName(); }
    const KernelImpl& getImpl() const { return *impl; }
    KernelImpl& getImpl() { return *impl; }
    template <class T> const T& getAs() const { return dynamic_cast<const T&>(*impl); }
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
private:
    void release() { if (impl && --impl->referenceCount == 0) delete impl; impl = 0; }
    KernelImpl* impl;
};

class KernelFactory {
public:
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
    virtual ~KernelFactory() {}
};

class Platform {
public:
    virtual ~Platform() {}
    virtual const std::string& getName() const = 0;
    virtual double getSpeed() const { return 1.0; }
    virtual bool supportsDoublePrecision() const { return true; }
    const std::string& getPropertyDefaultValue(const std::string& property) const;
    void setPropertyDefaultValue(const std::string& property, const std::string& value);
    virtual void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {}
    virtual void contextDestroyed(ContextImpl& context) const {}
    void registerKernelFactory(const std::string& name, KernelFactory* factory);
    bool supportsKernels(const std::vector<std::string>& kernelNames) const;
    Kernel createKernel(const std::string& name, ContextImpl& context) const;
    static void registerPlatform(Platform* platform);
    static int getNumPlatforms();
    static Platform& getPlatform(int index);
    static Platform& getPlatformByName(const std::string& name);
protected:
    std::map<std::string, std::string> defaultProperties;
private:
    std::map<std::string, KernelFactory*> kernelFactories;
};

/* ---- Integrator, Context ------------------------------------------------------------------------------ */

class Integrator {
public:
    Integrator() : context(0), owner(0), stepSize(0), constraintTol(1e-5) {}
    virtual ~Integrator() {}
    virtual double getStepSize() const { return stepSize; }
    virtual void setStepSize(double size) { stepSize = size; }
    virtual double getConstraintTolerance() const { return constraintTol; }
    virtual void setConstraintTolerance(double tol) { constraintTol = tol; }
    virtual void step(int steps) = 0;
    virtual int getIntegrationForceGroups() const { return 0xFFFFFFFF; }
protected:
    friend class Context;
    friend class ContextImpl;
    ContextImpl* context;
    Context* owner;
    virtual void initialize(ContextImpl& context) = 0;
    virtual void cleanup() {}
    virtual std::vector<std::string> getKernelNames() = 0;
    virtual void stateChanged(State::DataType) {}
    virtual double computeKineticEnergy() = 0;
    virtual bool kineticEnergyRequiresForce() const { return true; }
private:
    double stepSize, constraintTol;
};

/* exists only because the reference's tests include the header */
class VerletIntegrator : public Integrator {
public:
    explicit VerletIntegrator(double stepSize) { setStepSize(stepSize); }
    void step(int) { throw OpenMMException("openmm_shim: VerletIntegrator is not implemented"); }
protected:
    void initialize(ContextImpl&) {}
    std::vector<std::string> getKernelNames() { return std::vector<std::string>(); }
    double computeKineticEnergy() { return 0.0; }
};

class Context {
public:
    Context(const System& system, Integrator& integrator);
    Context(const System& system, Integrator& integrator, Platform& platform);
    Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties);
    ~Context();
    const System& getSystem() const { return system; }
    const Integrator& getIntegrator() const { return integrator; }
    Integrator& getIntegrator() { return integrator; }
    const Platform& getPlatform() const { return *platform; }
    Platform& getPlatform() { return *platform; }
    State getState(int types, bool enforcePeriodicBox = false, int groups = 0xFFFFFFFF) const;
    void setTime(double time);
    double getTime() const;
    void setPositions(const std::vector<Vec3>& positions);
    void setVelocities(const std::vector<Vec3>& velocities);
    void setVelocitiesToTemperature(double temperature, int randomSeed = 0);
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c);
    void applyConstraints(double tol);
    void applyVelocityConstraints(double tol);
    void reinitialize(bool preserveState = false);
    ContextImpl& getImpl() { return *impl; }
private:
    friend class ContextImpl;
    Context(const Context&);
    Context& operator=(const Context&);
    void create();
    const System& system;
    Integrator& integrator;
    Platform* platform;
    std::map<std::string, std::string> properties;
    ContextImpl* impl;
};

/* shim: what a platform must provide to ContextImpl (the CUDA stand-in implements it) */
class ShimPlatformContext {
public:
    virtual ~ShimPlatformContext() {}
    virtual double getTime() const = 0;
    virtual void setTime(double t) = 0;
    virtual long long getStepCount() const = 0;
    virtual void getPositions(std::vector<Vec3>& positions) = 0;             /* as Context::getState returns them */
    virtual void getForcePositions(std::vector<Vec3>& positions) = 0;        /* what force kernels read */
    /* the same for the listed particles only (entries of the others are left alone); false = not supported, fetch all */
    virtual bool getForcePositionsOf(const std::vector<int>& particles, std::vector<Vec3>& positions) { return false; }
    virtual void setPositions(const std::vector<Vec3>& positions) = 0;
    virtual void getVelocities(std::vector<Vec3>& velocities) = 0;
    virtual void setVelocities(const std::vector<Vec3>& velocities) = 0;
    virtual void setForces(const std::vector<Vec3>& forces) = 0;
    virtual void getForces(std::vector<Vec3>& forces) = 0;
    virtual void applyConstraints(double tol) = 0;
    virtual void applyVelocityConstraints(double tol) = 0;
};

class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform, const std::map<std::string, std::string>& properties);
    ~ContextImpl();
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    const Integrator& getIntegrator() const { return integrator; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    const void* getPlatformData() const { return platformData; }
    void setPlatformData(void* data) { platformData = data; }
    double getTime() const;
    void setTime(double t);
    long long getStepCount() const;
    void updateContextState() {}
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    int getLastForceGroups() const { return lastForceGroups; }
    /* shim */
    ShimPlatformContext* shimContext;
    long long shimForceEvaluations, shimEnergyEvaluations;
    std::map<int, std::vector<int> > shimEnergySubsets;     /* force-group mask -> particles its energy depends on ({-1} = all) */
    std::vector<Vec3> shimPositions;
    std::vector<double> shimMasses;
    long long shimForceNs, shimEnergyNs;        /* host time spent in force passes / energy-only passes */
    double calcKineticEnergy() { return integrator.computeKineticEnergy(); }   /* OpenMM: ContextImpl::calcKineticEnergy() */
private:
    friend class Context;
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    void* platformData;
    int lastForceGroups;
};

}  // namespace OpenMM
#endif
