/* openmm_shim (NOT OpenMM): stand-in for the CUDA platform classes the SEEKR2 plugin uses -- CudaArray,
 * CudaContext, CudaIntegrationUtilities, CudaPlatform -- with the buffer layouts of OpenMM's CUDA platform
 * (posq real4, posqCorrection real4, velm mixed4, force int64 SoA x 2^32, posDelta mixed4, random float4
 * pool of 32 x paddedNumAtoms, stepSize mixed2).  Forces, constraints and random numbers are produced on the
 * host and uploaded: correctness scaffolding, not a simulation engine.  See ShimCore.h. */
#ifndef OPENMM_SHIM_CUDA_H_
#define OPENMM_SHIM_CUDA_H_
#include "openmm/ShimCore.h"
#include <cassert>      /* the real CudaContext.h pulls these in; the reference relies on that */
#include <sstream>
#include <cuda.h>
#include <cuda_runtime.h>
#include <functional>
#include <string>
#include <vector>

namespace OpenMM {

class CudaContext;

class CudaArray {
public:
    CudaArray() : cu(0), pointer(0), size(0), elementSize(0) {}
    CudaArray(CudaContext& context, size_t size, int elementSize, const std::string& name) : cu(0), pointer(0), size(0), elementSize(0) { initialize(context, size, elementSize, name); }
    ~CudaArray();
    template <class T> static CudaArray* create(CudaContext& context, size_t size, const std::string& name) { return new CudaArray(context, size, sizeof(T), name); }
    void initialize(CudaContext& context, size_t size, int elementSize, const std::string& name);
    template <class T> void initialize(CudaContext& context, size_t size, const std::string& name) { initialize(context, size, sizeof(T), name); }
    bool isInitialized() const { return pointer != 0; }
    size_t getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    const std::string& getName() const { return name; }
    CUdeviceptr& getDevicePointer() { return pointer; }
    void upload(const void* data, bool blocking = true);
    void download(void* data, bool blocking = true) const;
    /* convert = true: a vector<double> may feed a float array and vice versa (OpenMM's ArrayInterface) */
    template <class T> void upload(const std::vector<T>& data, bool convert = false) {
        if (convert && sizeof(T) != (size_t) elementSize) { uploadConverted(&data[0], sizeof(T), data.size()); return; }
        if (sizeof(T) != (size_t) elementSize || data.size() != size) throw OpenMMException("Error uploading array " + name + ": The specified vector does not match the size of the array");
        upload(&data[0], true);
    }
    template <class T> void download(std::vector<T>& data) const {
        if (sizeof(T) != (size_t) elementSize) throw OpenMMException("Error downloading array " + name + ": The specified vector has the wrong element size");
        data.resize(size);
        download(&data[0], true);
    }
private:
    CudaArray(const CudaArray&);
    CudaArray& operator=(const CudaArray&);
    void uploadConverted(const void* data, size_t hostElementSize, size_t count);
    CudaContext* cu;
    CUdeviceptr pointer;
    size_t size;
    int elementSize;
    std::string name;
};

class CudaIntegrationUtilities {
public:
    explicit CudaIntegrationUtilities(CudaContext& context, const System& system);
    CudaArray& getPosDelta() { return posDelta; }
    CudaArray& getRandom() { return random; }
    CudaArray& getStepSize() { return stepSize; }
    void setNextStepSize(double size);
    double getLastStepSize() const { return lastStepSize; }
    void applyConstraints(double tol);
    void applyVelocityConstraints(double tol);
    void initRandomNumberGenerator(unsigned int randomNumberSeed);
    int prepareRandomNumbers(int numValues);
    void computeVirtualSites() {}
    double computeKineticEnergy(double timeShift);
    /* shim: host-side constraint passes on plain vectors (used for setVelocitiesToTemperature and KE too) */
    void shimShake(const std::vector<Vec3>& x, std::vector<Vec3>& delta, double tol) const;
    void shimRattle(const std::vector<Vec3>& x, std::vector<Vec3>& v, double tol) const;
    long long shimPoolFills;
private:
    CudaContext& context;
    const System& system;
    CudaArray posDelta, random, stepSize;
    int randomPos;
    unsigned int seed;
    bool seeded;
    double lastStepSize;
};

class CudaPlatform : public Platform {
public:
    class PlatformData;
    CudaPlatform();
    const std::string& getName() const { static const std::string name = "CUDA"; return name; }
    static const std::string& CudaDeviceIndex() { static const std::string key = "DeviceIndex"; return key; }
    static const std::string& CudaPrecision() { static const std::string key = "Precision"; return key; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const;
    void contextDestroyed(ContextImpl& context) const;
};

class CudaPlatform::PlatformData {
public:
    PlatformData(ContextImpl* context, const System& system, const std::string& precision, int deviceIndex);
    ~PlatformData();
    void initializeContexts(const System& system) {}
    ContextImpl* context;
    std::vector<CudaContext*> contexts;
};

class CudaContext : public ShimPlatformContext {
public:
    CudaContext(const System& system, int deviceIndex, const std::string& precision, CudaPlatform::PlatformData& platformData);
    ~CudaContext();
    void setAsCurrent();
    int getDeviceIndex() const { return deviceIndex; }
    CUstream getCurrentStream() { return stream; }
    CudaPlatform::PlatformData& getPlatformData() { return platformData; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return useMixed; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    CudaArray& getPosq() { return posq; }
    CudaArray& getPosqCorrection() { return posqCorrection; }
    CudaArray& getVelm() { return velm; }
    CudaArray& getForce() { return force; }
    const std::vector<int>& getAtomIndex() const { return atomIndex; }
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
    CudaIntegrationUtilities& getIntegrationUtilities() { return *integration; }
    /* NVRTC, with the typedef prelude OpenMM's createModule() puts in front of every kernel source */
    CUmodule createModule(const std::string source, const std::map<std::string, std::string>& defines, const char* optimizationFlags = NULL);
    CUmodule createModule(const std::string source, const char* optimizationFlags = NULL) { return createModule(source, std::map<std::string, std::string>(), optimizationFlags); }
    CUfunction getKernel(CUmodule& module, const std::string& name);
    void executeKernel(CUfunction kernel, void** arguments, int threads, int blockSize = -1, unsigned int sharedSize = 0);
    static std::string getErrorString(CUresult result);
    double getTime() const { return time; }
    void setTime(double t) { time = t; }
    long long getStepCount() const { return stepCount; }
    void setStepCount(long long steps) { stepCount = steps; }
    void reorderAtoms();
    /* ShimPlatformContext */
    void getPositions(std::vector<Vec3>& positions);
    void getForcePositions(std::vector<Vec3>& positions);
    bool getForcePositionsOf(const std::vector<int>& particles, std::vector<Vec3>& positions);
    void setPositions(const std::vector<Vec3>& positions);
    void getVelocities(std::vector<Vec3>& velocities);
    void setVelocities(const std::vector<Vec3>& velocities);
    void setForces(const std::vector<Vec3>& forces);
    void getForces(std::vector<Vec3>& forces);
    void applyConstraints(double tol);
    void applyVelocityConstraints(double tol);
    /* shim */
    std::vector<double> shimInverseMasses() const;
    const System& getSystem() const { return system; }
    long long shimKernelLaunches, shimReorders = 0;
    int shimIndexOfAtom(int particle) const { return indexOfAtom[particle]; }
private:
    int at(int particle) const { return indexOfAtom[particle]; }
    std::vector<int> indexOfAtom;            /* inverse of atomIndex */
    /* gather of a few positions for energy-only passes (getForcePositionsOf) */
    CudaArray gatherIndex, gatherOut;
    std::vector<int> gatherParticles;
    long long gatherReorders = -1;
    CUfunction gatherKernel = 0;
    int reorderEvery, reorderCalls;
    const System& system;
    CudaPlatform::PlatformData& platformData;
    int deviceIndex, numAtoms, paddedNumAtoms, multiprocessors, computeCapability;
    bool useDouble, useMixed;
    CUstream stream;
    CudaArray posq, posqCorrection, velm, force;
    std::vector<int> atomIndex;
    Vec3 box[3];
    CudaIntegrationUtilities* integration;
    double time;
    long long stepCount;
    std::vector<CUmodule> modules;
};

/* shim: test hooks.  A random source replaces the shim's own generator when a parity test injects numbers:
   fill(pool, count, fillIndex) must write `count` float4 values (x, y, z, w) for the fillIndex-th refill. */
typedef std::function<void(float* pool4, size_t count, long long fillIndex)> ShimRandomSource;
void shimSetRandomSource(ShimRandomSource source);

}  // namespace OpenMM
#endif
