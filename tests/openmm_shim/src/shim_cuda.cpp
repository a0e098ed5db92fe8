/* tests/openmm_shim -- TEST INFRASTRUCTURE ONLY.  NOT OpenMM (see include/openmm/cuda/ShimCuda.h).
 * CUDA side: device arrays in the OpenMM CUDA-platform layouts, NVRTC module creation with the typedef prelude
 * OpenMM injects, kernel launches, and host-side stand-ins for the random pool, SHAKE / RATTLE and the
 * time-shifted kinetic energy. */
#include "openmm/cuda/ShimCuda.h"
#include <nvrtc.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sstream>

namespace OpenMM {

namespace {
void checkRt(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw OpenMMException(std::string("openmm_shim: ") + what + ": " + cudaGetErrorString(e));
}
void checkDrv(CUresult r, const char* what) {
    if (r != CUDA_SUCCESS) throw OpenMMException(std::string("openmm_shim: ") + what + ": " + CudaContext::getErrorString(r));
}
ShimRandomSource& randomSource() { static ShimRandomSource source; return source; }
}  // namespace

void shimSetRandomSource(ShimRandomSource source) { randomSource() = source; }

/* ---- CudaArray ------------------------------------------------------------------------------------------ */

CudaArray::~CudaArray() { if (pointer) cudaFree((void*) pointer); }

void CudaArray::initialize(CudaContext& context, size_t newSize, int newElementSize, const std::string& newName) {
    if (pointer) throw OpenMMException("CudaArray has already been initialized");
    cu = &context; size = newSize; elementSize = newElementSize; name = newName;
    void* p = 0;
    checkRt(cudaMalloc(&p, std::max<size_t>(1, size * elementSize)), ("allocating " + name).c_str());
    checkRt(cudaMemset(p, 0, std::max<size_t>(1, size * elementSize)), "cudaMemset");
    pointer = (CUdeviceptr) p;
}
void CudaArray::upload(const void* data, bool) {
    if (!pointer) throw OpenMMException("CudaArray has not been initialized");
    checkRt(cudaMemcpy((void*) pointer, data, size * elementSize, cudaMemcpyHostToDevice), ("uploading " + name).c_str());
}
void CudaArray::download(void* data, bool) const {
    if (!pointer) throw OpenMMException("CudaArray has not been initialized");
    checkRt(cudaMemcpy(data, (const void*) pointer, size * elementSize, cudaMemcpyDeviceToHost), ("downloading " + name).c_str());
}
void CudaArray::uploadConverted(const void* data, size_t hostElementSize, size_t count) {
    if (count != size) throw OpenMMException("Error uploading array " + name + ": The specified vector does not match the size of the array");
    if (hostElementSize == 8 && elementSize == 4) {
        std::vector<float> tmp(count);
        for (size_t i = 0; i < count; i++) tmp[i] = (float) ((const double*) data)[i];
        upload(tmp.data(), true);
    } else if (hostElementSize == 4 && elementSize == 8) {
        std::vector<double> tmp(count);
        for (size_t i = 0; i < count; i++) tmp[i] = (double) ((const float*) data)[i];
        upload(tmp.data(), true);
    } else
        throw OpenMMException("Error uploading array " + name + ": The specified vector has the wrong element size");
}

/* ---- CudaPlatform ---------------------------------------------------------------------------------------- */

CudaPlatform::CudaPlatform() {
    defaultProperties["Precision"] = "single";     /* OpenMM's default */
    defaultProperties["DeviceIndex"] = "0";
}
void CudaPlatform::contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {
    std::string precision = getPropertyDefaultValue("Precision"), device = getPropertyDefaultValue("DeviceIndex");
    for (const auto& kv : properties) {
        if (kv.first == "Precision" || kv.first == "CudaPrecision") precision = kv.second;
        if (kv.first == "DeviceIndex" || kv.first == "CudaDeviceIndex") device = kv.second;
    }
    PlatformData* data = new PlatformData(&context, context.getSystem(), precision, std::atoi(device.c_str()));
    context.setPlatformData(data);
    context.shimContext = data->contexts[0];
}
void CudaPlatform::contextDestroyed(ContextImpl& context) const {
    delete static_cast<PlatformData*>(context.getPlatformData());
    context.setPlatformData(0);
    context.shimContext = 0;
}
CudaPlatform::PlatformData::PlatformData(ContextImpl* context, const System& system, const std::string& precision, int deviceIndex) : context(context) {
    contexts.push_back(new CudaContext(system, deviceIndex, precision, *this));
}
CudaPlatform::PlatformData::~PlatformData() { for (CudaContext* c : contexts) delete c; }

/* ---- CudaContext ----------------------------------------------------------------------------------------- */

CudaContext::CudaContext(const System& system, int deviceIndex, const std::string& precision, CudaPlatform::PlatformData& platformData)
    : shimKernelLaunches(0), system(system), platformData(platformData), deviceIndex(deviceIndex), stream(0), integration(0), time(0), stepCount(0) {
    if (precision == "single") { useDouble = false; useMixed = false; }
    else if (precision == "mixed") { useDouble = false; useMixed = true; }
    else if (precision == "double") { useDouble = true; useMixed = false; }
    else throw OpenMMException("Illegal value for Precision: " + precision);
    numAtoms = system.getNumParticles();
    paddedNumAtoms = (numAtoms + 31) / 32 * 32;      /* TileSize = 32 */
    setAsCurrent();
    checkRt(cudaFree(0), "creating the CUDA context");
    checkRt(cudaDeviceGetAttribute(&multiprocessors, cudaDevAttrMultiProcessorCount, deviceIndex), "cudaDeviceGetAttribute");
    int major = 0, minor = 0;
    checkRt(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, deviceIndex), "cudaDeviceGetAttribute");
    checkRt(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, deviceIndex), "cudaDeviceGetAttribute");
    computeCapability = 10 * major + minor;
    const int real4 = useDouble ? 32 : 16, mixed4 = (useDouble || useMixed) ? 32 : 16;
    posq.initialize(*this, paddedNumAtoms, real4, "posq");
    if (useMixed) posqCorrection.initialize(*this, paddedNumAtoms, real4, "posqCorrection");
    velm.initialize(*this, paddedNumAtoms, mixed4, "velm");
    force.initialize(*this, 3 * (size_t) paddedNumAtoms, sizeof(long long), "force");
    atomIndex.resize(numAtoms);
    indexOfAtom.resize(numAtoms);
    for (int i = 0; i < numAtoms; i++) atomIndex[i] = indexOfAtom[i] = i;
    /* velm.w = 1/mass, 0 for massless particles and padding */
    std::vector<Vec3> zero(numAtoms);
    setVelocities(zero);
    reorderCalls = 0;
    reorderEvery = 0;
    if (const char* r = std::getenv("OPENMM_SHIM_REORDER_EVERY")) reorderEvery = std::atoi(r);
    system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    integration = new CudaIntegrationUtilities(*this, system);
}
CudaContext::~CudaContext() {
    setAsCurrent();
    delete integration;
    for (CUmodule m : modules) cuModuleUnload(m);
}
void CudaContext::setAsCurrent() { checkRt(cudaSetDevice(deviceIndex), "cudaSetDevice"); }

std::string CudaContext::getErrorString(CUresult result) {
    const char* s = 0;
    if (cuGetErrorName(result, &s) == CUDA_SUCCESS && s) return s;
    return "CUDA error";
}

std::vector<double> CudaContext::shimInverseMasses() const {
    std::vector<double> w(numAtoms);
    for (int i = 0; i < numAtoms; i++) { const double m = system.getParticleMass(i); w[i] = m == 0 ? 0.0 : 1.0 / m; }
    return w;
}

/* The block OpenMM's CudaContext::createModule() puts in front of every kernel source (restated, like
   oracle/ref_prelude.h): real / mixed types and constructors, SQRT & co, USE_*_PRECISION.  OpenMM also passes
   --use_fast_math; the shim does so only when OPENMM_SHIM_FAST_MATH=1, so that by default the reference's
   kernels produce the IEEE results the bit-exact comparisons are made against. */
CUmodule CudaContext::createModule(const std::string source, const std::map<std::string, std::string>& defines, const char*) {
    setAsCurrent();
    std::stringstream src;
    for (const auto& kv : defines) src << "#define " << kv.first << " " << kv.second << "\n";
    if (useDouble) src << "#define USE_DOUBLE_PRECISION 1\n";
    if (useMixed) src << "#define USE_MIXED_PRECISION 1\n";
    const char* real = useDouble ? "double" : "float";
    const char* mixed = (useDouble || useMixed) ? "double" : "float";
    for (int n = 1; n <= 4; n++) {
        const std::string suffix = n == 1 ? "" : std::to_string(n);
        src << "typedef " << real << suffix << " real" << suffix << ";\n";
        src << "typedef " << mixed << suffix << " mixed" << suffix << ";\n";
        if (n > 1) {
            src << "#define make_real" << n << " make_" << real << n << "\n";
            src << "#define make_mixed" << n << " make_" << mixed << n << "\n";
        }
    }
    if (useDouble) src << "#define SQRT sqrt\n#define RSQRT rsqrt\n#define RECIP(x) (1.0/(x))\n#define EXP exp\n#define LOG log\n#define POW pow\n#define COS cos\n#define SIN sin\n#define TAN tan\n#define ACOS acos\n#define ASIN asin\n#define ATAN atan\n#define ERF erf\n#define ERFC erfc\n";
    else src << "#define SQRT sqrtf\n#define RSQRT rsqrtf\n#define RECIP(x) (1.0f/(x))\n#define EXP expf\n#define LOG logf\n#define POW powf\n#define COS cosf\n#define SIN sinf\n#define TAN tanf\n#define ACOS acosf\n#define ASIN asinf\n#define ATAN atanf\n#define ERF erff\n#define ERFC erfcf\n";
    src << "#define PADDED_NUM_ATOMS " << paddedNumAtoms << "\n#define NUM_ATOMS " << numAtoms << "\n";
    src << source;
    const std::string text = src.str();

    nvrtcProgram program;
    if (nvrtcCreateProgram(&program, text.c_str(), "seekr2_kernels.cu", 0, NULL, NULL) != NVRTC_SUCCESS)
        throw OpenMMException("openmm_shim: nvrtcCreateProgram failed");
    const std::string arch = "--gpu-architecture=compute_" + std::to_string(computeCapability);
    std::vector<const char*> options = {arch.c_str()};
    const char* fast = std::getenv("OPENMM_SHIM_FAST_MATH");
    if (fast && fast[0] == '1') options.push_back("--use_fast_math");
    const nvrtcResult result = nvrtcCompileProgram(program, (int) options.size(), options.data());
    if (result != NVRTC_SUCCESS) {
        size_t logSize = 0;
        nvrtcGetProgramLogSize(program, &logSize);
        std::string log(logSize, '\0');
        nvrtcGetProgramLog(program, &log[0]);
        nvrtcDestroyProgram(&program);
        throw OpenMMException("Error compiling kernel: " + log);
    }
    size_t ptxSize = 0;
    nvrtcGetPTXSize(program, &ptxSize);
    std::string ptx(ptxSize, '\0');
    nvrtcGetPTX(program, &ptx[0]);
    nvrtcDestroyProgram(&program);
    CUmodule module;
    checkDrv(cuModuleLoadDataEx(&module, ptx.c_str(), 0, NULL, NULL), "Error loading CUDA module");
    modules.push_back(module);
    return module;
}

CUfunction CudaContext::getKernel(CUmodule& module, const std::string& name) {
    CUfunction function;
    CUresult r = cuModuleGetFunction(&function, module, name.c_str());
    if (r != CUDA_SUCCESS) throw OpenMMException("Error creating kernel " + name + ": " + getErrorString(r));
    return function;
}

void CudaContext::executeKernel(CUfunction kernel, void** arguments, int threads, int blockSize, unsigned int sharedSize) {
    if (blockSize == -1) blockSize = 64;                         /* ThreadBlockSize */
    const int numThreadBlocks = 4 * multiprocessors;             /* the kernels use grid-stride loops */
    const int gridSize = std::max(1, std::min((threads + blockSize - 1) / blockSize, numThreadBlocks));
    CUresult r = cuLaunchKernel(kernel, gridSize, 1, 1, blockSize, 1, 1, sharedSize, stream, arguments, NULL);
    if (r != CUDA_SUCCESS) throw OpenMMException("Error invoking kernel: " + getErrorString(r) + " (" + std::to_string((int) r) + ")");
    shimKernelLaunches++;
}

void CudaContext::getForcePositions(std::vector<Vec3>& positions) {
    positions.resize(numAtoms);
    if (useDouble) {
        std::vector<double> p(4 * (size_t) paddedNumAtoms);
        posq.download(p.data());
        for (int i = 0; i < numAtoms; i++) positions[i] = Vec3(p[4 * at(i)], p[4 * at(i) + 1], p[4 * at(i) + 2]);
    } else {
        std::vector<float> p(4 * (size_t) paddedNumAtoms);
        posq.download(p.data());
        for (int i = 0; i < numAtoms; i++) positions[i] = Vec3(p[4 * at(i)], p[4 * at(i) + 1], p[4 * at(i) + 2]);
    }
}
bool CudaContext::getForcePositionsOf(const std::vector<int>& particles, std::vector<Vec3>& positions) {
    const int n = (int) particles.size();
    if (n == 0) return true;
    setAsCurrent();
    if (!gatherKernel) {
        CUmodule module = createModule("extern \"C\" __global__ void shimGather(const real4* __restrict__ posq, const int* __restrict__ index, int n, real4* __restrict__ out) {\n"
                                       "    const int i = blockIdx.x * blockDim.x + threadIdx.x;\n    if (i < n) out[i] = posq[index[i]];\n}\n");
        gatherKernel = getKernel(module, "shimGather");
    }
    if (particles != gatherParticles || gatherReorders != shimReorders) {          /* (re)build the index list */
        if (!gatherIndex.isInitialized() || gatherIndex.getSize() < (size_t) n) {
            if (gatherIndex.isInitialized()) return false;                         /* one subset size per context is enough here */
            gatherIndex.initialize(*this, n, sizeof(int), "shimGatherIndex");
            gatherOut.initialize(*this, n, useDouble ? 32 : 16, "shimGatherOut");
        }
        if (gatherIndex.getSize() != (size_t) n) return false;
        std::vector<int> index(n);
        for (int i = 0; i < n; i++) index[i] = at(particles[i]);
        gatherIndex.upload(index.data());
        gatherParticles = particles;
        gatherReorders = shimReorders;
    }
    int count = n;
    void* args[] = {&posq.getDevicePointer(), &gatherIndex.getDevicePointer(), &count, &gatherOut.getDevicePointer()};
    const CUresult r = cuLaunchKernel(gatherKernel, (n + 63) / 64, 1, 1, 64, 1, 1, 0, stream, args, NULL);
    if (r != CUDA_SUCCESS) throw OpenMMException("Error invoking shimGather: " + getErrorString(r));
    if (useDouble) {
        std::vector<double> p(4 * (size_t) n);
        gatherOut.download(p.data());
        for (int i = 0; i < n; i++) positions[particles[i]] = Vec3(p[4 * i], p[4 * i + 1], p[4 * i + 2]);
    } else {
        std::vector<float> p(4 * (size_t) n);
        gatherOut.download(p.data());
        for (int i = 0; i < n; i++) positions[particles[i]] = Vec3(p[4 * i], p[4 * i + 1], p[4 * i + 2]);
    }
    return true;
}
void CudaContext::getPositions(std::vector<Vec3>& positions) {
    getForcePositions(positions);
    if (useMixed) {
        std::vector<float> c(4 * (size_t) paddedNumAtoms);
        posqCorrection.download(c.data());
        for (int i = 0; i < numAtoms; i++) positions[i] += Vec3(c[4 * at(i)], c[4 * at(i) + 1], c[4 * at(i) + 2]);
    }
}
void CudaContext::setPositions(const std::vector<Vec3>& positions) {
    if (useDouble) {
        std::vector<double> p(4 * (size_t) paddedNumAtoms, 0.0);
        posq.download(p.data());
        for (int i = 0; i < numAtoms; i++) for (int k = 0; k < 3; k++) p[4 * at(i) + k] = positions[i][k];
        posq.upload(p.data());
    } else {
        std::vector<float> p(4 * (size_t) paddedNumAtoms, 0.0f), c(4 * (size_t) paddedNumAtoms, 0.0f);
        posq.download(p.data());
        for (int i = 0; i < numAtoms; i++)
            for (int k = 0; k < 3; k++) {
                p[4 * at(i) + k] = (float) positions[i][k];
                c[4 * at(i) + k] = (float) (positions[i][k] - (double) p[4 * at(i) + k]);
            }
        posq.upload(p.data());
        if (useMixed) posqCorrection.upload(c.data());
    }
}
void CudaContext::getVelocities(std::vector<Vec3>& velocities) {
    velocities.resize(numAtoms);
    if (useDouble || useMixed) {
        std::vector<double> v(4 * (size_t) paddedNumAtoms);
        velm.download(v.data());
        for (int i = 0; i < numAtoms; i++) velocities[i] = Vec3(v[4 * at(i)], v[4 * at(i) + 1], v[4 * at(i) + 2]);
    } else {
        std::vector<float> v(4 * (size_t) paddedNumAtoms);
        velm.download(v.data());
        for (int i = 0; i < numAtoms; i++) velocities[i] = Vec3(v[4 * at(i)], v[4 * at(i) + 1], v[4 * at(i) + 2]);
    }
}
void CudaContext::setVelocities(const std::vector<Vec3>& velocities) {
    const std::vector<double> w = shimInverseMasses();
    if (useDouble || useMixed) {
        std::vector<double> v(4 * (size_t) paddedNumAtoms, 0.0);
        for (int i = 0; i < numAtoms; i++) { for (int k = 0; k < 3; k++) v[4 * at(i) + k] = velocities[i][k]; v[4 * at(i) + 3] = w[i]; }
        velm.upload(v.data());
    } else {
        std::vector<float> v(4 * (size_t) paddedNumAtoms, 0.0f);
        for (int i = 0; i < numAtoms; i++) { for (int k = 0; k < 3; k++) v[4 * at(i) + k] = (float) velocities[i][k]; v[4 * at(i) + 3] = (float) w[i]; }
        velm.upload(v.data());
    }
}
void CudaContext::setForces(const std::vector<Vec3>& forces) {
    std::vector<long long> f(3 * (size_t) paddedNumAtoms, 0);
    for (int i = 0; i < numAtoms; i++)
        for (int k = 0; k < 3; k++) f[at(i) + (size_t) k * paddedNumAtoms] = (long long) std::llrint(forces[i][k] * 4294967296.0);
    force.upload(f.data());
}
void CudaContext::getForces(std::vector<Vec3>& forces) {
    std::vector<long long> f(3 * (size_t) paddedNumAtoms);
    force.download(f.data());
    forces.resize(numAtoms);
    for (int i = 0; i < numAtoms; i++)
        forces[i] = Vec3(f[at(i)] / 4294967296.0, f[at(i) + (size_t) paddedNumAtoms] / 4294967296.0, f[at(i) + 2 * (size_t) paddedNumAtoms] / 4294967296.0);
}
/* OpenMM's CudaContext::reorderAtoms() sorts atoms along a space-filling curve every few hundred steps when a cutoff is in
   use; integrator kernels must therefore never assume particle i sits at buffer index i.  The stand-in has no neighbour
   list, so it offers the same hazard on request: OPENMM_SHIM_REORDER_EVERY=n applies a fixed permutation (rotate by 7
   within blocks of 16, then reverse) to posq / posqCorrection / velm and to atomIndex on every n-th call. */
void CudaContext::reorderAtoms() {
    if (reorderEvery <= 0 || ++reorderCalls % reorderEvery != 0 || numAtoms < 2) return;
    std::vector<int> perm(numAtoms);                            /* new index -> old index */
    for (int i = 0; i < numAtoms; i++) {
        const int block = i / 16 * 16, width = std::min(16, numAtoms - block);
        perm[numAtoms - 1 - i] = block + (i - block + 7) % width;
    }
    auto permute = [&](CudaArray& a) {
        const size_t elem = a.getElementSize();
        std::vector<char> in(a.getSize() * elem), out(a.getSize() * elem);
        a.download(in.data());
        out = in;
        for (int i = 0; i < numAtoms; i++) std::memcpy(&out[i * elem], &in[perm[i] * elem], elem);
        a.upload(out.data());
    };
    permute(posq);
    if (useMixed) permute(posqCorrection);
    permute(velm);
    std::vector<int> newIndex(numAtoms);
    for (int i = 0; i < numAtoms; i++) newIndex[i] = atomIndex[perm[i]];
    atomIndex = newIndex;
    for (int i = 0; i < numAtoms; i++) indexOfAtom[atomIndex[i]] = i;
    shimReorders++;
}

void CudaContext::applyConstraints(double tol) {
    /* Context::applyConstraints: move the positions themselves onto the constraints */
    if (system.getNumConstraints() == 0) return;
    std::vector<Vec3> x, delta(numAtoms);
    getPositions(x);
    integration->shimShake(x, delta, tol);
    for (int i = 0; i < numAtoms; i++) x[i] += delta[i];
    setPositions(x);
}
void CudaContext::applyVelocityConstraints(double tol) {
    if (system.getNumConstraints() == 0) return;
    std::vector<Vec3> x, v;
    getPositions(x);
    getVelocities(v);
    integration->shimRattle(x, v, tol);
    setVelocities(v);
}

/* ---- CudaIntegrationUtilities ---------------------------------------------------------------------------- */

CudaIntegrationUtilities::CudaIntegrationUtilities(CudaContext& context, const System& system)
    : shimPoolFills(0), context(context), system(system), randomPos(0), seed(0), seeded(false), lastStepSize(-1) {
    const bool wide = context.getUseDoublePrecision() || context.getUseMixedPrecision();
    posDelta.initialize(context, context.getPaddedNumAtoms(), wide ? 32 : 16, "posDelta");
    stepSize.initialize(context, 1, wide ? 16 : 8, "stepSize");
}

void CudaIntegrationUtilities::setNextStepSize(double size) {
    if (size == lastStepSize) return;
    lastStepSize = size;
    if (stepSize.getElementSize() == 16) { const double v[2] = {size, size}; stepSize.upload(v); }
    else { const float v[2] = {(float) size, (float) size}; stepSize.upload(v); }
}

void CudaIntegrationUtilities::initRandomNumberGenerator(unsigned int randomNumberSeed) {
    if (seeded) {
        if (randomNumberSeed != seed) throw OpenMMException("CudaIntegrationUtilities::initRandomNumberGenerator(): Requested two different values for the random number seed");
        return;
    }
    seeded = true;
    seed = randomNumberSeed;
    random.initialize(context, 32 * (size_t) context.getPaddedNumAtoms(), 16, "random");
    randomPos = (int) random.getSize();          /* the first prepareRandomNumbers() fills the pool */
}

int CudaIntegrationUtilities::prepareRandomNumbers(int numValues) {
    if (randomPos + numValues <= (int) random.getSize()) {
        const int oldPos = randomPos;
        randomPos += numValues;
        return oldPos;
    }
    if (numValues > (int) random.getSize()) throw OpenMMException("openmm_shim: random pool too small");
    /* timing runs (bench.py's adapter leg): OPENMM_SHIM_REUSE_POOL=1 fills the pool once and then wraps around -- OpenMM
       refills it with a GPU kernel, this stand-in with a host loop over 32 x paddedNumAtoms normals that would swamp the
       integrator kernels being timed */
    static const bool reuse = [] { const char* r = std::getenv("OPENMM_SHIM_REUSE_POOL"); return r && r[0] == '1'; }();
    if (reuse && shimPoolFills > 0) { randomPos = numValues; return 0; }
    std::vector<float> pool(4 * random.getSize());
    if (randomSource()) {
        randomSource()(pool.data(), random.getSize(), shimPoolFills);
    } else {
        /* (seed, fill index) -> normals; OpenMM draws seed 0 from the OS, the shim keeps it deterministic */
        std::mt19937_64 engine(((unsigned long long) seed << 32) ^ (unsigned long long) shimPoolFills ^ 0x5EEC2ull);
        std::normal_distribution<float> normal(0.0f, 1.0f);
        for (float& v : pool) v = normal(engine);
    }
    random.upload(pool.data());
    shimPoolFills++;
    randomPos = numValues;
    return 0;
}

/* SHAKE on the displacements: find delta' so that |x_i + delta'_i - x_j - delta'_j| = d for every constraint */
void CudaIntegrationUtilities::shimShake(const std::vector<Vec3>& x, std::vector<Vec3>& delta, double tol) const {
    const std::vector<double> w = context.shimInverseMasses();
    const int numConstraints = system.getNumConstraints();
    for (int iteration = 0; iteration < 500; iteration++) {
        bool converged = true;
        for (int c = 0; c < numConstraints; c++) {
            int i, j; double d;
            system.getConstraintParameters(c, i, j, d);
            if (w[i] + w[j] == 0) continue;
            const Vec3 r = x[i] - x[j];
            const Vec3 rp = r + delta[i] - delta[j];
            const double diff = d * d - rp.dot(rp);
            if (std::fabs(diff) < 2 * tol * d * d * 0.1) continue;       /* ten times tighter than asked */
            converged = false;
            const double g = diff / (2 * rp.dot(r) * (w[i] + w[j]));
            delta[i] += r * (g * w[i]);
            delta[j] -= r * (g * w[j]);
        }
        if (converged) return;
    }
    throw OpenMMException("openmm_shim: SHAKE did not converge");
}

/* RATTLE velocity stage: remove the velocity component along every constraint */
void CudaIntegrationUtilities::shimRattle(const std::vector<Vec3>& x, std::vector<Vec3>& v, double tol) const {
    const std::vector<double> w = context.shimInverseMasses();
    const int numConstraints = system.getNumConstraints();
    for (int iteration = 0; iteration < 500; iteration++) {
        bool converged = true;
        for (int c = 0; c < numConstraints; c++) {
            int i, j; double d;
            system.getConstraintParameters(c, i, j, d);
            if (w[i] + w[j] == 0) continue;
            const Vec3 r = x[i] - x[j];
            const double rv = r.dot(v[i] - v[j]), r2 = r.dot(r);
            if (std::fabs(rv) < tol * r2 * 0.1) continue;
            converged = false;
            const double g = -rv / (r2 * (w[i] + w[j]));
            v[i] += r * (g * w[i]);
            v[j] -= r * (g * w[j]);
        }
        if (converged) return;
    }
    throw OpenMMException("openmm_shim: RATTLE did not converge");
}

namespace {
void downloadDelta(CudaContext& cu, CudaArray& posDelta, int n, std::vector<Vec3>& delta, std::vector<double>& raw64, std::vector<float>& raw32) {
    delta.resize(n);
    if (posDelta.getElementSize() == 32) {
        raw64.resize(4 * posDelta.getSize());
        posDelta.download(raw64.data());
        for (int i = 0; i < n; i++) { const int a = cu.shimIndexOfAtom(i); delta[i] = Vec3(raw64[4 * a], raw64[4 * a + 1], raw64[4 * a + 2]); }
    } else {
        raw32.resize(4 * posDelta.getSize());
        posDelta.download(raw32.data());
        for (int i = 0; i < n; i++) { const int a = cu.shimIndexOfAtom(i); delta[i] = Vec3(raw32[4 * a], raw32[4 * a + 1], raw32[4 * a + 2]); }
    }
}
}  // namespace

/* between Part1 and Part2: constrain x + posDelta (CudaSeekr2Kernels.cpp:227) */
void CudaIntegrationUtilities::applyConstraints(double tol) {
    if (system.getNumConstraints() == 0) return;
    const int n = context.getNumAtoms();
    std::vector<Vec3> x, delta;
    std::vector<double> raw64; std::vector<float> raw32;
    context.getPositions(x);
    downloadDelta(context, posDelta, n, delta, raw64, raw32);
    shimShake(x, delta, tol);
    if (posDelta.getElementSize() == 32) {
        for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) raw64[4 * context.shimIndexOfAtom(i) + k] = delta[i][k];
        posDelta.upload(raw64.data());
    } else {
        for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) raw32[4 * context.shimIndexOfAtom(i) + k] = (float) delta[i][k];
        posDelta.upload(raw32.data());
    }
}

void CudaIntegrationUtilities::applyVelocityConstraints(double tol) { context.applyVelocityConstraints(tol); }

/* kinetic energy of the velocities shifted by timeShift along the current forces (leapfrog: half a step) */
double CudaIntegrationUtilities::computeKineticEnergy(double timeShift) {
    const int n = context.getNumAtoms();
    std::vector<Vec3> v, f, x;
    context.getVelocities(v);
    const std::vector<double> w = context.shimInverseMasses();
    if (timeShift != 0) {
        context.getForces(f);
        for (int i = 0; i < n; i++) v[i] += f[i] * (timeShift * w[i]);
        if (system.getNumConstraints() > 0) { context.getPositions(x); shimRattle(x, v, 1e-4); }
    }
    double energy = 0;
    for (int i = 0; i < n; i++) if (w[i] != 0) energy += 0.5 * v[i].dot(v[i]) / w[i];
    return energy;
}

}  // namespace OpenMM
