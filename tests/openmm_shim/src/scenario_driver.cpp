/* tests/openmm_shim -- TEST INFRASTRUCTURE ONLY.
 *
 * Runs one scenario through the plugin's public C++ API exactly as a user program would: OpenMM::System +
 * CustomCentroidBondForce boundary definitions, one of the reference's four integrator classes (the reference's
 * own, unmodified sources), Context on the "CUDA" platform, integrator.step(n).  Which kernels do the work is
 * decided at link time: libref_cuda_platform.so (the reference's CUDA platform) or libb200_platform.so (this
 * repository's adapter).  tests/test_gpu_openmm_shim.py runs both binaries on the same scenario file with the same
 * injected random numbers and compares crossing files, statistics files, saved states and the final device
 * buffers bit for bit.
 *
 * usage: scenario <scenario file>
 * scenario file (one directive per line, written by the test):
 *   precision single|mixed|double
 *   integrator mmvt|mmvt_middle|elber|elber_middle <temperature> <friction> <stepSize>
 *   output <path> | savestate <prefix> | savestats <path> | counter <n> | endonsrc 0|1 | seed <n>
 *   milestone <forceGroup> | src <forceGroup> | dest <forceGroup>
 *   particle <mass> <x> <y> <z> <vx> <vy> <vz>
 *   constraint <i> <j> <distance>
 *   harmonic <i> <j> <length> <k> [forceGroup]
 *   tether <k>                                  F = -k (x - x0), x0 = the initial positions
 *   centroid <groupsPerBond> <forceGroup> <expression>       followed by
 *     param <name> | group <n> <i...> | wgroup <n> <i...> <w...> | bond <g...> : <p...> | end
 *   pool <path> <count>                         float32[count][4] normals, consumed 32*Np per pool refill
 *   steps <n> [<n> ...]                         several integrator.step() calls
 *   settime <t>                                 context.setTime(t) before the next steps directive
 *   settle                                      ContextImpl::calcKineticEnergy() -- what getState(getEnergy=True) reaches: the
 *                                               point at which an asynchronous platform must have its files and clock up to date
 *   dump <path>                                 final posq | posqCorrection | velm bytes + "time stepCount"
 */
#include "ElberLangevinIntegrator.h"
#include "ElberLangevinMiddleIntegrator.h"
#include "MmvtLangevinIntegrator.h"
#include "MmvtLangevinMiddleIntegrator.h"
#include "openmm/Context.h"
#include "openmm/CustomCentroidBondForce.h"
#include "openmm/HarmonicBondForce.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/cuda/CudaContext.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/internal/ContextImpl.h"
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>

using namespace OpenMM;
using namespace Seekr2Plugin;

extern "C" void registerSeekr2CudaKernelFactories();

namespace {
/* the synthetic force model of the benchmarks; reads what force kernels read: posq */
class TetherForce : public Force {
public:
    TetherForce(double k, const std::vector<Vec3>& x0) : k(k), x0(x0) {}
    double shimEvaluate(const std::vector<Vec3>& x, const std::vector<double>&, std::vector<Vec3>* f) const {
        double e = 0;
        for (size_t i = 0; i < x.size(); i++) {
            const Vec3 d = x[i] - x0[i];
            e += 0.5 * k * d.dot(d);
            if (f) (*f)[i] -= d * k;
        }
        return e;
    }
private:
    double k;
    std::vector<Vec3> x0;
};

struct Action { enum Kind { Steps, SetTime, Settle } kind; int steps; double time; };
}  // namespace

int main(int argc, char* argv[]) {
    if (argc != 2) { std::fprintf(stderr, "usage: %s <scenario file>\n", argv[0]); return 2; }
    try {
        std::ifstream in(argv[1]);
        if (!in) throw OpenMMException(std::string("cannot open ") + argv[1]);
        std::string precision = "mixed", kind, output, saveState, saveStats, dumpPath, poolPath;
        double temperature = 300, friction = 1, stepSize = 0.002, tetherK = 0;
        int counter = 0, endOnSrc = 0, seed = 0;
        long long poolCount = 0;
        std::vector<int> milestones, srcs, dests;
        std::vector<Vec3> positions, velocities;
        std::vector<Action> actions;
        System system;
        CustomCentroidBondForce* centroid = 0;
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            std::string word;
            if (!(ls >> word) || word[0] == '#') continue;
            if (word == "precision") ls >> precision;
            else if (word == "integrator") ls >> kind >> temperature >> friction >> stepSize;
            else if (word == "output") ls >> output;
            else if (word == "savestate") ls >> saveState;
            else if (word == "savestats") ls >> saveStats;
            else if (word == "counter") ls >> counter;
            else if (word == "endonsrc") ls >> endOnSrc;
            else if (word == "seed") ls >> seed;
            else if (word == "milestone") { int g; ls >> g; milestones.push_back(g); }
            else if (word == "src") { int g; ls >> g; srcs.push_back(g); }
            else if (word == "dest") { int g; ls >> g; dests.push_back(g); }
            else if (word == "particle") {
                double m, x, y, z, vx, vy, vz;
                ls >> m >> x >> y >> z >> vx >> vy >> vz;
                system.addParticle(m);
                positions.push_back(Vec3(x, y, z));
                velocities.push_back(Vec3(vx, vy, vz));
            }
            else if (word == "constraint") { int i, j; double d; ls >> i >> j >> d; system.addConstraint(i, j, d); }
            else if (word == "harmonic") {
                int i, j, group = 0; double length, k; ls >> i >> j >> length >> k;
                HarmonicBondForce* f = new HarmonicBondForce();
                f->addBond(i, j, length, k);
                if (ls >> group) f->setForceGroup(group);
                system.addForce(f);
            }
            else if (word == "tether") ls >> tetherK;
            else if (word == "centroid") {
                int groupsPerBond, forceGroup;
                ls >> groupsPerBond >> forceGroup;
                std::string expression;
                std::getline(ls, expression);
                centroid = new CustomCentroidBondForce(groupsPerBond, expression);
                centroid->setForceGroup(forceGroup);
                system.addForce(centroid);
            }
            else if (word == "param") { std::string name; ls >> name; centroid->addPerBondParameter(name); }
            else if (word == "group" || word == "wgroup") {
                int n; ls >> n;
                std::vector<int> atoms(n);
                for (int& a : atoms) ls >> a;
                std::vector<double> weights(word == "wgroup" ? n : 0);
                for (double& w : weights) ls >> w;
                centroid->addGroup(atoms, weights);
            }
            else if (word == "bond") {
                std::vector<int> groups; std::vector<double> params;
                std::string tok; bool afterColon = false;
                while (ls >> tok) {
                    if (tok == ":") { afterColon = true; continue; }
                    if (afterColon) params.push_back(std::stod(tok)); else groups.push_back(std::stoi(tok));
                }
                centroid->addBond(groups, params);
            }
            else if (word == "end") centroid = 0;
            else if (word == "pool") ls >> poolPath >> poolCount;
            else if (word == "steps") { int n; while (ls >> n) actions.push_back(Action{Action::Steps, n, 0}); }
            else if (word == "settime") { double t; ls >> t; actions.push_back(Action{Action::SetTime, 0, t}); }
            else if (word == "settle") actions.push_back(Action{Action::Settle, 0, 0});
            else if (word == "dump") ls >> dumpPath;
            else throw OpenMMException("unknown directive " + word);
        }
        if (tetherK > 0) system.addForce(new TetherForce(tetherK, positions));

        std::vector<float> pool;
        if (!poolPath.empty()) {
            pool.resize(4 * (size_t) poolCount);
            std::ifstream pf(poolPath, std::ios::binary);
            pf.read((char*) pool.data(), pool.size() * sizeof(float));
            if (!pf) throw OpenMMException("cannot read " + poolPath);
            shimSetRandomSource([&pool](float* out, size_t count, long long fill) {
                for (size_t i = 0; i < 4 * count; i++) {
                    const size_t src = (size_t) fill * 4 * count + i;
                    out[i] = src < pool.size() ? pool[src] : 0.0f;
                }
            });
        }

        registerSeekr2CudaKernelFactories();
        Platform& platform = Platform::getPlatformByName("CUDA");
        std::unique_ptr<Integrator> integrator;
        if (kind == "mmvt" || kind == "mmvt_middle") {
            auto configure = [&](auto* integ) {
                integ->setRandomNumberSeed(seed);
                integ->setSaveStateFileName(saveState);
                integ->setSaveStatisticsFileName(saveStats);
                integ->setBounceCounter(counter);
                for (int g : milestones) integ->addMilestoneGroup(g);
            };
            if (kind == "mmvt") { auto* p = new MmvtLangevinIntegrator(temperature, friction, stepSize, output); configure(p); integrator.reset(p); }
            else { auto* p = new MmvtLangevinMiddleIntegrator(temperature, friction, stepSize, output); configure(p); integrator.reset(p); }
        } else if (kind == "elber" || kind == "elber_middle") {
            auto configure = [&](auto* integ) {
                integ->setRandomNumberSeed(seed);
                integ->setSaveStateFileName(saveState);
                integ->setCrossingCounter(counter);
                integ->setEndOnSrcMilestone(endOnSrc != 0);
                for (int g : srcs) integ->addSrcMilestoneGroup(g);
                for (int g : dests) integ->addDestMilestoneGroup(g);
            };
            if (kind == "elber") { auto* p = new ElberLangevinIntegrator(temperature, friction, stepSize, output); configure(p); integrator.reset(p); }
            else { auto* p = new ElberLangevinMiddleIntegrator(temperature, friction, stepSize, output); configure(p); integrator.reset(p); }
        } else
            throw OpenMMException("unknown integrator " + kind);

        std::map<std::string, std::string> properties;
        properties["CudaPrecision"] = precision;
        int status = 0;
        {
            Context context(system, *integrator, platform, properties);
            context.setPositions(positions);
            context.setVelocities(velocities);
            std::vector<double> stepSeconds;                      /* wall time of every steps directive, device drained */
            try {
                for (const Action& a : actions) {
                    if (a.kind == Action::SetTime) context.setTime(a.time);
                    else if (a.kind == Action::Settle) context.getImpl().calcKineticEnergy();
                    else {
                        cudaDeviceSynchronize();
                        const auto t0 = std::chrono::steady_clock::now();
                        integrator->step(a.steps);
                        cudaDeviceSynchronize();
                        stepSeconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
                    }
                }
            } catch (const std::exception& e) {
                std::cout << "step exception: " << e.what() << std::endl;
                status = 3;
            }
            ContextImpl& impl = context.getImpl();
            CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(impl.getPlatformData())->contexts[0];
            if (!dumpPath.empty()) {
                std::ofstream dump(dumpPath, std::ios::binary);
                for (CudaArray* a : {&cu.getPosq(), cu.getUseMixedPrecision() ? &cu.getPosqCorrection() : (CudaArray*) 0, &cu.getVelm()}) {
                    if (!a) continue;
                    std::vector<char> bytes(a->getSize() * a->getElementSize());
                    a->download(bytes.data());
                    dump.write(bytes.data(), bytes.size());
                }
                const double t = cu.getTime();
                const long long sc = cu.getStepCount();
                dump.write((const char*) &t, sizeof t);
                dump.write((const char*) &sc, sizeof sc);
            }
            std::cout << "time " << cu.getTime() << " stepCount " << cu.getStepCount()
                      << " forceEvaluations " << impl.shimForceEvaluations << " energyEvaluations " << impl.shimEnergyEvaluations
                      << " shimKernelLaunches " << cu.shimKernelLaunches << " shimReorders " << cu.shimReorders
                      << " forceNs " << impl.shimForceNs << " energyNs " << impl.shimEnergyNs << std::endl;
            std::cout << "stepSeconds";
            for (double t : stepSeconds) std::cout << " " << t;
            std::cout << std::endl;
        }
        std::cout << (status == 0 ? "Done" : "Failed") << std::endl;
        return status;
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        return 1;
    }
}
