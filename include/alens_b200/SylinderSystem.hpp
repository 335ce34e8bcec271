// SylinderSystem.hpp -- the hot members of SimToolbox/Sylinder/SylinderSystem.{hpp,cpp} with the
// reference's names and call sequence, implemented over the C ABI (include/alens_b200.h):
//   prepareStep :886-932, calcMobOperator :719-722, calcVelocityNonCon :724-800, collectPairCollision
//   :1152-1160, resolveConstraints :829-866, saveForceVelocityConstraints :971-1018, sumForceVelocity
//   :802-814, stepEuler :816-827, runStep :948-969, setForceNonBrown/setVelocityNonBrown :934-946.
// Not here (they stay with the host application, SURVEY.md section 8 "out of scope"): file/VTK I/O, YAML,
// Brownian noise, domain decomposition by FDPS.  Boundaries: collectBoundaryCollision :1093-1150, links:
// collectLinkBilateral :1386-1482 (both on the device).
#ifndef ALENS_B200_SYLINDERSYSTEM_HPP_
#define ALENS_B200_SYLINDERSYSTEM_HPP_

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <map>
#include <random>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ConstraintSolver.hpp"
#include "Sylinder.hpp"
#include "SylinderConfig.hpp"

// the per-rod host loops run on all cores when the host program is compiled with OpenMP, as the reference's do
#ifdef _OPENMP
#define ALENS_OMP_FOR _Pragma("omp parallel for schedule(static)")
#else
#define ALENS_OMP_FOR
#endif

class SylinderSystem {
    alens_ctx *ctx_ = nullptr;
    int stepCount = 0;
    bool brownianOnDevice = false; // runStep() calls calcVelocityBrown() itself when KBT > 0 (else: setVelocityBrown)
    std::vector<Sylinder> sylinderContainer;
    std::shared_ptr<ConstraintSolver> conSolverPtr;
    std::shared_ptr<ConstraintCollector> conCollectorPtr;
    Teuchos::RCP<const TV> forceUniRcp, velocityUniRcp, forceBiRcp, velocityBiRcp;
    Teuchos::RCP<TV> forcePartNonBrownRcp, velocityPartNonBrownRcp, velocityBrownRcp, velocityNonConRcp;
    Teuchos::RCP<const TCOMM> commRcp;
    Teuchos::RCP<TMAP> sylinderMapRcp, sylinderMobilityMapRcp;
    Teuchos::RCP<TOP> mobilityOperatorRcp; ///< placeholder: the mobility lives on the device
    bool multiRank_ = false; ///< slab decomposition: one SylinderSystem (one GPU) per rank

    void ck(int rc) const {
        if (rc != ALENS_OK) throw std::runtime_error(alens_last_error(ctx_));
    }
    void updateSylinderMap() { // :868-880
        const int nLocal = (int)sylinderContainer.size();
        // multi-rank: the exclusive scan of the ranks' rod counts is kept by the device layer (alens_set_decomposition at
        // start, renumbered by every alens_migrate_rods), as getTMAPFromLocalSize's MPI scan does in the reference
        int offset = 0;
        if (multiRank_) ck(alens_get_rod_identity(ctx_, nullptr, &offset, nullptr, nullptr, nullptr, nullptr));
        sylinderMapRcp = getTMAPFromLocalSize(nLocal, commRcp, offset);
        sylinderMobilityMapRcp = getTMAPFromLocalSize(nLocal * 6, commRcp, 6 * offset);
        const int base = sylinderMapRcp->getMinGlobalIndex();
        ALENS_OMP_FOR
        for (int i = 0; i < nLocal; i++) sylinderContainer[i].globalIndex = i + base;
    }

  public:
    SylinderConfig runConfig;

    SylinderSystem() = default;
    SylinderSystem(const SylinderConfig &config, std::vector<Sylinder> rods, int device = 0) {
        initialize(config, std::move(rods), device);
    }
    /// the reference's constructors (SylinderSystem.hpp:172-186, SylinderSystem.cpp:27-33)
    SylinderSystem(const std::string &configFile, const std::string &posFile, int argc, char **argv) {
        initialize(SylinderConfig(configFile), posFile, argc, argv);
    }
    SylinderSystem(const SylinderConfig &config, const std::string &posFile, int argc, char **argv) {
        initialize(config, posFile, argc, argv);
    }
    ~SylinderSystem() {
        if (ctx_) alens_destroy(ctx_);
    }
    SylinderSystem(const SylinderSystem &) = delete;
    SylinderSystem &operator=(const SylinderSystem &) = delete;

    /// SylinderSystem::initialize(config, posFile, argc, argv) (:35-104): rods from `posFile` if it exists
    /// (setInitialFromFile :317-375), else drawn from the configuration (setInitialFromConfig :218-261); links from the
    /// `L prev next` lines of the same file (:377-405).  The CUDA device is the process's local rank
    /// (ALENS_DEVICE, LOCAL_RANK or OMPI_COMM_WORLD_LOCAL_RANK; default 0).
    void initialize(const SylinderConfig &config, const std::string &posFile, int /*argc*/, char ** /*argv*/) {
        runConfig = config;
        std::vector<Sylinder> rods;
        bool haveFile = false;
        {
            std::ifstream probe(posFile.c_str());
            haveFile = !posFile.empty() && probe.good(); // IOHelper::fileExist
        }
        if (haveFile) {
            rods = readSylinderFile(posFile);
            setLinkMapFromFile(posFile);
        } else {
            rods = drawInitialRods(config);
        }
        int device = 0;
        for (const char *name : {"ALENS_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"})
            if (const char *v = std::getenv(name)) {
                device = std::atoi(v);
                break;
            }
        initialize(config, std::move(rods), device);
    }

    /// Equatn::FromTwoVectors((0,0,1), d) (Eigen's setFromTwoVectors; d antiparallel to z: a half turn about x)
    static void orientationFromDirection(const double d_[3], double q[4]) {
        const double n = std::sqrt(d_[0] * d_[0] + d_[1] * d_[1] + d_[2] * d_[2]);
        const double d[3] = {d_[0] / n, d_[1] / n, d_[2] / n};
        const double c = d[2];
        if (c < -1.0 + 1e-12) {
            q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0;
            return;
        }
        const double s = std::sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
        q[0] = -d[1] * invs; q[1] = d[0] * invs; q[2] = 0.0; q[3] = s * 0.5; // axis = z x d
    }
    /// one `C|S gid radius mx my mz px py pz [group]` line per rod (parseSylinder, :320-344)
    static std::vector<Sylinder> readSylinderFile(const std::string &filename) {
        std::ifstream f(filename);
        if (!f) throw std::runtime_error("cannot open " + filename);
        std::string line;
        std::getline(f, line); // two header lines
        std::getline(f, line);
        std::vector<Sylinder> rods;
        while (std::getline(f, line)) {
            if (line.empty() || (line[0] != 'C' && line[0] != 'S')) continue;
            std::stringstream ss(line);
            char type;
            int gid, group = -1;
            double radius = 0, m[3] = {0, 0, 0}, p[3] = {0, 0, 0};
            ss >> type >> gid >> radius >> m[0] >> m[1] >> m[2] >> p[0] >> p[1] >> p[2];
            if (ss.fail()) throw std::runtime_error("readSylinderFile: malformed line: " + line);
            ss >> group;
            Sylinder sy;
            for (int k = 0; k < 3; k++) sy.pos[k] = (m[k] + p[k]) * 0.5;
            sy.gid = gid;
            sy.group = group;
            sy.isImmovable = type == 'S';
            sy.radius = sy.radiusCollision = radius;
            sy.length = std::sqrt(std::pow(p[0] - m[0], 2) + std::pow(p[1] - m[1], 2) + std::pow(p[2] - m[2], 2));
            sy.lengthCollision = sy.length;
            const double dir[3] = {p[0] - m[0], p[1] - m[1], p[2] - m[2]}, ez[3] = {0, 0, 1};
            orientationFromDirection(sy.length > 1e-7 ? dir : ez, sy.orientation);
            sy.clear();
            rods.push_back(sy);
        }
        return rods;
    }
    void setLinkMapFromFile(const std::string &filename) { // :377-405
        std::ifstream f(filename);
        std::string line;
        std::getline(f, line);
        std::getline(f, line);
        linkMap.clear();
        while (std::getline(f, line)) {
            if (line.empty() || line[0] != 'L') continue;
            std::stringstream ss(line);
            char h;
            int prev, next;
            ss >> h >> prev >> next;
            linkMap.emplace(prev, next);
        }
    }
    /// setInitialFromConfig (:218-261): sylinderNumber rods, uniform in the init box, lengths sylinderLength (log-normal
    /// with sylinderLengthSigma > 0, redrawn until shorter than half the smallest box edge), orientation by getOrient
    /// (:190-216: a component of initOrient outside [-1, 1] is random).  The reference draws from per-thread TRNG streams;
    /// here one std::mt19937_64 seeded with rngSeed (documented deviation: TRNG is not a dependency of this library).
    static std::vector<Sylinder> drawInitialRods(const SylinderConfig &cfg) {
        std::mt19937_64 gen(cfg.rngSeed);
        std::uniform_real_distribution<double> u01(0.0, 1.0);
        std::lognormal_distribution<double> ln(cfg.sylinderLength > 0 ? cfg.sylinderLength : 1.0,
                                               cfg.sylinderLengthSigma > 0 ? cfg.sylinderLengthSigma : 1.0);
        double edge[3], minEdge = 1e300;
        for (int k = 0; k < 3; k++) {
            edge[k] = cfg.initBoxHigh[k] - cfg.initBoxLow[k];
            minEdge = std::min(minEdge, edge[k]);
        }
        const double pi = 3.14159265358979323846;
        std::vector<Sylinder> rods((size_t)std::max(cfg.sylinderNumber, 0));
        for (int i = 0; i < (int)rods.size(); i++) {
            double length = cfg.sylinderLength;
            if (cfg.sylinderLengthSigma > 0) do length = ln(gen); while (length >= minEdge * 0.5);
            double pos[3], p[3], q[4];
            for (int k = 0; k < 3; k++) pos[k] = u01(gen) * edge[k] + cfg.initBoxLow[k];
            bool allRandom = true;
            for (int k = 0; k < 3; k++) {
                const double o = cfg.initOrient[k];
                if (o < -1 || o > 1) p[k] = 2 * u01(gen) - 1;
                else { p[k] = o; allRandom = false; }
            }
            if (allRandom) { // EquatnHelper::setUnitRandomEquatn: uniform on SO(3)
                const double u1 = u01(gen), u2 = u01(gen), u3 = u01(gen);
                const double a = std::sqrt(1 - u1), b = std::sqrt(u1);
                q[3] = a * std::sin(2 * pi * u2); q[0] = a * std::cos(2 * pi * u2);
                q[1] = b * std::sin(2 * pi * u3); q[2] = b * std::cos(2 * pi * u3);
            } else {
                orientationFromDirection(p, q);
            }
            rods[i] = Sylinder(i, cfg.sylinderDiameter / 2, cfg.sylinderDiameter / 2, length, length, pos, q);
        }
        return rods;
    }

    /// setInitialFromVTKFile (:406-476): the rods of a Sylinder_<snap>.pvtp written by writeResult (or by the reference):
    /// centre = midpoint of the two end points, orientation = FromTwoVectors(ez, znorm), the Float32 cell arrays gid, group,
    /// isImmovable, length, lengthCollision, radius, radiusCollision, vel, omega
    static std::vector<Sylinder> readSylinderVTK(const std::string &pvtpFileName) {
        const alens_vtk::PolyData pd = alens_vtk::readParallel(pvtpFileName);
        const size_t n = pd.numberOfPoints() / 2; // two points per sylinder
        const auto &gid = pd.cell("gid"), &group = pd.cell("group"), &imm = pd.cell("isImmovable");
        const auto &len = pd.cell("length"), &lenC = pd.cell("lengthCollision"), &rad = pd.cell("radius");
        const auto &radC = pd.cell("radiusCollision"), &znorm = pd.cell("znorm"), &vel = pd.cell("vel"), &omega = pd.cell("omega");
        for (const alens_vtk::Array *a : {&gid, &group, &imm, &len, &lenC, &rad, &radC})
            if (a->v.size() != n) throw std::runtime_error("readSylinderVTK: array length != number of rods in " + pvtpFileName);
        for (const alens_vtk::Array *a : {&znorm, &vel, &omega})
            if (a->v.size() != 3 * n) throw std::runtime_error("readSylinderVTK: vector array length != 3 x rods in " + pvtpFileName);
        std::vector<Sylinder> rods(n);
        ALENS_OMP_FOR
        for (long long i = 0; i < (long long)n; i++) {
            Sylinder &sy = rods[(size_t)i];
            for (int k = 0; k < 3; k++) {
                sy.pos[k] = (pd.points[6 * (size_t)i + k] + pd.points[6 * (size_t)i + 3 + k]) * 0.5;
                sy.vel[k] = vel.v[3 * (size_t)i + k];
                sy.omega[k] = omega.v[3 * (size_t)i + k];
            }
            sy.gid = (int)gid.v[(size_t)i];
            sy.group = (int)group.v[(size_t)i];
            sy.isImmovable = imm.v[(size_t)i] > 0;
            sy.length = len.v[(size_t)i];
            sy.lengthCollision = lenC.v[(size_t)i];
            sy.radius = rad.v[(size_t)i];
            sy.radiusCollision = radC.v[(size_t)i];
            const double d[3] = {znorm.v[3 * (size_t)i], znorm.v[3 * (size_t)i + 1], znorm.v[3 * (size_t)i + 2]};
            orientationFromDirection(d, sy.orientation);
        }
        return rods;
    }

    /// what a restart file (TimeStepInfo.txt, written by writeResult :509-521) holds
    struct RestartInfo {
        unsigned rngSeed = 0;
        int stepCount = 0, snapID = 0;
        std::string pvtpFileName, asciiFileName;
    };
    static RestartInfo readRestartFile(const std::string &restartFile) {
        std::ifstream f(restartFile);
        if (!f) throw std::runtime_error("reinitialize: cannot open " + restartFile);
        RestartInfo r;
        f >> r.rngSeed >> r.stepCount >> r.snapID >> r.pvtpFileName;
        if (f.fail()) throw std::runtime_error("reinitialize: malformed restart file " + restartFile);
        r.asciiFileName = r.pvtpFileName; // Sylinder_<snap>.pvtp -> SylinderAscii_<snap>.dat (:143-147)
        const size_t dot = r.asciiFileName.find_last_of('.'), us = r.asciiFileName.find_last_of('_');
        if (dot == std::string::npos || us == std::string::npos) throw std::runtime_error("reinitialize: bad pvtp name in " + restartFile);
        r.asciiFileName.replace(dot, 5, ".dat");
        r.asciiFileName.replace(us, 1, "Ascii_");
        return r;
    }

    /// SylinderSystem::reinitialize (:106-175): resume from the snapshot a restart file names -- rods from the .pvtp, links
    /// from the .dat next to it (both in getCurrentResultFolder()), one Euler step with the stored velocities (the
    /// snapshot is written before the step that follows it), then stepCount and snapID move on by one.  The reference
    /// re-seeds its generator with restartRngSeed + 1; the device-side Brownian generator is keyed by (rngSeed, stepCount),
    /// so runConfig.rngSeed takes that value.
    void reinitialize(const SylinderConfig &config, const std::string &restartFile, int /*argc*/, char ** /*argv*/,
                      bool eulerStep = true) {
        const RestartInfo info = readRestartFile(restartFile);
        commRcp = getMPIWORLDTCOMM();
        snapID = info.snapID;
        const std::string baseFolder = getCurrentResultFolder();
        std::vector<Sylinder> rods = readSylinderVTK(baseFolder + info.pvtpFileName);
        setLinkMapFromFile(baseFolder + info.asciiFileName);
        if (eulerStep && !config.sylinderFixed) {
            const long long n = (long long)rods.size();
            ALENS_OMP_FOR
            for (long long i = 0; i < n; i++) rods[(size_t)i].stepEuler(config.dt);
        }
        int device = 0;
        for (const char *name : {"ALENS_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"})
            if (const char *v = std::getenv(name)) {
                device = std::atoi(v);
                break;
            }
        SylinderConfig resumed = config;
        resumed.initPreSteps = 0; // no initial collision resolution on a restart (:106-175 has none)
        restartRngSeed = info.rngSeed + 1;
        resumed.rngSeed = restartRngSeed;
        if (ctx_) {
            alens_destroy(ctx_);
            ctx_ = nullptr;
        }
        initialize(resumed, std::move(rods), device);
        runConfig.initPreSteps = config.initPreSteps;
        stepCount = info.stepCount + 1;
        snapID = info.snapID + 1;
    }

    /// rods are handed over by the caller
    void initialize(const SylinderConfig &config, std::vector<Sylinder> rods, int device = 0) {
        runConfig = config;
        stepCount = 0;
        restartRngSeed = runConfig.rngSeed; // :41
        commRcp = getMPIWORLDTCOMM();
        if (alens_create(device, 0, 1, &ctx_) != ALENS_OK) throw std::runtime_error(alens_last_error(nullptr));
        conSolverPtr = std::make_shared<ConstraintSolver>(ctx_);
        conCollectorPtr = std::make_shared<ConstraintCollector>();
        sylinderContainer = std::move(rods);
        const int pbc[3] = {runConfig.simBoxPBC[0], runConfig.simBoxPBC[1], runConfig.simBoxPBC[2]};
        ck(alens_set_domain(ctx_, runConfig.simBoxLow, runConfig.simBoxHigh, pbc)); // setDomainInfo :569-610
        if (!runConfig.sylinderFixed) { // initial collision resolution, :88-101
            for (int i = 0; i < runConfig.initPreSteps; i++) {
                prepareStep();
                calcVelocityNonCon();
                resolveConstraints();
                saveForceVelocityConstraints();
                sumForceVelocity();
                stepEuler();
            }
        }
    }


    // ---- more than one rank: slabs along one box axis, one SylinderSystem (one GPU) per rank.  What FDPS's DomainInfo +
    // exchangeParticle and Tpetra's maps do in the reference (SylinderSystem.cpp:569-620, :868-880) is done by the device
    // layer: ghost rods, halo of the rod velocities, allreduce of the BBPGD scalars, rod migration.
    struct Decomposition {
        int rank = 0, nranks = 1;
        int axis = 2;                        ///< slab axis (z: the slowest axis of the cell order, boundaries are contiguous)
        double skin = 0;                     ///< how far a rod may drift out of its slab before it is migrated (>= one step's motion)
        double globalMaxBoundingRadius = 0;  ///< max over ALL ranks of lengthCollision / 2 + radiusCollision
        int globalIndexBase = 0;             ///< number of rods on the ranks below this one (MPI_Exscan of the local counts)
        long long maxLocalRods = 0;          ///< capacity of the communication window: the SAME number on every rank
    };
    /// rank `dec.rank` of `dec.nranks`: this rank's rods are those of `rods` (they must lie in its slab).  The communicator
    /// is connected afterwards (connectLocal, or exportCommBlob + connectComm around the host's MPI_Allgather); the initial
    /// collision resolution (:88-101) then runs with resolveInitialCollisions().
    void initialize(const SylinderConfig &config, std::vector<Sylinder> rods, int device, const Decomposition &dec) {
        runConfig = config;
        stepCount = 0;
        multiRank_ = dec.nranks > 1;
        commRcp = getMPIWORLDTCOMM(dec.rank, dec.nranks);
        if (alens_create(device, dec.rank, dec.nranks, &ctx_) != ALENS_OK) throw std::runtime_error(alens_last_error(nullptr));
        conSolverPtr = std::make_shared<ConstraintSolver>(ctx_);
        conCollectorPtr = std::make_shared<ConstraintCollector>();
        sylinderContainer = std::move(rods);
        const int pbc[3] = {runConfig.simBoxPBC[0], runConfig.simBoxPBC[1], runConfig.simBoxPBC[2]};
        ck(alens_set_domain(ctx_, runConfig.simBoxLow, runConfig.simBoxHigh, pbc));
        ck(alens_set_collision_params(ctx_, runConfig.sylinderDiameterColRatio, runConfig.sylinderLengthColRatio,
                                      runConfig.sylinderColBuf));
        if (multiRank_) {
            const double lo = runConfig.simBoxLow[dec.axis], w = (runConfig.simBoxHigh[dec.axis] - lo) / dec.nranks;
            ck(alens_set_decomposition(ctx_, dec.axis, lo + dec.rank * w, lo + (dec.rank + 1) * w, dec.skin,
                                       dec.globalMaxBoundingRadius, dec.globalIndexBase));
            if (dec.maxLocalRods <= 0) throw std::invalid_argument("Decomposition::maxLocalRods: the window capacity (same on all ranks)");
            ck(alens_comm_create(ctx_, dec.maxLocalRods));
        }
    }
    std::vector<char> exportCommBlob() { // what this rank publishes to its peers (cudaIpc handle of its window)
        std::vector<char> blob((size_t)alens_comm_blob_size());
        ck(alens_comm_export(ctx_, blob.data()));
        return blob;
    }
    void connectComm(const std::vector<char> &blobsInRankOrder) { ck(alens_comm_connect(ctx_, blobsInRankOrder.data())); }
    /// ranks living in one process (one host thread each)
    static void connectLocal(const std::vector<SylinderSystem *> &ranks) {
        std::vector<alens_ctx *> c;
        for (auto *r : ranks) c.push_back(r->ctx_);
        if (alens_comm_connect_local(c.data(), (int)c.size()) != ALENS_OK) throw std::runtime_error(alens_last_error(c[0]));
    }
    void resolveInitialCollisions() { // :88-101
        if (runConfig.sylinderFixed) return;
        for (int i = 0; i < runConfig.initPreSteps; i++) {
            prepareStep();
            calcVelocityNonCon();
            resolveConstraints();
            saveForceVelocityConstraints();
            sumForceVelocity();
            stepEuler();
        }
    }
    void setBrownianOnDevice(bool on) { brownianOnDevice = on; }
    bool isMultiRank() const { return multiRank_; }

    alens_ctx *deviceContext() { return ctx_; }
    const std::vector<Sylinder> &getContainer() { return sylinderContainer; }
    std::vector<Sylinder> &getContainerNonConst() { return sylinderContainer; }
    Teuchos::RCP<const TCOMM> &getCommRcp() { return commRcp; }
    ConstraintBlockPool &getConstraintPoolNonConst() { return *(conCollectorPtr->constraintPoolPtr); }
    std::shared_ptr<ConstraintSolver> &getConstraintSolver() { return conSolverPtr; }
    std::shared_ptr<ConstraintCollector> &getConstraintCollector() { return conCollectorPtr; }
    int getStepCount() { return stepCount; }

    void prepareStep() { // :886-932
        ck(alens_set_collision_params(ctx_, runConfig.sylinderDiameterColRatio, runConfig.sylinderLengthColRatio,
                                      runConfig.sylinderColBuf));
        const int nLocal = (int)sylinderContainer.size();
        ALENS_OMP_FOR
        for (int i = 0; i < nLocal; i++) {
            auto &sy = sylinderContainer[i];
            sy.clear();
            sy.radiusCollision = sy.radius * runConfig.sylinderDiameterColRatio;
            sy.lengthCollision = sy.length * runConfig.sylinderLengthColRatio;
            sy.rank = commRcp->getRank();
            sy.colBuf = runConfig.sylinderColBuf;
        }
        if (runConfig.monolayer) { // :907-918 (direction flattened into the xy plane)
            const double monoZ = (runConfig.simBoxHigh[2] + runConfig.simBoxLow[2]) / 2;
            ALENS_OMP_FOR
            for (int i = 0; i < nLocal; i++) {
                auto &sy = sylinderContainer[i];
                sy.pos[2] = monoZ;
                const double *q = sy.orientation;
                double dx = 2 * (q[0] * q[2] + q[3] * q[1]), dy = 2 * (q[1] * q[2] - q[3] * q[0]);
                const double n = std::sqrt(dx * dx + dy * dy);
                if (n > 0) { // FromTwoVectors(z, d) with d in the xy plane
                    dx /= n; dy /= n;
                    const double s = std::sqrt(2.0);
                    sy.orientation[0] = -dy / s; sy.orientation[1] = dx / s; sy.orientation[2] = 0; sy.orientation[3] = s * 0.5;
                }
            }
        }
        updateSylinderMap();
        // upload (+ applyBoxBC + cell list) and bring the wrapped positions back into the container
        ck(alens_set_rods_aos(ctx_, nLocal, sylinderContainer.data(), sizeof(Sylinder), 1));
        if (nLocal > 0) {
            const std::unique_ptr<double[]> pos(new double[3 * (size_t)nLocal]); // (not zero-filled: the download fills it)
            ck(alens_get_positions(ctx_, pos.get()));
            ALENS_OMP_FOR
            for (int i = 0; i < nLocal; i++)
                for (int k = 0; k < 3; k++) sylinderContainer[i].pos[k] = pos[3 * (size_t)i + k];
        }
        calcMobOperator();
        conCollectorPtr->clear();
        forcePartNonBrownRcp.reset();
        velocityPartNonBrownRcp.reset();
        velocityBrownRcp.reset();
    }

    void calcMobMatrix() { ck(alens_calc_mobility(ctx_, runConfig.viscosity)); } // :622-717
    void calcMobOperator() { calcMobMatrix(); }                                   // :719-722

    void setForceNonBrown(const std::vector<double> &f) { // :934-939
        if (f.size() != 6 * sylinderContainer.size()) throw std::invalid_argument("forceNonBrown.size() != 6 * nLocal");
        forcePartNonBrownRcp = getTVFromVector(f, commRcp);
    }
    void setVelocityNonBrown(const std::vector<double> &v) { // :941-946
        if (v.size() != 6 * sylinderContainer.size()) throw std::invalid_argument("velNonBrown.size() != 6 * nLocal");
        velocityPartNonBrownRcp = getTVFromVector(v, commRcp);
    }
    /// Brownian velocity is generated by the host application (RNG is out of scope); 6 per rod
    void setVelocityBrown(const std::vector<double> &v) { velocityBrownRcp = getTVFromVector(v, commRcp); }

    void calcVelocityNonCon() { // :724-800
        const int nLocal = (int)sylinderContainer.size();
        velocityNonConRcp = Teuchos::RCP<TV>(std::make_shared<TV>(Teuchos::RCP<const TMAP>(sylinderMobilityMapRcp), true));
        double *v = velocityNonConRcp->data();
        auto monoZero = [&](double *p) {
            if (!runConfig.monolayer) return;
            ALENS_OMP_FOR
            for (int i = 0; i < nLocal; i++) p[6 * i + 2] = p[6 * i + 3] = p[6 * i + 4] = 0;
        };
        if (!forcePartNonBrownRcp.is_null()) {
            ck(alens_mobility_apply(ctx_, forcePartNonBrownRcp->data(), v)); // mobilityOperatorRcp->apply
            monoZero(v);
            const double *f = forcePartNonBrownRcp->data();
            ALENS_OMP_FOR
            for (int i = 0; i < nLocal; i++)
                for (int k = 0; k < 3; k++) {
                    sylinderContainer[i].forceNonB[k] = f[6 * i + k];
                    sylinderContainer[i].torqueNonB[k] = f[6 * i + 3 + k];
                }
        }
        if (!velocityPartNonBrownRcp.is_null()) {
            monoZero(velocityPartNonBrownRcp->data());
            velocityNonConRcp->update(1.0, *velocityPartNonBrownRcp, 1.0);
        }
        ALENS_OMP_FOR
        for (int i = 0; i < nLocal; i++)
            for (int k = 0; k < 3; k++) {
                sylinderContainer[i].velNonB[k] = v[6 * i + k];
                sylinderContainer[i].omegaNonB[k] = v[6 * i + 3 + k];
            }
        if (!velocityBrownRcp.is_null()) {
            monoZero(velocityBrownRcp->data());
            velocityNonConRcp->update(1.0, *velocityBrownRcp, 1.0);
            const double *b = velocityBrownRcp->data();
            ALENS_OMP_FOR
            for (int i = 0; i < nLocal; i++)
                for (int k = 0; k < 3; k++) {
                    sylinderContainer[i].velBrown[k] = b[6 * i + k];
                    sylinderContainer[i].omegaBrown[k] = b[6 * i + 3 + k];
                }
        }
    }

    void collectPairCollision() { // :1152-1160
        long long n = 0;
        ck(alens_collect_pair_collision(ctx_, &n));
    }

    // links prev gid -> next gid (SylinderSystem.hpp: linkMap, addNewLink / getLinkMap, SylinderSystem.cpp:1363-1384)
    std::multimap<int, int> linkMap;
    void addNewLink(const std::vector<Link> &links) {
        for (const auto &l : links) linkMap.emplace(l.prev, l.next);
    }
    const std::multimap<int, int> &getLinkMap() const { return linkMap; }

    // ---- output, as the reference lays it out (SylinderSystem.cpp:478-560): ./result/result<lo>-<hi>/ holds
    // SylinderAscii_<snap>.dat, Sylinder_r<rank>_<snap>.vtp + Sylinder_<snap>.pvtp, ConBlock_r<rank>_<snap>.vtp +
    // ConBlock_<snap>.pvtp; ./TimeStepInfo.txt and ./result/simBox.vtk next to it
    int snapID = 0;
    unsigned restartRngSeed = 0;
    int getSnapID() { return snapID; }
    std::string getCurrentResultFolder() { return getResultFolderWithID(snapID); }
    std::string getResultFolderWithID(int snapID_) {
        const int num = std::max(400 / commRcp->getSize(), 1);
        const int k = snapID_ / num;
        return "./result/result" + std::to_string(k * num) + "-" + std::to_string(k * num + num - 1) + "/";
    }
    bool getIfWriteResultCurrentStep() { return stepCount % static_cast<int>(runConfig.timeSnap / runConfig.dt) == 0; }
    void writeBox() const { // :535-551
        FILE *f = std::fopen("./result/simBox.vtk", "w");
        if (!f) throw std::runtime_error("writeBox: cannot open ./result/simBox.vtk");
        std::fprintf(f, "# vtk DataFile Version 3.0\nvtk file\nASCII\nDATASET RECTILINEAR_GRID\nDIMENSIONS 2 2 2\n");
        const char ax[3] = {'X', 'Y', 'Z'};
        for (int k = 0; k < 3; k++)
            std::fprintf(f, "%c_COORDINATES 2 float\n%g %g\n", ax[k], runConfig.simBoxLow[k], runConfig.simBoxHigh[k]);
        std::fprintf(f, "CELL_DATA 1\nPOINT_DATA 8\n");
        std::fclose(f);
    }
    /// writeResult (:553-560): pulls every block back from the device (gamma written back, stress scaled) for the
    /// constraint file; `makeFolder` is the host's mkdir (the reference uses IOHelper::makeSubFolder)
    void writeResult() {
        const std::string base = getCurrentResultFolder();
        writeAscii(base + "SylinderAscii_" + std::to_string(snapID) + ".dat");
        const int rank = commRcp->getRank(), size = commRcp->getSize();
        Sylinder::writeVTP(sylinderContainer, (int)sylinderContainer.size(), base, std::to_string(snapID), rank);
        conCollectorPtr->pullFromDevice(ctx_, true, true);
        conCollectorPtr->writeVTP(base, "", std::to_string(snapID), rank);
        if (rank == 0) {
            Sylinder::writePVTP(base, std::to_string(snapID), size);
            conCollectorPtr->writePVTP(base, "", std::to_string(snapID), size);
            FILE *f = std::fopen((base + "../../TimeStepInfo.txt").c_str(), "w"); // :509-521
            if (f) {
                std::fprintf(f, "%u\n%u\n%u\nSylinder_%d.pvtp\n", restartRngSeed, (unsigned)stepCount, (unsigned)snapID, snapID);
                std::fclose(f);
            }
        }
        snapID++;
    }

    /// SylinderAscii_<snapID>.dat (SylinderSystem.cpp:489-507): header, one line per rod, then the links `L prev next`
    void writeAscii(const std::string &fileName) const {
        FILE *fptr = std::fopen(fileName.c_str(), "w");
        if (!fptr) throw std::runtime_error("writeAscii: cannot open " + fileName);
        SylinderAsciiHeader header;
        header.nparticle = (int)sylinderContainer.size();
        header.time = stepCount * runConfig.dt;
        header.writeAscii(fptr);
        for (const auto &sy : sylinderContainer) sy.writeAscii(fptr);
        for (const auto &kv : linkMap) std::fprintf(fptr, "L %d %d\n", kv.first, kv.second);
        std::fclose(fptr);
    }

    void collectLinkBilateral() { // :1386-1482, on the device (gid lookup = device hash table instead of the ZDD directory)
        if (linkMap.empty()) return;
        std::vector<int> prev, next;
        prev.reserve(linkMap.size());
        next.reserve(linkMap.size());
        for (const auto &kv : linkMap) {
            prev.push_back(kv.first);
            next.push_back(kv.second);
        }
        long long n = 0;
        ck(alens_collect_link_bilateral(ctx_, prev.data(), next.data(), (long long)prev.size(), runConfig.linkKappa,
                                        runConfig.linkGap, &n));
    }

    void collectBoundaryCollision() { // :1093-1150, on the device; blocks follow the pair collisions in the list
        if (runConfig.boundaries.empty()) return;
        long long n = 0;
        ck(alens_collect_boundary_collision(ctx_, runConfig.boundaries.data(), (int)runConfig.boundaries.size(), &n));
    }

    void resolveConstraints() { // :829-866
        collectPairCollision();
        collectBoundaryCollision();
        collectLinkBilateral();
        // protein constraints (SRC/TubuleSystem.cpp:694-745): the host application pushes those blocks into the pool
        conSolverPtr->setup(*conCollectorPtr, mobilityOperatorRcp, velocityNonConRcp, runConfig.dt);
        conSolverPtr->setControlParams(runConfig.conResTol, runConfig.conMaxIte, runConfig.conSolverChoice);
        conSolverPtr->solveConstraints();
        // writebackGamma is deferred: conSolverPtr->writebackGamma() on snapshot steps (272 B/constraint D2H)
        saveForceVelocityConstraints();
    }

    void saveForceVelocityConstraints() { // :971-1018
        forceUniRcp = conSolverPtr->getForceUni();
        velocityUniRcp = conSolverPtr->getVelocityUni();
        forceBiRcp = conSolverPtr->getForceBi();
        velocityBiRcp = conSolverPtr->getVelocityBi();
        const double *vu = velocityUniRcp->data(), *vb = velocityBiRcp->data();
        const double *fu = forceUniRcp->data(), *fb = forceBiRcp->data();
        const int n = (int)sylinderContainer.size();
        ALENS_OMP_FOR
        for (int i = 0; i < n; i++) {
            auto &sy = sylinderContainer[i];
            for (int k = 0; k < 3; k++) {
                sy.velCol[k] = vu[6 * i + k];     sy.omegaCol[k] = vu[6 * i + 3 + k];
                sy.velBi[k] = vb[6 * i + k];      sy.omegaBi[k] = vb[6 * i + 3 + k];
                sy.forceCol[k] = fu[6 * i + k];   sy.torqueCol[k] = fu[6 * i + 3 + k];
                sy.forceBi[k] = fb[6 * i + k];    sy.torqueBi[k] = fb[6 * i + 3 + k];
            }
        }
    }

    void sumForceVelocity() { // :802-814
        const int nAll = (int)sylinderContainer.size();
        ALENS_OMP_FOR
        for (int i = 0; i < nAll; i++) {
            auto &sy = sylinderContainer[i];
            for (int k = 0; k < 3; k++) {
                sy.vel[k] = sy.velNonB[k] + sy.velBrown[k] + sy.velCol[k] + sy.velBi[k];
                sy.omega[k] = sy.omegaNonB[k] + sy.omegaBrown[k] + sy.omegaCol[k] + sy.omegaBi[k];
                sy.force[k] = sy.forceNonB[k] + sy.forceCol[k] + sy.forceBi[k];
                sy.torque[k] = sy.torqueNonB[k] + sy.torqueCol[k] + sy.torqueBi[k];
            }
        }
    }

    /// Euler step on the device copy (vel = velNonCon + velUni + velBi, quaternion rotated by omega*dt,
    /// Sylinder.cpp:91-99 / EquatnHelper.hpp:74-90), then pos/orientation are mirrored into the container
    void stepEuler() { // :816-827
        if (runConfig.sylinderFixed) return;
        ck(alens_step_euler(ctx_, runConfig.dt));
        if (multiRank_) { // rods that left the slab move to the neighbour rank, on the device (:617-620 in the reference)
            long long sent = 0, received = 0;
            ck(alens_migrate_rods(ctx_, &sent, &received));
            if (sent + received > 0) {
                refillContainerFromDevice();
                return;
            }
        }
        const size_t n = sylinderContainer.size();
        const std::unique_ptr<double[]> pos(new double[3 * n + 1]), q(new double[4 * n + 1]);
        ck(alens_get_rod_state(ctx_, pos.get(), q.get()));
        ALENS_OMP_FOR
        for (size_t i = 0; i < n; i++) {
            for (int k = 0; k < 3; k++) sylinderContainer[i].pos[k] = pos[3 * i + k];
            for (int k = 0; k < 4; k++) sylinderContainer[i].orientation[k] = q[4 * i + k];
        }
    }
    /// the rank's rod set changed: gid, shape, state and group (the rod's tag) come from the device, the per-step fields
    /// (velocities, forces, collision sizes, globalIndex) are rebuilt by the next prepareStep as for every rod
    void refillContainerFromDevice() {
        int n = 0, base = 0;
        ck(alens_get_rod_identity(ctx_, &n, &base, nullptr, nullptr, nullptr, nullptr));
        std::vector<int> gid((size_t)n + 1);
        std::vector<double> len((size_t)n + 1), rad((size_t)n + 1), pos(3 * (size_t)n + 3), q(4 * (size_t)n + 4);
        std::vector<unsigned char> imm((size_t)n + 1);
        std::vector<long long> tag((size_t)n + 1);
        ck(alens_get_rod_identity(ctx_, &n, &base, gid.data(), len.data(), rad.data(), imm.data()));
        ck(alens_get_rod_state(ctx_, pos.data(), q.data()));
        ck(alens_get_rod_tags(ctx_, tag.data()));
        sylinderContainer.assign((size_t)n, Sylinder());
        ALENS_OMP_FOR
        for (int i = 0; i < n; i++) {
            Sylinder &sy = sylinderContainer[i];
            sy.gid = gid[i];
            sy.group = (int)tag[i];
            sy.isImmovable = imm[i] != 0;
            sy.radius = rad[i];
            sy.length = len[i];
            sy.radiusCollision = rad[i] * runConfig.sylinderDiameterColRatio;
            sy.lengthCollision = len[i] * runConfig.sylinderLengthColRatio;
            sy.colBuf = runConfig.sylinderColBuf;
            for (int k = 0; k < 3; k++) sy.pos[k] = pos[3 * (size_t)i + k];
            for (int k = 0; k < 4; k++) sy.orientation[k] = q[4 * (size_t)i + k];
        }
    }

    // :1020-1091 on the device.  The reference draws from per-thread TRNG streams (schedule dependent); here the deviates
    // come from a counter-based generator keyed by (rngSeed, stepCount, gid).  A host application that must reproduce
    // its own stream calls setVelocityBrown() with its numbers instead (or alens_calc_velocity_brown with `normals12`).
    void calcVelocityBrown() {
        const int nLocal = (int)sylinderContainer.size();
        std::vector<double> v(6 * (size_t)nLocal, 0.0);
        ck(alens_calc_velocity_brown(ctx_, runConfig.KBT, runConfig.dt, nullptr, (unsigned long long)runConfig.rngSeed,
                                     (unsigned long long)stepCount, v.data()));
        velocityBrownRcp = getTVFromVector(v, commRcp);
    }

    // ---- per-step diagnostics of the reference (called by SRC/TubuleSystem.cpp after every step).  Their reductions over
    // the ranks go through `sumOverRanks` / `maxOverRanks`: identity on one rank; a multi-rank host program assigns its own
    // (e.g. [](double *v, int n) { MPI_Allreduce(MPI_IN_PLACE, v, n, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD); }), the
    // counterpart of Teuchos::reduceAll on commRcp.
    std::function<void(double *, int)> sumOverRanks = [](double *, int) {};
    std::function<void(int *, int)> maxOverRanks = [](int *, int) {};
    bool printRecords = true; ///< the reference's spdlog "RECORD:" lines on stdout (rank 0)

    struct ConStress {
        double uni[9], bi[9]; ///< row-major, in units of n kBT (scaled by 1 / (nGlobal KBT))
    };
    /// calcConStress (:1226-1263): the gamma-weighted stress sums of the last solve, reduced on the device
    /// (alens_sum_constraint_stress) instead of over a host pool
    ConStress calcConStress() {
        ConStress r;
        ck(alens_sum_constraint_stress(ctx_, 0, r.uni, r.bi));
        double n = (double)sylinderContainer.size();
        sumOverRanks(&n, 1);
        const double scaleFactor = 1 / (n * runConfig.KBT);
        double both[18];
        for (int k = 0; k < 9; k++) {
            both[k] = r.uni[k] * scaleFactor;
            both[9 + k] = r.bi[k] * scaleFactor;
        }
        sumOverRanks(both, 18);
        for (int k = 0; k < 9; k++) {
            r.uni[k] = both[k];
            r.bi[k] = both[9 + k];
        }
        if (printRecords && commRcp->getRank() == 0 && runConfig.logLevel <= 2) {
            std::printf("RECORD: ColXF,%g,%g,%g,%g,%g,%g,%g,%g,%g\n", r.uni[0], r.uni[1], r.uni[2], r.uni[3], r.uni[4], r.uni[5],
                        r.uni[6], r.uni[7], r.uni[8]);
            std::printf("RECORD: BiXF,%g,%g,%g,%g,%g,%g,%g,%g,%g\n", r.bi[0], r.bi[1], r.bi[2], r.bi[3], r.bi[4], r.bi[5], r.bi[6],
                        r.bi[7], r.bi[8]);
        }
        return r;
    }

    struct OrderParameter {
        double p[3], Q[9]; ///< polar vector and nematic tensor, averaged over all rods
    };
    /// calcOrderParameter (:1265-1310): direction = orientation * ez as Eigen evaluates it
    OrderParameter calcOrderParameter() {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0, s8 = 0, s9 = 0, s10 = 0, s11 = 0;
        const int nLocal = (int)sylinderContainer.size();
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11)
#endif
        for (int i = 0; i < nLocal; i++) {
            double d[3];
            sylinderContainer[i].direction(d);
            s0 += d[0]; s1 += d[1]; s2 += d[2];
            s3 += d[0] * d[0] - 1 / 3.0; s4 += d[0] * d[1]; s5 += d[0] * d[2];
            s6 += d[1] * d[0]; s7 += d[1] * d[1] - 1 / 3.0; s8 += d[1] * d[2];
            s9 += d[2] * d[0]; s10 += d[2] * d[1]; s11 += d[2] * d[2] - 1 / 3.0;
        }
        double pQ[13] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, (double)nLocal};
        sumOverRanks(pQ, 13);
        OrderParameter r;
        for (int k = 0; k < 3; k++) r.p[k] = pQ[k] * (1.0 / pQ[12]);
        for (int k = 0; k < 9; k++) r.Q[k] = pQ[3 + k] * (1.0 / pQ[12]);
        if (printRecords && commRcp->getRank() == 0 && runConfig.logLevel <= 2)
            std::printf("RECORD: Order P,%g,%g,%g,Q,%g,%g,%g,%g,%g,%g,%g,%g,%g\n", r.p[0], r.p[1], r.p[2], r.Q[0], r.Q[1], r.Q[2],
                        r.Q[3], r.Q[4], r.Q[5], r.Q[6], r.Q[7], r.Q[8]);
        return r;
    }

    /// calcVolFrac (:292-312): returns the volume fraction (the reference logs it)
    double calcVolFrac() {
        double vol = 0;
        const int nLocal = (int)sylinderContainer.size();
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : vol)
#endif
        for (int i = 0; i < nLocal; i++) {
            const auto &sy = sylinderContainer[i];
            vol += 3.1415926535 * (0.25 * sy.length * std::pow(sy.radius * 2, 2) + std::pow(sy.radius * 2, 3) / 6);
        }
        sumOverRanks(&vol, 1);
        const double boxVolume = (runConfig.simBoxHigh[0] - runConfig.simBoxLow[0]) *
                                 (runConfig.simBoxHigh[1] - runConfig.simBoxLow[1]) *
                                 (runConfig.simBoxHigh[2] - runConfig.simBoxLow[2]);
        if (printRecords && commRcp->getRank() == 0)
            std::printf("Volume Sylinder = %g\nVolume fraction = %g\n", vol, vol / boxVolume);
        return vol / boxVolume;
    }

    /// getMaxGid (:1162-1174): (local, global)
    std::pair<int, int> getMaxGid() {
        int maxGidLocal = 0;
        for (const auto &sy : sylinderContainer) maxGidLocal = std::max(maxGidLocal, sy.gid);
        int maxGidGlobal = maxGidLocal;
        maxOverRanks(&maxGidGlobal, 1);
        return std::pair<int, int>(maxGidLocal, maxGidGlobal);
    }

    /// printTimingSummary (:1484-1489): the device layer's phase times of the last step instead of Teuchos' timers
    void printTimingSummary(const bool zeroOut = true) {
        if (runConfig.timerLevel > 2) return;
        alens_timers t{};
        if (alens_get_timers(ctx_, &t) != ALENS_OK) return;
        std::printf("SylinderSystem timing (ms): upload %g, collect %g, setup %g, solve %g, split %g\n", t.upload_ms, t.collect_ms,
                    t.setup_ms, t.solve_ms, t.split_ms);
        if (zeroOut) alens_reset_timers(ctx_);
    }

    void runStep(bool count_flag = true) { // :948-969 (writeResult is the host's business)
        if (runConfig.KBT > 0 && brownianOnDevice) calcVelocityBrown();
        calcVelocityNonCon();
        resolveConstraints();
        sumForceVelocity();
        stepEuler();
        if (count_flag) stepCount++;
    }

    Teuchos::RCP<TV> getVelocityNonCon() const { return velocityNonConRcp; }
    Teuchos::RCP<const TV> getForceUni() const { return forceUniRcp; }
    Teuchos::RCP<const TV> getVelocityUni() const { return velocityUniRcp; }
    Teuchos::RCP<const TV> getForceBi() const { return forceBiRcp; }
    Teuchos::RCP<const TV> getVelocityBi() const { return velocityBiRcp; }
};

#endif
