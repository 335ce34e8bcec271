// oracle/ref_system_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" driver around the reference's OWN, UNMODIFIED sources, compiled in place from /root/reference
// (never copied into this repo) against the stand-in headers of oracle/stubs/ (mpi.h, Eigen, Teuchos/Kokkos/Tpetra,
// yaml-cpp, TRNG, Zoltan_DD, VTK readers -- none of which exist in this image):
//     SimToolbox/Sylinder/{SylinderSystem,Sylinder,SylinderConfig}.cpp     (+ SylinderNear.hpp, FDPS, DCPQuery.hpp)
//     SimToolbox/Constraint/{ConstraintCollector,ConstraintSolver,ConstraintOperator,BCQPSolver}.cpp
//     SimToolbox/Boundary/Boundary.cpp, SimToolbox/Trilinos/TpetraUtil.cpp
// Recipe: oracle/Makefile `refsys` -> oracle/_ref/libalens_refsys.so (git-ignored, travels to the GPU box).
// What is the reference's and what is not: every line of the pipeline (prepareStep, FDPS tree search, the pair
// functor, boundary and link collection, calcMobMatrix, D^T assembly, ConstraintOperator, BBPGD/APGD, the uni/bi split,
// stepEuler, calcVelocityBrown/NonCon) is reference code; the vector/matrix CONTAINERS underneath it are the stubs,
// one MPI rank.  Tests use it (1) to pin oracle/alens_oracle.c and (2) as the CPU baseline "kind": "reference".
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

#include <mpi.h>

#include "Constraint/BCQPSolver.hpp"
#include "Constraint/ConstraintSolver.hpp"
#include "MPI/MixPairInteraction.hpp"
#include "Sylinder/SylinderSystem.hpp"
#include "Util/Logger.hpp"

static_assert(sizeof(ConstraintBlock) == 272, "ConstraintBlock layout");
static_assert(sizeof(Sylinder) == 568, "Sylinder layout");

namespace {
bool g_init = false;
void initOnce(int nthreads) {
    if (!g_init) {
        int argc = 0;
        char **argv = nullptr;
        MPI_Init(&argc, &argv);
        // FDPS sizes its per-thread buffers from omp_get_max_threads() at Initialize: use every core once
        omp_set_num_threads(omp_get_num_procs());
        FILE *keep = stderr;
        (void)keep;
        PS::Initialize(argc, argv);
        Logger::setup_mpi_spdlog(spdlog::level::err);
        g_init = true;
    }
    if (nthreads > 0) omp_set_num_threads(std::min(nthreads, omp_get_num_procs()));
}

struct Handle {
    SylinderSystem sys;
    std::vector<double> lastHist;
};

void flatten(ConstraintBlockPool &pool, std::vector<ConstraintBlock> &out) {
    out.clear();
    for (auto &q : pool) out.insert(out.end(), q.begin(), q.end());
}

void copyTV(const Teuchos::RCP<const TV> &v, double *out, size_t n) {
    if (!out) return;
    if (v.is_null()) {
        std::fill(out, out + n, 0.0);
        return;
    }
    auto p = v->getLocalView<Kokkos::HostSpace>();
    for (size_t i = 0; i < n; i++) out[i] = p(i, 0);
}
} // namespace


// ---- two-species search (SimToolbox/MPI/MixPairInteraction.hpp), driven like MPI/MixPairInteraction_test.cpp
namespace {
struct MixPoint { // the particle type of that test (Point / Query)
    int gid;
    double RSearch;
    double pos[3];
    inline double getRSearch() const { return RSearch; }
    inline const PS::F64vec3 getPos() const { return PS::F64vec3(pos[0], pos[1], pos[2]); }
    inline void setPos(const PS::F64vec3 &newPos) { pos[0] = newPos.x; pos[1] = newPos.y; pos[2] = newPos.z; }
    inline void copyFromFP(const MixPoint &other) { *this = other; }
};
struct MixCount {
    int nbCount = 0;
    void clear() { nbCount = 0; }
};
struct MixRecord { // every (target, source image) FDPS hands to the functor whose distance passes `r2 <= max(rs)^2`
    std::vector<long long> *out; // trg gid, src gid
    std::vector<double> *dist;
    void operator()(const MixEPI<MixPoint> *const trgPtr, const PS::S32 nTrg, const MixEPJ<MixPoint> *const srcPtr,
                    const PS::S32 nSrc, MixCount *const forcePtr) {
        for (int t = 0; t < nTrg; t++) {
            forcePtr[t].clear();
            if (!trgPtr[t].trgFlag) continue;
            const auto tp = trgPtr[t].getPos();
            const double rt = trgPtr[t].getRSearch();
            for (int s = 0; s < nSrc; s++) {
                if (!srcPtr[s].srcFlag) continue;
                const double r2 = tp.getDistanceSQ(srcPtr[s].getPos());
                const double rr = std::max(rt, srcPtr[s].getRSearch());
                if (r2 <= rr * rr) {
                    forcePtr[t].nbCount++;
#pragma omp critical
                    {
                        out->push_back(trgPtr[t].epTrg.gid);
                        out->push_back(srcPtr[s].epSrc.gid);
                        dist->push_back(std::sqrt(r2));
                    }
                }
            }
        }
    }
};
} // namespace

extern "C" {

int refsys_sizeof_block() { return (int)sizeof(ConstraintBlock); }
int refsys_sizeof_sylinder() { return (int)sizeof(Sylinder); }

// SylinderSystem(configFile, posFile, argc, argv) run inside `workdir` (the reference writes ./result/ there)
void *refsys_create(const char *workdir, const char *yamlPath, const char *posFile, int nthreads) {
    initOnce(nthreads);
    if (workdir && workdir[0] && chdir(workdir) != 0) return nullptr;
    Handle *h = new Handle();
    try {
        h->sys.initialize(SylinderConfig(std::string(yamlPath)), std::string(posFile ? posFile : ""), 0, nullptr);
    } catch (const std::exception &e) {
        fprintf(stderr, "refsys_create: %s\n", e.what());
        delete h;
        return nullptr;
    }
    return h;
}
void refsys_destroy(void *hp) { delete (Handle *)hp; }
void refsys_set_threads(int n) { initOnce(n); }
void refsys_set_log_level(void *, int level) { spdlog::set_level((spdlog::level::level_enum)level); }

int refsys_num_rods(void *hp) { return ((Handle *)hp)->sys.getContainer().getNumberOfParticleLocal(); }
// the container as it is: n Sylinder records of 568 bytes
void refsys_get_sylinders(void *hp, void *out) {
    auto &c = ((Handle *)hp)->sys.getContainerNonConst();
    const int n = c.getNumberOfParticleLocal();
    for (int i = 0; i < n; i++) memcpy((char *)out + (size_t)i * sizeof(Sylinder), &c[i], sizeof(Sylinder));
}
// replace the container's contents (call refsys_prepare_step afterwards)
void refsys_set_rods(void *hp, int n, const int *gid, const double *pos, const double *quat, const double *length,
                     const double *radius, const unsigned char *immovable) {
    auto &c = ((Handle *)hp)->sys.getContainerNonConst();
    c.setNumberOfParticleLocal(n);
    for (int i = 0; i < n; i++) {
        Sylinder sy(gid[i], radius[i], radius[i], length[i], length[i], pos + 3 * i, quat + 4 * i);
        sy.isImmovable = immovable ? immovable[i] != 0 : false;
        c[i] = sy;
    }
}
void refsys_set_config(void *hp, double dt, double conResTol, int conMaxIte, int conSolverChoice, double viscosity,
                       double KBT, double colBuf, double dRatio, double lRatio, double linkKappa, double linkGap,
                       int monolayer) {
    auto &rc = ((Handle *)hp)->sys.runConfig;
    rc.dt = dt; rc.conResTol = conResTol; rc.conMaxIte = conMaxIte; rc.conSolverChoice = conSolverChoice;
    rc.viscosity = viscosity; rc.KBT = KBT; rc.sylinderColBuf = colBuf; rc.sylinderDiameterColRatio = dRatio;
    rc.sylinderLengthColRatio = lRatio; rc.linkKappa = linkKappa; rc.linkGap = linkGap; rc.monolayer = monolayer != 0;
}

void refsys_prepare_step(void *hp) { ((Handle *)hp)->sys.prepareStep(); }
void refsys_set_force_nonbrown(void *hp, const double *f, int n6) {
    ((Handle *)hp)->sys.setForceNonBrown(std::vector<double>(f, f + n6));
}
void refsys_set_velocity_nonbrown(void *hp, const double *v, int n6) {
    ((Handle *)hp)->sys.setVelocityNonBrown(std::vector<double>(v, v + n6));
}
void refsys_calc_velocity_brown(void *hp) { ((Handle *)hp)->sys.calcVelocityBrown(); }
void refsys_calc_velocity_noncon(void *hp) { ((Handle *)hp)->sys.calcVelocityNonCon(); }
void refsys_get_velocity(void *hp, double *velNonCon, double *velBrown, double *velNonBrown) {
    Handle *h = (Handle *)hp;
    const size_t n6 = 6 * (size_t)refsys_num_rods(hp);
    copyTV(h->sys.getVelocityNonCon(), velNonCon, n6);
    copyTV(h->sys.getVelocityBrown(), velBrown, n6);
    copyTV(h->sys.getVelocityNonBrown(), velNonBrown, n6);
}
// y = M x with the reference's mobility matrix (calcMobMatrix, built by prepareStep)
void refsys_mobility_apply(void *hp, const double *x, double *y) {
    Handle *h = (Handle *)hp;
    auto op = h->sys.getMobOperator();
    auto comm = h->sys.getCommRcp();
    const size_t n6 = 6 * (size_t)refsys_num_rods(hp);
    Teuchos::RCP<TV> X = getTVFromVector(std::vector<double>(x, x + n6), comm);
    Teuchos::RCP<TV> Y = Teuchos::rcp(new TV(X->getMap(), true));
    op->apply(*X, *Y);
    copyTV(Y, y, n6);
}

long long refsys_num_constraints(void *hp) {
    long long n = 0;
    for (auto &q : ((Handle *)hp)->sys.getConstraintPoolNonConst()) n += (long long)q.size();
    return n;
}
void refsys_get_constraints(void *hp, void *out) {
    std::vector<ConstraintBlock> flat;
    flatten(((Handle *)hp)->sys.getConstraintPoolNonConst(), flat);
    if (!flat.empty()) memcpy(out, flat.data(), flat.size() * sizeof(ConstraintBlock));
}
void refsys_clear_constraints(void *hp) {
    for (auto &q : ((Handle *)hp)->sys.getConstraintPoolNonConst()) q.clear();
}
void refsys_append_constraints(void *hp, const void *blocks, long long n) {
    auto &q = ((Handle *)hp)->sys.getConstraintPoolNonConst()[0];
    const ConstraintBlock *b = (const ConstraintBlock *)blocks;
    for (long long i = 0; i < n; i++) q.push_back(b[i]);
}
long long refsys_collect_pair_collision(void *hp) {
    ((Handle *)hp)->sys.collectPairCollision();
    return refsys_num_constraints(hp);
}
long long refsys_collect_boundary_collision(void *hp) {
    ((Handle *)hp)->sys.collectBoundaryCollision();
    return refsys_num_constraints(hp);
}
long long refsys_collect_link_bilateral(void *hp) {
    ((Handle *)hp)->sys.collectLinkBilateral();
    return refsys_num_constraints(hp);
}
void refsys_add_links(void *hp, const int *prev, const int *next, int n) {
    std::vector<Link> l(n);
    for (int i = 0; i < n; i++) {
        l[i].prev = prev[i];
        l[i].next = next[i];
    }
    ((Handle *)hp)->sys.addNewLink(l);
}
// collectPairCollision + collectBoundaryCollision + collectLinkBilateral + ConstraintSolver::{setup, solveConstraints,
// writebackGamma} + saveForceVelocityConstraints, appended to whatever the pool already holds
void refsys_resolve_constraints(void *hp) { ((Handle *)hp)->sys.resolveConstraints(); }
void refsys_get_force_velocity(void *hp, double *fu, double *vu, double *fb, double *vb) {
    Handle *h = (Handle *)hp;
    const size_t n6 = 6 * (size_t)refsys_num_rods(hp);
    copyTV(h->sys.getForceUni(), fu, n6);
    copyTV(h->sys.getVelocityUni(), vu, n6);
    copyTV(h->sys.getForceBi(), fb, n6);
    copyTV(h->sys.getVelocityBi(), vb, n6);
}
void refsys_sum_force_velocity(void *hp) { ((Handle *)hp)->sys.sumForceVelocity(); }
void refsys_step_euler(void *hp) { ((Handle *)hp)->sys.stepEuler(); }
void refsys_run_step(void *hp) { ((Handle *)hp)->sys.runStep(false); }
// SylinderSystem::writeResult (SylinderSystem.cpp:553-560): SylinderAscii, Sylinder / ConBlock .vtp + .pvtp, TimeStepInfo
void refsys_write_result(void *hp) { ((Handle *)hp)->sys.writeResult(); }
void refsys_timing_summary(void *hp) { ((Handle *)hp)->sys.printTimingSummary(true); }

// The reference's ConstraintCollector + ConstraintSolver + BCQPSolver on a GIVEN block list (queue 0 of a fresh pool),
// with the system's mobility operator (prepareStep must have run) and a given velocityNonCon.  gamma in list order.
// History rows at full precision come from a second, direct BCQPSolver run on the same operator and q (the
// ConstraintSolver keeps its IteHistory local, ConstraintSolver.cpp:71-93); both runs must agree on gamma.
int refsys_solve_blocks(void *hp, const void *blocksIn, long long n, const double *velNC, double dt, double res, int maxIte,
                        int choice, void *blocksOut, double *gamma, double *fu, double *vu, double *fb, double *vb,
                        double *hist6, int histCap, int *nHist) {
    Handle *h = (Handle *)hp;
    const size_t n6 = 6 * (size_t)refsys_num_rods(hp);
    auto comm = h->sys.getCommRcp();
    Teuchos::RCP<TOP> mobOp = h->sys.getMobOperator();
    Teuchos::RCP<TV> velnc = getTVFromVector(std::vector<double>(velNC, velNC + n6), comm);
    const ConstraintBlock *b = (const ConstraintBlock *)blocksIn;

    ConstraintCollector col;
    auto &q0 = (*col.constraintPoolPtr)[0];
    for (long long i = 0; i < n; i++) q0.push_back(b[i]);
    ConstraintSolver solver;
    solver.setup(col, mobOp, velnc, dt);
    solver.setControlParams(res, maxIte, choice);
    solver.solveConstraints();
    solver.writebackGamma();
    copyTV(solver.getForceUni(), fu, n6);
    copyTV(solver.getVelocityUni(), vu, n6);
    copyTV(solver.getForceBi(), fb, n6);
    copyTV(solver.getVelocityBi(), vb, n6);
    for (long long i = 0; i < n; i++) {
        if (gamma) gamma[i] = q0[i].gamma;
        if (blocksOut) ((ConstraintBlock *)blocksOut)[i] = q0[i];
    }

    if (histCap <= 0 || !hist6) { // no history wanted: one run is enough
        if (nHist) *nHist = 0;
        return 0;
    }
    // the same problem through BCQPSolver directly, to read the IteHistory (steps of ConstraintSolver::setup and
    // ::solveConstraints, ConstraintSolver.cpp:4-34,60-93, spelled out with the reference's public classes)
    ConstraintCollector col2;
    auto &q2 = (*col2.constraintPoolPtr)[0];
    for (long long i = 0; i < n; i++) q2.push_back(b[i]);
    Teuchos::RCP<TCMAT> DMatTrans;
    Teuchos::RCP<TV> delta0, invKappa, biFlag, gam;
    Teuchos::RCP<const TMAP> mobMap = mobOp->getDomainMap();
    col2.buildConstraintMatrixVector(mobMap, DMatTrans, delta0, invKappa, biFlag, gam);
    delta0->scale(1.0 / dt);
    invKappa->scale(1.0 / dt);
    Teuchos::RCP<TV> deltanc = Teuchos::rcp(new TV(delta0->getMap(), true));
    DMatTrans->apply(*velnc, *deltanc);
    Teuchos::RCP<ConstraintOperator> MOp = Teuchos::rcp(new ConstraintOperator(mobOp, DMatTrans, invKappa));
    Teuchos::RCP<TV> qv = Teuchos::rcp(new TV(delta0->getMap(), true));
    qv->update(1.0, *delta0, 1.0, *deltanc, 0.0);
    BCQPSolver bcqp(MOp, qv);
    bcqp.getLowerBound()->scale(-std::numeric_limits<double>::max() * .1, *biFlag);
    IteHistory history;
    int rc = choice == 1 ? bcqp.solveAPGD(gam, res * (1.0 / dt), maxIte, history)
                         : bcqp.solveBBPGD(gam, res * (1.0 / dt), maxIte, history);
    int rows = 0;
    for (auto &r : history) {
        if (rows < histCap && hist6) memcpy(hist6 + 6 * (size_t)rows, r.data(), 6 * sizeof(double));
        rows++;
    }
    if (nHist) *nHist = rows;
    auto gp = gam->getLocalView<Kokkos::HostSpace>();
    for (long long i = 0; i < n; i++)
        if (gp(i, 0) != q0[i].gamma) return -1000 - rc; // the two reference runs disagree: driver error
    return rc;
}

// ConstraintOperator::apply (ConstraintOperator.cpp:30-71) on the D^T the reference's collector builds from a given
// block list: y = (D^T M D + K^-1/dt) x, plus the operator's cached force = D x and vel = M D x
void refsys_operator_apply(void *hp, const void *blocksIn, long long n, double dt, const double *x, double *y, double *force,
                           double *vel) {
    Handle *h = (Handle *)hp;
    const size_t n6 = 6 * (size_t)refsys_num_rods(hp);
    auto comm = h->sys.getCommRcp();
    Teuchos::RCP<TOP> mobOp = h->sys.getMobOperator();
    const ConstraintBlock *b = (const ConstraintBlock *)blocksIn;
    ConstraintCollector col;
    auto &q0 = (*col.constraintPoolPtr)[0];
    for (long long i = 0; i < n; i++) q0.push_back(b[i]);
    Teuchos::RCP<TCMAT> DMatTrans;
    Teuchos::RCP<TV> delta0, invKappa, biFlag, gam;
    Teuchos::RCP<const TMAP> mobMap = mobOp->getDomainMap();
    col.buildConstraintMatrixVector(mobMap, DMatTrans, delta0, invKappa, biFlag, gam);
    invKappa->scale(1.0 / dt);
    ConstraintOperator op(mobOp, DMatTrans, invKappa);
    Teuchos::RCP<TV> X = getTVFromVector(std::vector<double>(x, x + n), comm);
    Teuchos::RCP<TV> Y = Teuchos::rcp(new TV(X->getMap(), true));
    op.apply(*X, *Y);
    copyTV(Y, y, (size_t)n);
    copyTV(op.getForce(), force, n6);
    copyTV(op.getVel(), vel, n6);
}

// BCQPSolver on a caller-supplied CSR matrix, b and bounds (BCQPSolver.hpp:52-97 as a caller sees it).
int refbcqp_solve_csr(int n, const long long *rowPtr, const int *colInd, const double *values, const double *bvec,
                      const double *lb, const double *ub, double *x, double tol, int maxIte, int choice, double *hist6,
                      int histCap, int *nHist, int nthreads) {
    initOnce(nthreads);
    Teuchos::RCP<const TCOMM> comm = getMPIWORLDTCOMM();
    Teuchos::RCP<TMAP> map = getTMAPFromLocalSize(n, comm);
    Teuchos::RCP<const TMAP> cmap = map;
    Kokkos::View<size_t *> rp("rp", n + 1);
    for (int i = 0; i <= n; i++) rp[i] = (size_t)rowPtr[i];
    Kokkos::View<int *> ci("ci", rp[n]);
    Kokkos::View<double *> va("va", rp[n]);
    for (size_t k = 0; k < rp[n]; k++) {
        ci[k] = colInd[k];
        va[k] = values[k];
    }
    Teuchos::RCP<TCMAT> A = Teuchos::rcp(new TCMAT(cmap, cmap, rp, ci, va));
    A->fillComplete(cmap, cmap);
    Teuchos::RCP<TV> bv = getTVFromVector(std::vector<double>(bvec, bvec + n), comm);
    Teuchos::RCP<const TOP> Aop = Teuchos::rcp_dynamic_cast<const TOP>(A, true);
    Teuchos::RCP<const TV> bc = bv;
    BCQPSolver solver(Aop, bc);
    if (lb) solver.setLowerBound(getTVFromVector(std::vector<double>(lb, lb + n), comm));
    if (ub) solver.setUpperBound(getTVFromVector(std::vector<double>(ub, ub + n), comm));
    solver.prepareSolver();
    Teuchos::RCP<TV> xs = getTVFromVector(std::vector<double>(x, x + n), comm);
    IteHistory history;
    const int rc = choice == 1 ? solver.solveAPGD(xs, tol, maxIte, history) : solver.solveBBPGD(xs, tol, maxIte, history);
    copyTV(xs, x, n);
    int rows = 0;
    for (auto &r : history) {
        if (rows < histCap && hist6) memcpy(hist6 + 6 * (size_t)rows, r.data(), 6 * sizeof(double));
        rows++;
    }
    if (nHist) *nHist = rows;
    return rc;
}

// The reference's own self test: BCQPSolver(localSize, diagonal) builds a random SPD problem with random bounds and
// selfTest() solves it, dumping Amat/bvec/lbvec/ubvec/xsol{BBPGD,APGD} as MatrixMarket files into `workdir`
// (BCQPSolver.cpp:38-132,391-429; driven like BCQPSolver_test.cpp).
int refbcqp_selftest(const char *workdir, int localSize, double diagonal, double tol, int maxIte, int choice, int nthreads) {
    initOnce(nthreads);
    if (workdir && workdir[0] && chdir(workdir) != 0) return -1;
    BCQPSolver test(localSize, diagonal);
    return test.selfTest(tol, maxIte, choice);
}

// CalcSylinderNearForce::operator() on one (target, source) pair of SylinderNearEP built from raw fields
// (SylinderNear.hpp:197-519): returns 1 and the pushed ConstraintBlock on a hit
int refsys_pair_functor(const double *posI, const double *dirI, double lcI, double rcI, double colBufI, int gidI,
                        const double *posJ, const double *dirJ, double lcJ, double rcJ, double colBufJ, int gidJ, void *out) {
    initOnce(0);
    SylinderNearEP a, b;
    auto fill = [](SylinderNearEP &e, const double *p, const double *d, double lc, double rc, double cb, int gid) {
        e.gid = gid; e.globalIndex = gid; e.rank = 0;
        e.radius = rc; e.length = lc; e.radiusCollision = rc; e.lengthCollision = lc; e.colBuf = cb;
        for (int k = 0; k < 3; k++) { e.pos[k] = p[k]; e.direction[k] = d[k]; }
    };
    fill(a, posI, dirI, lcI, rcI, colBufI, gidI);
    fill(b, posJ, dirJ, lcJ, rcJ, colBufJ, gidJ);
    auto pool = std::make_shared<ConstraintBlockPool>();
    pool->resize(omp_get_max_threads());
    CalcSylinderNearForce f(pool);
    ForceNear force;
    f(&a, 1, &b, 1, &force);
    for (auto &q : *pool)
        if (!q.empty()) {
            memcpy(out, &q[0], sizeof(ConstraintBlock));
            return 1;
        }
    return 0;
}

// The 12 standard normal deviates per rod that calcVelocityBrown (SylinderSystem.cpp:1053-1056) hands to
// Wrot, Wpos, Wrfdrot, Wrfdpos when a FRESH TRngPool(seed) is consumed by one thread: the same constructor expressions,
// compiled by the same compiler (the evaluation order of the three getN01 calls inside one constructor call is the
// compiler's choice), so that out12 is in the order the reference used them.
void refsys_brown_normals(int seed, int nRods, double *out12) {
    initOnce(0);
    TRngPool pool(seed);
    TRngPool *rngPoolPtr = &pool;
    const int threadId = 0;
    for (int i = 0; i < nRods; i++) {
        Evec3 Wrot(rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId));
        Evec3 Wpos(rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId));
        Evec3 Wrfdrot(rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId));
        Evec3 Wrfdpos(rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId), rngPoolPtr->getN01(threadId));
        for (int k = 0; k < 3; k++) {
            out12[12 * i + k] = Wrot[k];
            out12[12 * i + 3 + k] = Wpos[k];
            out12[12 * i + 6 + k] = Wrfdrot[k];
            out12[12 * i + 9 + k] = Wrfdpos[k];
        }
    }
}

// CalcSylinderNearForce::collideStress (SylinderNear.hpp:432-484), the public static the link and protein code call
void refsys_collide_stress(const double *dirI, const double *dirJ, const double *centerI, const double *centerJ, double lenI,
                           double lenJ, double radI, double radJ, double rho, const double *Ploc, const double *Qloc,
                           double *stress9) {
    Emat3 st;
    CalcSylinderNearForce::collideStress(Evec3(dirI[0], dirI[1], dirI[2]), Evec3(dirJ[0], dirJ[1], dirJ[2]),
                                         Evec3(centerI[0], centerI[1], centerI[2]), Evec3(centerJ[0], centerJ[1], centerJ[2]),
                                         lenI, lenJ, radI, radJ, rho, Evec3(Ploc[0], Ploc[1], Ploc[2]),
                                         Evec3(Qloc[0], Qloc[1], Qloc[2]), st);
    ConstraintBlock blk;
    blk.setStress(st);
    for (int k = 0; k < 9; k++) stress9[k] = blk.stress[k];
}

// Sylinder::calcDragCoeff (Sylinder.cpp:69-82)
void refsys_drag_coeff(double length, double radius, double viscosity, double *para, double *perp, double *rot) {
    const double pos[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1};
    Sylinder sy(0, radius, radius, length, length, pos, q);
    sy.calcDragCoeff(viscosity, *para, *perp, *rot);
}
// Sylinder::stepEuler (Sylinder.cpp:91-99)
void refsys_sylinder_step_euler(double *pos, double *quat, const double *vel, const double *omega, double dt) {
    Sylinder sy(0, 1, 1, 1, 1, pos, quat);
    for (int k = 0; k < 3; k++) { sy.vel[k] = vel[k]; sy.omega[k] = omega[k]; }
    sy.stepEuler(dt);
    memcpy(pos, sy.pos, 24);
    memcpy(quat, sy.orientation, 32);
}
// Boundary::project for the three shipped shapes (Boundary.cpp:25-41,108-124,185-209); type 0 sphere, 1 wall, 2 tube
void refsys_boundary_project(int type, const double *center, const double *axis, double radius, int inside,
                             const double *query, double *project, double *delta) {
    double c[3] = {center[0], center[1], center[2]}, a[3] = {axis[0], axis[1], axis[2]};
    if (type == 0) SphereShell(c, radius, inside != 0).project(query, project, delta);
    else if (type == 1) Wall(c, a).project(query, project, delta);
    else Tube(c, a, radius, inside != 0).project(query, project, delta);
}

// MixPairInteraction<Trg, Src, Trg, Src, Count>: all (target, source image) pairs within max(rsTrg, rsSrc).
// pairs[2 k] = target index, pairs[2 k + 1] = source index, dist[k]; returns the number found (may exceed cap)
long long refmix_search(int nTrg, const double *trgPos, const double *trgRs, int nSrc, const double *srcPos,
                        const double *srcRs, const double *boxLow, const double *boxHigh, const int *pbc, long long cap,
                        long long *pairs, double *dist, int nthreads) {
    initOnce(std::max(1, nthreads));
    PS::ParticleSystem<MixPoint> sysTrg, sysSrc;
    sysTrg.initialize();
    sysSrc.initialize();
    sysTrg.setNumberOfParticleLocal(nTrg);
    sysSrc.setNumberOfParticleLocal(nSrc);
    for (int i = 0; i < nTrg; i++) {
        sysTrg[i].gid = i;
        sysTrg[i].RSearch = trgRs[i];
        for (int k = 0; k < 3; k++) sysTrg[i].pos[k] = trgPos[3 * i + k];
    }
    for (int i = 0; i < nSrc; i++) {
        sysSrc[i].gid = i;
        sysSrc[i].RSearch = srcRs[i];
        for (int k = 0; k < 3; k++) sysSrc[i].pos[k] = srcPos[3 * i + k];
    }
    PS::DomainInfo dinfo;
    dinfo.initialize();
    const int code = (pbc[0] ? 1 : 0) + (pbc[1] ? 2 : 0) + (pbc[2] ? 4 : 0);
    const PS::BOUNDARY_CONDITION bc[8] = {PS::BOUNDARY_CONDITION_OPEN,        PS::BOUNDARY_CONDITION_PERIODIC_X,
                                          PS::BOUNDARY_CONDITION_PERIODIC_Y,  PS::BOUNDARY_CONDITION_PERIODIC_XY,
                                          PS::BOUNDARY_CONDITION_PERIODIC_Z,  PS::BOUNDARY_CONDITION_PERIODIC_XZ,
                                          PS::BOUNDARY_CONDITION_PERIODIC_YZ, PS::BOUNDARY_CONDITION_PERIODIC_XYZ};
    dinfo.setBoundaryCondition(bc[code]);
    dinfo.setPosRootDomain(PS::F64vec3(boxLow[0], boxLow[1], boxLow[2]), PS::F64vec3(boxHigh[0], boxHigh[1], boxHigh[2]));
    dinfo.decomposeDomainAll(sysSrc);
    sysSrc.exchangeParticle(dinfo);
    sysTrg.exchangeParticle(dinfo);
    MixPairInteraction<MixPoint, MixPoint, MixPoint, MixPoint, MixCount> mix;
    mix.initialize();
    mix.updateSystem(sysTrg, sysSrc, dinfo);
    mix.updateTree();
    std::vector<long long> out;
    std::vector<double> d;
    MixRecord ftr{&out, &d};
    mix.computeForce<MixRecord>(ftr, dinfo);
    const long long n = (long long)d.size();
    for (long long k = 0; k < std::min(n, cap); k++) {
        pairs[2 * k] = out[2 * k];
        pairs[2 * k + 1] = out[2 * k + 1];
        dist[k] = d[k];
    }
    return n;
}

// SylinderConfig::SylinderConfig(file) (SylinderConfig.cpp:6-90): the parsed run parameters as a flat array, for the test
// of the mirror's own reader.  out[0..33]; returns the number of boundaries.
int refsys_parse_config(const char *yamlPath, double *out) {
    initOnce(0);
    SylinderConfig c(yamlPath);
    int k = 0;
    out[k++] = c.rngSeed; out[k++] = c.logLevel; out[k++] = c.timerLevel;
    for (int d = 0; d < 3; d++) out[k++] = c.simBoxLow[d];
    for (int d = 0; d < 3; d++) out[k++] = c.simBoxHigh[d];
    for (int d = 0; d < 3; d++) out[k++] = c.simBoxPBC[d] ? 1 : 0;
    out[k++] = c.monolayer ? 1 : 0;
    for (int d = 0; d < 3; d++) out[k++] = c.initBoxLow[d];
    for (int d = 0; d < 3; d++) out[k++] = c.initBoxHigh[d];
    for (int d = 0; d < 3; d++) out[k++] = c.initOrient[d];
    out[k++] = c.initCircularX ? 1 : 0; out[k++] = c.initPreSteps;
    out[k++] = c.viscosity; out[k++] = c.KBT; out[k++] = c.linkKappa; out[k++] = c.linkGap;
    out[k++] = c.sylinderFixed ? 1 : 0; out[k++] = c.sylinderNumber; out[k++] = c.sylinderLength;
    out[k++] = c.sylinderLengthSigma; out[k++] = c.sylinderDiameter; out[k++] = c.sylinderDiameterColRatio;
    out[k++] = c.sylinderLengthColRatio; out[k++] = c.sylinderColBuf;
    out[k++] = c.dt; out[k++] = c.timeTotal; out[k++] = c.timeSnap; out[k++] = c.conResTol; out[k++] = c.conMaxIte;
    out[k++] = c.conSolverChoice;
    out[k] = k; // self-check: number of values written before this one
    return (int)c.boundaryPtr.size();
}
}
