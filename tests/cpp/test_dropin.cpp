// tests/cpp/test_dropin.cpp -- drives the C++ drop-in classes (include/alens_b200/*.hpp) the way the
// reference's SylinderSystem_main.cpp:36-42 / SRC/TubuleSystem.cpp:144-230 drive the originals, and dumps
// the results as raw doubles for tests/test_gpu_dropin.py to compare with the ctypes path and the oracle.
//
//   test_dropin <in.bin> <out.bin> [nsteps]
// in.bin : int n, double boxlo[3], boxhi[3], int pbc[3], double colbuf, mu, dt, res, int maxIte,
//          then n x {int gid, double radius, length, pos[3], quat[4]}, then 6n doubles velNonBrown,
//          int nb, nb x ConstraintBlock (272 B) host blocks pushed into the pool before runStep
// out.bin: long long nc, int iterations, double residual, 6n x4 doubles (fU, vU, fB, vB), n x (pos[3],quat[4])
//          after the step, nc doubles gamma, nc x ConstraintBlock after writebackGamma, then the per-step diagnostics:
//          18 doubles calcConStress (uni, bi), 12 doubles calcOrderParameter (p, Q), 1 double calcVolFrac, 2 ints getMaxGid
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "alens_b200/SylinderSystem.hpp"

template <class T>
static void rd(FILE *f, T *p, size_t n = 1) {
    if (fread(p, sizeof(T), n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
}
template <class T>
static void wr(FILE *f, const T *p, size_t n = 1) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char **argv) {
    if (argc < 3) return 1;
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 1;
    SylinderConfig cfg;
    int n, pbc[3], maxIte, nb;
    rd(fi, &n);
    rd(fi, cfg.simBoxLow, 3);
    rd(fi, cfg.simBoxHigh, 3);
    rd(fi, pbc, 3);
    for (int k = 0; k < 3; k++) cfg.simBoxPBC[k] = pbc[k] != 0;
    rd(fi, &cfg.sylinderColBuf);
    rd(fi, &cfg.viscosity);
    rd(fi, &cfg.dt);
    rd(fi, &cfg.conResTol);
    rd(fi, &maxIte);
    cfg.conMaxIte = maxIte;
    cfg.initPreSteps = 0;
    cfg.KBT = 0.5; // scale of calcConStress (no Brownian motion: the velocities are the test's input)
    std::vector<Sylinder> rods(n);
    for (int i = 0; i < n; i++) {
        int gid;
        double radius, length, pos[3], q[4];
        rd(fi, &gid); rd(fi, &radius); rd(fi, &length); rd(fi, pos, 3); rd(fi, q, 4);
        rods[i] = Sylinder(gid, radius, radius, length, length, pos, q);
    }
    std::vector<double> vnb(6 * (size_t)n);
    rd(fi, vnb.data(), vnb.size());
    rd(fi, &nb);
    std::vector<ConstraintBlock> host(nb);
    if (nb) rd(fi, host.data(), nb);
    fclose(fi);

    try {
        SylinderSystem sys(cfg, rods, 0);
        // one step, in the order TubuleSystem::step uses: prepareStep, set non-Brownian input, push host
        // blocks into the per-thread pool, runStep
        sys.prepareStep();
        sys.setVelocityNonBrown(vnb);
        auto &pool = sys.getConstraintPoolNonConst();
        for (int i = 0; i < nb; i++) pool[i % pool.size()].push_back(host[i]);
        // the pool is flattened queue by queue: remember that order for the caller
        std::vector<ConstraintBlock> hostOrder = sys.getConstraintCollector()->flatten();
        sys.runStep();
        auto &solver = *sys.getConstraintSolver();
        solver.printRecord(stderr);
        const auto rep = solver.getReport();
        FILE *fo = fopen(argv[2], "wb");
        wr(fo, &rep.n_constraints);
        wr(fo, &rep.iterations);
        wr(fo, &rep.residual);
        wr(fo, sys.getForceUni()->data(), 6 * (size_t)n);
        wr(fo, sys.getVelocityUni()->data(), 6 * (size_t)n);
        wr(fo, sys.getForceBi()->data(), 6 * (size_t)n);
        wr(fo, sys.getVelocityBi()->data(), 6 * (size_t)n);
        for (auto &sy : sys.getContainer()) {
            wr(fo, sy.pos, 3);
            wr(fo, sy.orientation, 4);
        }
        auto g = solver.getGamma();
        wr(fo, g->data(), (size_t)rep.n_constraints);
        sys.printRecords = false;
        const auto stress = sys.calcConStress(); // device-side sum: before the pool is refilled from the device
        solver.writebackGamma();
        auto blocks = sys.getConstraintCollector()->flatten();
        if ((long long)blocks.size() != rep.n_constraints) return 3;
        wr(fo, blocks.data(), blocks.size());
        // velCol of rod 0 must equal velocityUni[0..2] (saveForceVelocityConstraints)
        const auto &s0 = sys.getContainer()[0];
        if (n > 0 && s0.velCol[0] != sys.getVelocityUni()->data()[0]) return 4;
        wr(fo, stress.uni, 9);
        wr(fo, stress.bi, 9);
        { // ... and the same sums from the refilled pool, as the reference takes them
            double u[9], b[9];
            sys.getConstraintCollector()->sumLocalConstraintStress(u, b, false);
            for (int k = 0; k < 9; k++) {
                const double su = u[k] / (n * cfg.KBT), sb = b[k] / (n * cfg.KBT);
                if (std::fabs(su - stress.uni[k]) > 1e-12 * (1 + std::fabs(su)) || std::fabs(sb - stress.bi[k]) > 1e-12 * (1 + std::fabs(sb)))
                    return 6;
            }
        }
        const auto op = sys.calcOrderParameter();
        wr(fo, op.p, 3);
        wr(fo, op.Q, 9);
        const double phi = sys.calcVolFrac();
        wr(fo, &phi);
        const auto mg = sys.getMaxGid();
        wr(fo, &mg.first);
        wr(fo, &mg.second);
        sys.printTimingSummary();
        fclose(fo);
    } catch (const std::exception &e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 5;
    }
    return 0;
}
