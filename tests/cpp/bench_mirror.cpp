// tests/cpp/bench_mirror.cpp -- wall clock per step of the reference's main loop (prepareStep(); runStep();) through the C++
// mirror at the bench size: what a host application that switches its includes (option A) gets, host loops
// over the 568-byte Sylinder records included.  Not a test; `make bench_mirror && ./bench_mirror [nRods] [steps]`.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "alens_b200/SylinderSystem.hpp"

int main(int argc, char **argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 1000000, steps = argc > 2 ? std::atoi(argv[2]) : 5;
    try {
        const double L = 0.25, rad = 0.0125, phi = 0.10;
        const double vol = 3.14159265358979323846 * rad * rad * L + 4.0 / 3.0 * 3.14159265358979323846 * rad * rad * rad;
        const double edge = std::cbrt(n * vol / phi);
        SylinderConfig cfg;
        for (int k = 0; k < 3; k++) {
            cfg.simBoxLow[k] = 0;
            cfg.simBoxHigh[k] = edge;
            cfg.simBoxPBC[k] = true;
        }
        cfg.viscosity = 1.0;
        cfg.KBT = 0.0;
        cfg.dt = 1e-5;
        cfg.sylinderColBuf = 0.025;
        cfg.conResTol = 1e-5;
        cfg.conMaxIte = 10000;
        cfg.initPreSteps = 0;
        std::mt19937_64 gen(1234);
        std::uniform_real_distribution<double> u(0.0, 1.0);
        std::normal_distribution<double> g(0.0, 1.0);
        std::vector<Sylinder> rods(n);
        for (int i = 0; i < n; i++) {
            Sylinder &sy = rods[i];
            sy.gid = i;
            sy.radius = sy.radiusCollision = rad;
            sy.length = sy.lengthCollision = L;
            for (int k = 0; k < 3; k++) sy.pos[k] = u(gen) * edge;
            const double d[3] = {g(gen), g(gen), g(gen)};
            SylinderSystem::orientationFromDirection(d, sy.orientation);
        }
        SylinderSystem sys(cfg, rods, 0);
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        for (int s = 0; s < 4; s++) { // relaxation of the overlapping start (untimed)
            sys.prepareStep();
            sys.runStep();
        }
        double tPrep = 0, tRun = 0;
        for (int s = 0; s < steps; s++) {
            const auto t0 = now();
            sys.prepareStep();
            const auto t1 = now();
            sys.runStep();
            const auto t2 = now();
            tPrep += ms(t0, t1);
            tRun += ms(t1, t2);
        }
        // the per-step stress diagnostic (SylinderSystem::calcConStress): device-side reduction against the reference's way,
        // a walk over the host pool (which the mirror first has to refill from the device)
        sys.runConfig.KBT = 1.0;
        sys.printRecords = false;
        sys.calcConStress();
        const auto s0 = now();
        const auto cs = sys.calcConStress();
        const auto s1 = now();
        sys.getConstraintCollector()->pullFromDevice(sys.deviceContext(), true, true);
        double uni[9], bi[9];
        sys.getConstraintCollector()->sumLocalConstraintStress(uni, bi, false);
        const auto s2 = now();
        double err = 0, scale = 0;
        for (int k = 0; k < 9; k++) {
            err = std::max(err, std::fabs(uni[k] / n - cs.uni[k]));
            scale = std::max(scale, std::fabs(cs.uni[k]));
        }
        std::printf("{\"calcConStress_device_ms\": %.3f, \"pool_refill_and_host_sum_ms\": %.3f, \"rel_diff\": %.3g}\n", ms(s0, s1),
                    ms(s1, s2), err / scale);
        const auto &rep = sys.getConstraintSolver()->getReport();
        std::printf("{\"mirror_ms_per_step\": %.3f, \"prepareStep_ms\": %.3f, \"runStep_ms\": %.3f, \"rods\": %d, \"constraints\": %lld, "
                    "\"bbpgd_iterations\": %d, \"steps\": %d}\n",
                    (tPrep + tRun) / steps, tPrep / steps, tRun / steps, n, (long long)rep.n_constraints, rep.iterations, steps);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 5;
    }
    return 0;
}
