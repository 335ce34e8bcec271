// tests/cpp/test_system_main.cpp -- the mirror driven exactly like the reference's own main program
// (SimToolbox/Sylinder/SylinderSystem_main.cpp:16-48): a RunConfig.yaml and a SylinderInitial.dat in the working directory,
//   SylinderSystem system(runConfig, posFile, argc, argv);  loop { prepareStep(); runStep(); }  + writeResult()
// usage: test_system_main <nsteps> <out.bin> [restart]   (cwd holds RunConfig.yaml and SylinderInitial.dat, ./result/result0-399 exists)
// restart: afterwards a second system resumes from the snapshot just written (the restart branch of the reference's main
// programs: reinitialize(runConfig, "TimeStepInfo.txt", argc, argv), SylinderSystem.cpp:106-175) and is compared with the first
// out.bin: int n, n x Sylinder (568 B) after the steps, int nlinks, nlinks x (prev, next)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "alens_b200/SylinderSystem.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 1;
    const int nsteps = atoi(argv[1]);
    try {
        std::string runConfig = "RunConfig.yaml";
        std::string posFile = "SylinderInitial.dat";
        SylinderSystem system(runConfig, posFile, argc, argv);
        for (int i = 0; i < nsteps; i++) {
            system.prepareStep();
            system.runStep();
        }
        system.prepareStep(); // wrapped positions, as the next step would see them
        system.writeResult();
        FILE *f = fopen(argv[2], "wb");
        const int n = (int)system.getContainer().size();
        fwrite(&n, 4, 1, f);
        fwrite(system.getContainer().data(), sizeof(Sylinder), n, f);
        const int nl = (int)system.getLinkMap().size();
        fwrite(&nl, 4, 1, f);
        for (const auto &kv : system.getLinkMap()) {
            fwrite(&kv.first, 4, 1, f);
            fwrite(&kv.second, 4, 1, f);
        }
        fclose(f);
        if (argc > 3 && std::string(argv[3]) == "restart") {
            const SylinderConfig cfg(runConfig);
            auto fail = [](const char *what) {
                fprintf(stderr, "restart: %s\n", what);
                return 6;
            };
            // one step by hand with the snapshot between the solve and the move, where runStep takes it (:948-969): the
            // records then carry the velocities the restart steps with
            system.prepareStep();
            system.calcVelocityNonCon();
            system.resolveConstraints();
            system.sumForceVelocity();
            system.writeResult();
            system.stepEuler();
            SylinderSystem resumed;
            resumed.reinitialize(cfg, "TimeStepInfo.txt", argc, argv);
            if (resumed.getStepCount() != system.getStepCount() + 1 || resumed.getSnapID() != system.getSnapID()) return fail("counters");
            if (resumed.getLinkMap() != system.getLinkMap()) return fail("link map");
            if (resumed.runConfig.rngSeed != cfg.rngSeed + 1) return fail("rng seed");
            if ((int)resumed.getContainer().size() != n) return fail("rod count");
            double vmax = 0;
            for (int i = 0; i < n; i++) {
                const Sylinder &a = system.getContainer()[i], &b = resumed.getContainer()[i];
                if (a.gid != b.gid || a.group != b.group || a.isImmovable != b.isImmovable) return fail("identity");
                if (b.length != (double)(float)a.length || b.radius != (double)(float)a.radius) return fail("shape");
                double da[3], db[3];
                a.direction(da);
                b.direction(db);
                for (int k = 0; k < 3; k++) {
                    vmax = std::max(vmax, std::fabs(a.vel[k]));
                    // Float64 end points, Float32 velocities and axes in the file
                    if (std::fabs(a.pos[k] - b.pos[k]) > 1e-13 * (1 + std::fabs(a.pos[k])) + 1e-6 * std::fabs(a.vel[k]) * cfg.dt)
                        return fail("position after the Euler step of the stored velocity");
                    if (std::fabs(da[k] - db[k]) > 3e-7) return fail("axis after the Euler step");
                }
            }
            if (!(vmax > 0)) return fail("the snapshot holds no velocities");
            // one more step on both: the resumed run follows the original to the precision of the snapshot (Float32 axes)
            system.prepareStep();
            system.runStep();
            resumed.prepareStep();
            resumed.runStep();
            double moved = 0, diff = 0;
            for (int i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) {
                    moved = std::max(moved, std::fabs(system.getContainer()[i].vel[k]) * cfg.dt);
                    diff = std::max(diff, std::fabs(system.getContainer()[i].pos[k] - resumed.getContainer()[i].pos[k]));
                }
            fprintf(stderr, "restart: step after resume differs by %g (step size %g)\n", diff, moved);
            if (!(moved > 0) || diff > 1e-3 * moved + 1e-9) return fail("trajectory after the restart");
            printf("restart ok\n");
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 5;
    }
    return 0;
}
