// tests/cpp/test_multirank.cpp -- the C++ mirror on more than one rank: R SylinderSystem objects (one host thread and one
// alens_ctx each, slabs along x), driven like the reference's main program (prepareStep / runStep), against ONE
// SylinderSystem holding the whole suspension.  Brownian steps make rods cross slab faces: they are migrated on the
// device (alens_migrate_rods inside stepEuler), the containers follow, Sylinder::group travels with the rod, globalIndex
// stays the contiguous numbering of updateSylinderMap (SylinderSystem.cpp:868-880).
// usage: test_multirank <nranks> <steps>      devices: ALENS_TEST_DEVICES=0,1,... (default: all ranks on device 0)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <sstream>
#include <thread>

#include "alens_b200/SylinderSystem.hpp"

static std::vector<int> devicesFor(int R) {
    std::vector<int> d(R, 0);
    if (const char *e = std::getenv("ALENS_TEST_DEVICES")) {
        std::vector<int> v;
        std::stringstream ss(e);
        std::string item;
        while (std::getline(ss, item, ',')) v.push_back(std::atoi(item.c_str()));
        for (int r = 0; r < R && !v.empty(); r++) d[r] = v[r % v.size()];
    }
    return d;
}

int main(int argc, char **argv) {
    const int R = argc > 1 ? std::atoi(argv[1]) : 2, steps = argc > 2 ? std::atoi(argv[2]) : 5;
    try {
        SylinderConfig cfg;
        const double box[3] = {2.0 * R, 1.5, 1.5};
        for (int k = 0; k < 3; k++) {
            cfg.simBoxLow[k] = 0;
            cfg.simBoxHigh[k] = box[k];
            cfg.simBoxPBC[k] = true;
        }
        cfg.viscosity = 1.0;
        cfg.KBT = 20.0;
        cfg.dt = 1e-4;
        cfg.sylinderColBuf = 0.025;
        cfg.conResTol = 1e-11;
        cfg.conMaxIte = 20000;
        cfg.rngSeed = 9;
        cfg.initPreSteps = 0;
        const int n = 1500 * R;
        const double L = 0.25, rad = 0.0125;
        std::mt19937_64 gen(123);
        std::uniform_real_distribution<double> u(0.0, 1.0);
        std::normal_distribution<double> g(0.0, 1.0);
        std::vector<Sylinder> all(n);
        for (int i = 0; i < n; i++) {
            Sylinder &sy = all[i];
            sy.gid = i;
            sy.group = i % 7;
            sy.radius = sy.radiusCollision = rad;
            sy.length = sy.lengthCollision = L;
            for (int k = 0; k < 3; k++) sy.pos[k] = u(gen) * box[k];
            const double d[3] = {g(gen), g(gen), g(gen)};
            SylinderSystem::orientationFromDirection(d, sy.orientation);
            sy.isImmovable = (i % 97) == 0;
        }
        // ---- one rank
        std::map<int, Sylinder> ref;
        {
            SylinderSystem one(cfg, all, devicesFor(1)[0]);
            one.setBrownianOnDevice(true);
            for (int s = 0; s < steps; s++) {
                one.prepareStep();
                one.runStep();
            }
            one.prepareStep();
            for (const auto &sy : one.getContainer()) ref[sy.gid] = sy;
        }
        // ---- R ranks
        const std::vector<int> dev = devicesFor(R);
        std::vector<std::vector<Sylinder>> part(R);
        for (const auto &sy : all) part[std::min(R - 1, (int)std::floor(sy.pos[0] / (box[0] / R)))].push_back(sy);
        std::vector<std::unique_ptr<SylinderSystem>> sys(R);
        int base = 0;
        for (int r = 0; r < R; r++) {
            SylinderSystem::Decomposition dec;
            dec.rank = r;
            dec.nranks = R;
            dec.axis = 0;
            dec.skin = 0.1;
            dec.globalMaxBoundingRadius = 0.5 * L + rad;
            dec.globalIndexBase = base;
            dec.maxLocalRods = 2 * n;
            base += (int)part[r].size();
            sys[r].reset(new SylinderSystem());
            sys[r]->initialize(cfg, part[r], dev[r], dec);
            sys[r]->setBrownianOnDevice(true);
        }
        std::vector<SylinderSystem *> ptr;
        for (auto &s : sys) ptr.push_back(s.get());
        SylinderSystem::connectLocal(ptr);
        std::vector<std::string> err(R);
        std::vector<std::thread> th;
        for (int r = 0; r < R; r++)
            th.emplace_back([&, r]() {
                try {
                    for (int s = 0; s < steps; s++) {
                        sys[r]->prepareStep();
                        sys[r]->runStep();
                    }
                    sys[r]->prepareStep();
                } catch (const std::exception &e) {
                    err[r] = e.what();
                }
            });
        for (auto &t : th) t.join();
        for (int r = 0; r < R; r++)
            if (!err[r].empty()) {
                std::fprintf(stderr, "rank %d: %s\n", r, err[r].c_str());
                return 3;
            }
        // ---- compare
        int seen = 0, moved = 0, expectIndex = 0, bad = 0;
        double maxd = 0;
        std::map<int, int> count;
        for (int r = 0; r < R; r++) {
            const double lo = r * box[0] / R, hi = (r + 1) * box[0] / R;
            for (const auto &sy : sys[r]->getContainer()) {
                seen++;
                count[sy.gid]++;
                if (sy.globalIndex != expectIndex++) bad++;
                if (sy.rank != r || sy.group != sy.gid % 7 || sy.isImmovable != ((sy.gid % 97) == 0)) bad++;
                if (!(sy.pos[0] >= lo && sy.pos[0] < hi)) bad++; // prepareStep wrapped it, stepEuler migrated it
                const Sylinder &o = ref.at(sy.gid);
                if (std::floor(all[sy.gid].pos[0] / (box[0] / R)) != r) moved++;
                for (int k = 0; k < 3; k++) {
                    double d = sy.pos[k] - o.pos[k];
                    d -= box[k] * std::round(d / box[k]);
                    maxd = std::max(maxd, std::fabs(d));
                }
            }
        }
        for (const auto &kv : count)
            if (kv.second != 1) bad++;
        std::printf("ranks %d steps %d rods %d seen %d changed_rank %d max_position_difference %.3e inconsistencies %d\n", R, steps,
                    n, seen, moved, maxd, bad);
        if (seen != n || (int)count.size() != n || bad != 0 || moved < 5 || !(maxd < 1e-7)) return 2;
        std::printf("PASS\n");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 5;
    }
    return 0;
}
