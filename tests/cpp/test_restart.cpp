// tests/cpp/test_restart.cpp -- the restart path of the C++ mirror (SylinderSystem::reinitialize, SylinderSystem.cpp:106-175;
// setInitialFromVTKFile :406-476) without a GPU: rods with velocities are written as Sylinder_r<rank>_<snap>.vtp pieces +
// Sylinder_<snap>.pvtp + SylinderAscii_<snap>.dat + TimeStepInfo.txt by the mirror's writers, read back, and compared.
//   test_restart <folder>                    (the folder must exist; files go to <folder>/result/result0-399/)
//   test_restart read <file.pvtp> <out.bin>  rods of a snapshot as raw 568-byte Sylinder records (tests/test_cpp_restart.py
//                                            reads the REFERENCE's own files through this)
//   test_restart euler <in.bin> <dt> <out.bin>  Sylinder::stepEuler(dt) on raw Sylinder records (compared with the reference's)
#include <cmath>
#include <cstdio>
#include <random>

#include "alens_b200/SylinderSystem.hpp"

#define CHECK(c)                                                                                                       \
    do {                                                                                                               \
        if (!(c)) {                                                                                                    \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);                                        \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (std::string(argv[1]) == "euler") {
        if (argc < 5) return 2;
        FILE *f = std::fopen(argv[2], "rb");
        CHECK(f);
        std::vector<Sylinder> rods;
        Sylinder one;
        while (std::fread((void *)&one, sizeof(Sylinder), 1, f) == 1) rods.push_back(one);
        std::fclose(f);
        const double dt = std::atof(argv[3]);
        for (auto &sy : rods) sy.stepEuler(dt);
        f = std::fopen(argv[4], "wb");
        CHECK(f);
        std::fwrite((const void *)rods.data(), sizeof(Sylinder), rods.size(), f);
        std::fclose(f);
        return 0;
    }
    if (std::string(argv[1]) == "read") {
        if (argc < 4) return 2;
        std::vector<Sylinder> rods;
        try {
            rods = SylinderSystem::readSylinderVTK(argv[2]);
        } catch (const std::exception &e) {
            std::fprintf(stderr, "exception: %s\n", e.what());
            return 5;
        }
        FILE *f = std::fopen(argv[3], "wb");
        CHECK(f);
        std::fwrite((const void *)rods.data(), sizeof(Sylinder), rods.size(), f);
        std::fclose(f);
        return 0;
    }
    const std::string base = std::string(argv[1]) + "/result/result0-399/";
    std::mt19937_64 gen(7);
    std::uniform_real_distribution<double> u(-1, 1);
    const int n = 257, snap = 12;
    std::vector<Sylinder> rods(n);
    for (int i = 0; i < n; i++) {
        Sylinder &sy = rods[i];
        sy.gid = 3 * i + 1;
        sy.group = i % 5 - 1;
        sy.isImmovable = i % 7 == 0;
        sy.radius = 0.0125 * (1 + 0.25 * (i % 3));
        sy.radiusCollision = sy.radius * 0.5;
        sy.length = i % 11 == 0 ? 0.0 : 0.3 + 0.2 * u(gen);
        sy.lengthCollision = sy.length * 0.75;
        double d[3] = {u(gen), u(gen), i == 5 ? -50.0 : u(gen)};
        if (i == 6) d[0] = d[1] = 0, d[2] = -1; // antiparallel to z: the half-turn branch of FromTwoVectors
        SylinderSystem::orientationFromDirection(d, sy.orientation);
        for (int k = 0; k < 3; k++) {
            sy.pos[k] = 10 * u(gen);
            sy.vel[k] = u(gen);
            sy.omega[k] = i % 13 == 0 ? 0.0 : 3 * u(gen);
        }
    }
    // two piece files (ranks 0 and 1) + the index, as a 2-rank run of the reference leaves them
    const int n0 = 100;
    std::vector<Sylinder> part0(rods.begin(), rods.begin() + n0), part1(rods.begin() + n0, rods.end());
    Sylinder::writeVTP(part0, n0, base, std::to_string(snap), 0);
    Sylinder::writeVTP(part1, n - n0, base, std::to_string(snap), 1);
    Sylinder::writePVTP(base, std::to_string(snap), 2);

    const std::vector<Sylinder> back = SylinderSystem::readSylinderVTK(base + "Sylinder_" + std::to_string(snap) + ".pvtp");
    CHECK((int)back.size() == n);
    for (int i = 0; i < n; i++) {
        const Sylinder &a = rods[i], &b = back[i];
        CHECK(a.gid == b.gid && a.group == b.group && a.isImmovable == b.isImmovable);
        // scalars and velocities travel as Float32, the end points as Float64
        CHECK(b.radius == (double)(float)a.radius && b.radiusCollision == (double)(float)a.radiusCollision);
        CHECK(b.length == (double)(float)a.length && b.lengthCollision == (double)(float)a.lengthCollision);
        double da[3], db[3];
        a.direction(da);
        b.direction(db);
        for (int k = 0; k < 3; k++) {
            CHECK(std::fabs(a.pos[k] - b.pos[k]) < 1e-14 * (1 + std::fabs(a.pos[k])));
            CHECK(b.vel[k] == (double)(float)a.vel[k] && b.omega[k] == (double)(float)a.omega[k]);
            CHECK(std::fabs(da[k] - db[k]) < 2e-7); // znorm is Float32
        }
    }
    // a single piece is read directly as well
    CHECK((int)SylinderSystem::readSylinderVTK(base + "Sylinder_r1_" + std::to_string(snap) + ".vtp").size() == n - n0);

    // restart file -> names (TimeStepInfo.txt as writeResult writes it)
    {
        FILE *f = std::fopen((std::string(argv[1]) + "/TimeStepInfo.txt").c_str(), "w");
        CHECK(f);
        std::fprintf(f, "%u\n%u\n%u\nSylinder_%d.pvtp\n", 41u, 1200u, (unsigned)snap, snap);
        std::fclose(f);
    }
    const auto info = SylinderSystem::readRestartFile(std::string(argv[1]) + "/TimeStepInfo.txt");
    CHECK(info.rngSeed == 41 && info.stepCount == 1200 && info.snapID == snap);
    CHECK(info.pvtpFileName == "Sylinder_12.pvtp" && info.asciiFileName == "SylinderAscii_12.dat");

    // the host Euler step against a direct evaluation of rotateEquatn (EquatnHelper.hpp:74-90)
    for (int i = 0; i < n; i++) {
        Sylinder s = back[i];
        const double dt = 1e-3;
        s.stepEuler(dt);
        const Sylinder &o = back[i];
        const double w = std::sqrt(o.omega[0] * o.omega[0] + o.omega[1] * o.omega[1] + o.omega[2] * o.omega[2]);
        for (int k = 0; k < 3; k++) CHECK(s.pos[k] == o.pos[k] + o.vel[k] * dt);
        double nrm = 0;
        for (int k = 0; k < 4; k++) nrm += s.orientation[k] * s.orientation[k];
        CHECK(std::fabs(nrm - 1) < 1e-14);
        if (w == 0) {
            for (int k = 0; k < 4; k++) CHECK(s.orientation[k] == o.orientation[k]);
            continue;
        }
        // the rotated direction equals Rodrigues' rotation of the old one about omega by |omega| dt
        double d0[3], d1[3], ax[3] = {o.omega[0] / w, o.omega[1] / w, o.omega[2] / w};
        o.direction(d0);
        s.direction(d1);
        const double th = w * dt, c = std::cos(th), sn = std::sin(th);
        const double dot = ax[0] * d0[0] + ax[1] * d0[1] + ax[2] * d0[2];
        const double cr[3] = {ax[1] * d0[2] - ax[2] * d0[1], ax[2] * d0[0] - ax[0] * d0[2], ax[0] * d0[1] - ax[1] * d0[0]};
        for (int k = 0; k < 3; k++) CHECK(std::fabs(d1[k] - (d0[k] * c + cr[k] * sn + ax[k] * dot * (1 - c))) < 1e-10);
    }
    std::printf("restart ok: %d rods\n", n);
    return 0;
}
