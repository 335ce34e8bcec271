// tests/cpp/test_system_main.cpp -- the mirror driven exactly like the reference's own main program
// (SimToolbox/Sylinder/SylinderSystem_main.cpp:16-48): a RunConfig.yaml and a SylinderInitial.dat in the working directory,
//   SylinderSystem system(runConfig, posFile, argc, argv);  loop { prepareStep(); runStep(); }  + writeResult()
// usage: test_system_main <nsteps> <out.bin>     (cwd holds RunConfig.yaml and SylinderInitial.dat, ./result/result0-399 exists)
// out.bin: int n, n x Sylinder (568 B) after the steps, int nlinks, nlinks x (prev, next)
#include <cstdio>
#include <cstdlib>

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
    } catch (const std::exception &e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 5;
    }
    return 0;
}
