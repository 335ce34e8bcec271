// SylinderConfig.hpp -- the run parameters the hot path reads (SimToolbox/Sylinder/SylinderConfig.hpp).
// The YAML parser stays with the host application; the boundary list is the plain-data form of RunConfig::boundaryPtr
// (SimToolbox/Boundary/Boundary.hpp): fill one alens_boundary per SphereShell / Wall / Tube.
#ifndef ALENS_B200_SYLINDERCONFIG_HPP_
#define ALENS_B200_SYLINDERCONFIG_HPP_

#include <vector>

#include "../alens_b200.h"

class SylinderConfig {
  public:
    unsigned int rngSeed = 0;
    int logLevel = 2, timerLevel = 0;
    double simBoxHigh[3] = {1, 1, 1}, simBoxLow[3] = {0, 0, 0};
    bool simBoxPBC[3] = {false, false, false};
    bool monolayer = false;
    int initPreSteps = 100;
    double viscosity = 1.0, KBT = 0.0, linkKappa = 0.0, linkGap = 0.0;
    bool sylinderFixed = false;
    int sylinderNumber = 0;
    double sylinderLength = 0, sylinderLengthSigma = 0, sylinderDiameter = 0;
    double sylinderDiameterColRatio = 1.0, sylinderLengthColRatio = 1.0, sylinderColBuf = 0.3;
    double dt = 1e-5, timeTotal = 0, timeSnap = 0;
    double conResTol = 1e-5;
    int conMaxIte = 100000;
    int conSolverChoice = 0;
    std::vector<alens_boundary> boundaries; // SylinderConfig.hpp: boundaryPtr
};

#endif
