// SylinderConfig.hpp -- the run parameters of SimToolbox/Sylinder/SylinderConfig.{hpp,cpp}, including the
// constructor from a RunConfig.yaml file (SylinderConfig.cpp:6-90: required / optional keys and their defaults).
// The reference parses with yaml-cpp; here a small reader for the subset those files use (block mappings `key: value`,
// flow sequences `[a, b, c]`, comments, and the `boundaries:` block sequence of mappings).  The boundary list is the
// plain-data form of RunConfig::boundaryPtr (SimToolbox/Boundary/Boundary.hpp): one alens_boundary per
// SphereShell / Wall / Tube.
#ifndef ALENS_B200_SYLINDERCONFIG_HPP_
#define ALENS_B200_SYLINDERCONFIG_HPP_

#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../alens_b200.h"

class SylinderConfig {
  public:
    unsigned int rngSeed = 0;
    int logLevel = 2, timerLevel = 0;
    double simBoxHigh[3] = {1, 1, 1}, simBoxLow[3] = {0, 0, 0};
    bool simBoxPBC[3] = {false, false, false};
    bool monolayer = false;
    double initBoxHigh[3] = {1, 1, 1}, initBoxLow[3] = {0, 0, 0};
    double initOrient[3] = {2, 2, 2};
    bool initCircularX = false;
    int initPreSteps = 100;
    double thermEquilTime = 0;
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

    SylinderConfig() = default;

    /// SylinderConfig::SylinderConfig(std::string filename), SylinderConfig.cpp:6-90
    explicit SylinderConfig(const std::string &filename) {
        Yaml y(filename);
        // required (SylinderConfig.cpp:11-28): a missing key is fatal in the reference (spdlog::critical + exit)
        y.get("rngSeed", rngSeed, true);
        y.get("simBoxLow", simBoxLow, 3, true);
        y.get("simBoxHigh", simBoxHigh, 3, true);
        y.get("simBoxPBC", simBoxPBC, 3, true);
        y.get("viscosity", viscosity, true);
        y.get("KBT", KBT, true);
        y.get("sylinderNumber", sylinderNumber, true);
        y.get("sylinderLength", sylinderLength, true);
        y.get("sylinderDiameter", sylinderDiameter, true);
        y.get("dt", dt, true);
        y.get("timeTotal", timeTotal, true);
        y.get("timeSnap", timeSnap, true);
        y.get("conResTol", conResTol, true);
        y.get("conMaxIte", conMaxIte, true);
        y.get("conSolverChoice", conSolverChoice, true);
        // optional, with the reference's defaults (:30-73)
        logLevel = 2; // spdlog::level::info
        y.get("logLevel", logLevel);
        timerLevel = logLevel;
        y.get("timerLevel", timerLevel);
        monolayer = false;
        y.get("monolayer", monolayer);
        for (int k = 0; k < 3; k++) {
            initBoxLow[k] = simBoxLow[k];
            initBoxHigh[k] = simBoxHigh[k];
            initOrient[k] = 2;
        }
        y.get("initBoxLow", initBoxLow, 3);
        y.get("initBoxHigh", initBoxHigh, 3);
        y.get("initOrient", initOrient, 3);
        initCircularX = false;
        y.get("initCircularX", initCircularX);
        initPreSteps = 100;
        y.get("initPreSteps", initPreSteps);
        thermEquilTime = 0;
        y.get("thermEquilTime", thermEquilTime);
        linkKappa = 100;
        linkGap = 0.01;
        y.get("linkKappa", linkKappa);
        y.get("linkGap", linkGap);
        sylinderFixed = false;
        y.get("sylinderFixed", sylinderFixed);
        sylinderLengthSigma = -1;
        y.get("sylinderLengthSigma", sylinderLengthSigma);
        sylinderDiameterColRatio = 1.0;
        y.get("sylinderDiameterColRatio", sylinderDiameterColRatio);
        sylinderLengthColRatio = 1.0;
        y.get("sylinderLengthColRatio", sylinderLengthColRatio);
        sylinderColBuf = 0.3;
        y.get("sylinderColBuf", sylinderColBuf);
        boundaries.clear(); // :75-89
        for (const auto &b : y.boundaries) {
            alens_boundary o{};
            const std::string type = b.str("type");
            if (type == "sphere") { // SphereShell::initialize, Boundary.cpp:18-22
                o.type = 0;
                b.get("center", o.center, 3, true);
                b.get("radius", o.radius, true);
                bool in = true;
                b.get("inside", in, true);
                o.inside = in;
            } else if (type == "wall") { // Wall::initialize, Boundary.cpp:98-101
                o.type = 1;
                b.get("center", o.center, 3, true);
                b.get("norm", o.axis, 3, true);
            } else if (type == "tube") { // Tube::initialize, Boundary.cpp:168-173
                o.type = 2;
                b.get("center", o.center, 3, true);
                b.get("axis", o.axis, 3, true);
                bool in = true;
                b.get("inside", in, true);
                o.inside = in;
                b.get("radius", o.radius, true);
            } else {
                continue;
            }
            boundaries.push_back(o);
        }
    }

    /// SylinderConfig::dump (SylinderConfig.cpp:92-)
    void dump() const {
        std::printf("-------------------------------------------\nRun Setting: \n");
        std::printf("Random number seed: %d\nLog Level: %d\nTimer Level: %d\n", rngSeed, logLevel, timerLevel);
        std::printf("Simulation box Low: %g,%g,%g\n", simBoxLow[0], simBoxLow[1], simBoxLow[2]);
        std::printf("Simulation box High: %g,%g,%g\n", simBoxHigh[0], simBoxHigh[1], simBoxHigh[2]);
        std::printf("Periodicity: %d,%d,%d\n", simBoxPBC[0], simBoxPBC[1], simBoxPBC[2]);
        std::printf("Time step size: %g\nTotal Time: %g\nSnap Time: %g\n", dt, timeTotal, timeSnap);
        std::printf("-------------------------------------------\n");
    }

  private:
    // ---- the YAML subset of RunConfig.yaml
    struct Yaml {
        std::map<std::string, std::string> kv; // key -> raw value text
        std::vector<Yaml> boundaries;
        Yaml() = default;
        explicit Yaml(const std::string &filename) {
            std::ifstream f(filename);
            if (!f) throw std::runtime_error("SylinderConfig: cannot open " + filename);
            std::string line;
            bool inBoundaries = false;
            while (std::getline(f, line)) {
                const size_t hash = findComment(line);
                if (hash != std::string::npos) line = line.substr(0, hash);
                const size_t a = line.find_first_not_of(" \t\r");
                if (a == std::string::npos) continue;
                std::string t = trim(line);
                if (t == "---" || t == "...") continue;
                if (a == 0) inBoundaries = false;
                if (inBoundaries) {
                    if (t.rfind("- ", 0) == 0) {
                        boundaries.emplace_back();
                        t = trim(t.substr(2));
                    }
                    if (boundaries.empty()) continue;
                    const size_t c = t.find(':');
                    if (c != std::string::npos) boundaries.back().kv[trim(t.substr(0, c))] = trim(t.substr(c + 1));
                    continue;
                }
                const size_t c = t.find(':');
                if (c == std::string::npos) continue;
                const std::string key = trim(t.substr(0, c)), val = trim(t.substr(c + 1));
                if (key == "boundaries" && val.empty()) inBoundaries = true;
                else kv[key] = val;
            }
        }
        static size_t findComment(const std::string &s) {
            for (size_t i = 0; i < s.size(); i++)
                if (s[i] == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return i;
            return std::string::npos;
        }
        static std::string trim(const std::string &s) {
            const size_t a = s.find_first_not_of(" \t\r\n\"'"), b = s.find_last_not_of(" \t\r\n\"'");
            return a == std::string::npos ? "" : s.substr(a, b - a + 1);
        }
        bool has(const std::string &k) const { return kv.count(k) != 0; }
        std::string str(const std::string &k) const {
            auto it = kv.find(k);
            return it == kv.end() ? "" : it->second;
        }
        template <class T>
        static T conv(const std::string &s) {
            std::istringstream is(s);
            T v{};
            is >> v;
            if (is.fail()) throw std::runtime_error("SylinderConfig: bad value '" + s + "'");
            return v;
        }
        void missing(const std::string &k, bool required) const {
            if (required) throw std::runtime_error("Required parameter " + k + " in input yaml file not found");
        }
        template <class T>
        void get(const std::string &k, T &v, bool required = false) const {
            if (!has(k)) return missing(k, required);
            v = conv<T>(str(k));
        }
        void get(const std::string &k, bool &v, bool required = false) const {
            if (!has(k)) return missing(k, required);
            const std::string s = str(k);
            v = (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on" || s == "1");
        }
        template <class T>
        void get(const std::string &k, T *v, int dim, bool required = false) const {
            if (!has(k)) return missing(k, required);
            std::string s = str(k);
            const size_t a = s.find('['), b = s.rfind(']');
            if (a == std::string::npos || b == std::string::npos) throw std::runtime_error("SylinderConfig: " + k + " is not a list");
            std::stringstream ss(s.substr(a + 1, b - a - 1));
            std::string item;
            int n = 0;
            while (std::getline(ss, item, ',')) {
                if (n < dim) {
                    Yaml one;
                    one.kv["v"] = trim(item);
                    one.get("v", v[n]);
                }
                n++;
            }
            if (n != dim) throw std::runtime_error("Expecting " + std::to_string(dim) + " elements in " + k + " in input yaml file.");
        }
    };
};

#endif
