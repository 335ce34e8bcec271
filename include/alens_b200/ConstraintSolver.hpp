// ConstraintSolver.hpp -- per-step driver mirroring SimToolbox/Constraint/ConstraintSolver.hpp:32-120 /
// .cpp:4-107: setup (q = delta0/dt + D^T v_nc, bounds), solve with BBPGD/APGD, uni/bi split, write-back.
// All arithmetic happens behind the C ABI; this class keeps the reference's call sequence and result
// vectors (ownership: result RCPs outlive reset(), as SylinderSystem.cpp:973-976 relies on).
#ifndef ALENS_B200_CONSTRAINTSOLVER_HPP_
#define ALENS_B200_CONSTRAINTSOLVER_HPP_

#include <cstdio>
#include <stdexcept>

#include "BCQPSolver.hpp"
#include "ConstraintCollector.hpp"

class ConstraintSolver {
    alens_ctx *ctx_ = nullptr;
    ConstraintCollector conCollector;
    double dt = 0;
    double res = 1e-5;
    int maxIte = 1000000;
    int solverChoice = 0;
    Teuchos::RCP<const TMAP> mobMapRcp;
    Teuchos::RCP<TV> velncRcp, forceuRcp, forcebRcp, veluRcp, velbRcp, gammaRcp;
    Teuchos::RCP<TV> zeroRcp; ///< the bilateral force / velocity of a pool without a bilateral block (never written)
    IteHistory history;
    alens_solve_report report{};

    void ck(int rc) const {
        if (rc != ALENS_OK) throw std::runtime_error(alens_last_error(ctx_));
    }

  public:
    ConstraintSolver() = default;
    explicit ConstraintSolver(alens_ctx *ctx) : ctx_(ctx) {}
    void bindDevice(alens_ctx *ctx) { ctx_ = ctx; }

    void reset() { // ConstraintSolver.cpp:36-58
        setControlParams(1e-5, 1000000, 0);
        mobMapRcp.reset(); velncRcp.reset();
        forceuRcp.reset(); forcebRcp.reset(); veluRcp.reset(); velbRcp.reset(); gammaRcp.reset();
    }

    void setControlParams(double res_, int maxIte_, int solverChoice_) {
        res = res_;
        maxIte = maxIte_;
        solverChoice = solverChoice_;
    }

    /// mobOpRcp_ is accepted for signature compatibility; the mobility lives on the device
    /// (alens_calc_mobility).  Host-generated blocks waiting in the collector's pool are appended to the
    /// device list here, in queue order.
    void setup(ConstraintCollector &conCollector_, Teuchos::RCP<TOP> & /*mobOpRcp_*/, Teuchos::RCP<TV> &velncRcp_,
               double dt_) {
        if (!ctx_) throw std::invalid_argument("ConstraintSolver: no device bound");
        reset();
        conCollector = conCollector_; // shallow copy: same pool (ConstraintSolver.cpp:8)
        dt = dt_;
        velncRcp = velncRcp_;
        const auto blocks = conCollector.flatten();
        if (!blocks.empty())
            ck(alens_append_constraints(ctx_, reinterpret_cast<const alens_constraint_block *>(blocks.data()),
                                        (long long)blocks.size()));
        ck(alens_set_velocity_noncon(ctx_, velncRcp.is_null() ? nullptr : velncRcp->data()));
        ck(alens_setup_constraints(ctx_, nullptr, dt));
        mobMapRcp = velncRcp.is_null() ? Teuchos::RCP<const TMAP>() : velncRcp->getMap();
    }

    void solveConstraints() { // ConstraintSolver.cpp:60-107
        if (!ctx_) throw std::invalid_argument("ConstraintSolver: no device bound");
        ck(alens_solve_constraints(ctx_, nullptr, dt, res, maxIte, solverChoice, &report));
        history.clear();
        std::vector<double> rows(6 * (size_t)std::max(report.history_rows, 1));
        int n = 0;
        alens_get_history(ctx_, rows.data(), report.history_rows, &n);
        for (int i = 0; i < std::min(n, report.history_rows); i++)
            history.push_back({rows[6 * i], rows[6 * i + 1], rows[6 * i + 2], rows[6 * i + 3], rows[6 * i + 4],
                               rows[6 * i + 5]});
        auto comm = getMPIWORLDTCOMM();
        Teuchos::RCP<const TMAP> mob = mobMapRcp.is_null()
                                           ? Teuchos::RCP<const TMAP>(getTMAPFromLocalSize(6 * report.n_rods, comm))
                                           : mobMapRcp;
        // without a bilateral block force_b = D gamma_b and vel_b = M force_b (ConstraintSolver.cpp:98-101) are identically
        // zero: ONE zero vector (kept across steps while the size stays) stands for both, and 96 bytes per rod stay off the
        // PCIe bus.  The vectors a download fills completely are not zeroed first (48 MB each at 1M rods).
        long long nBi = 0;
        ck(alens_get_pool_stats(ctx_, nullptr, nullptr, &nBi));
        auto mk = [&](bool zero) { return Teuchos::RCP<TV>(std::make_shared<TV>(mob, zero)); };
        forceuRcp = mk(false); veluRcp = mk(false);
        if (nBi) {
            forcebRcp = mk(false); velbRcp = mk(false);
        } else {
            if (zeroRcp.is_null() || zeroRcp->getLocalLength() != (size_t)mob->getNodeNumElements()) zeroRcp = mk(true);
            forcebRcp = zeroRcp; velbRcp = zeroRcp;
        }
        ck(alens_get_force_velocity(ctx_, forceuRcp->data(), veluRcp->data(), nBi ? forcebRcp->data() : nullptr,
                                    nBi ? velbRcp->data() : nullptr));
    }

    /// the same log line as ConstraintSolver.cpp:92-93 ("RECORD: BCQP residue ...")
    void printRecord(FILE *f = stdout) const {
        if (history.empty()) return;
        const auto &p = history.back();
        fprintf(f, "RECORD: BCQP residue %g, %g, %g, %g, %g, %g\n", p[0], p[1], p[2], p[3], p[4] * dt, p[5]);
    }

    /// gamma -> blocks, stress *= gamma (ConstraintCollector::writeBackGamma): refills the shared pool
    /// from the device.  This moves 272 B per constraint over PCIe; call it on snapshot steps only.
    void writebackGamma() { conCollector.pullFromDevice(ctx_, true, true); }

    Teuchos::RCP<const TV> getForceUni() const { return forceuRcp; }
    Teuchos::RCP<const TV> getVelocityUni() const { return veluRcp; }
    Teuchos::RCP<const TV> getForceBi() const { return forcebRcp; }
    Teuchos::RCP<const TV> getVelocityBi() const { return velbRcp; }
    const IteHistory &getHistory() const { return history; }
    const alens_solve_report &getReport() const { return report; }
    /// solved gamma in device (= pullFromDevice) order
    Teuchos::RCP<TV> getGamma() {
        auto comm = getMPIWORLDTCOMM();
        gammaRcp = Teuchos::RCP<TV>(std::make_shared<TV>(
            Teuchos::RCP<const TMAP>(getTMAPFromLocalSize((int)report.n_constraints, comm)), true));
        ck(alens_get_gamma(ctx_, gammaRcp->data(), report.n_constraints));
        return gammaRcp;
    }
};

#endif
