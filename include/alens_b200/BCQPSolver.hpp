// BCQPSolver.hpp -- bound-constrained QP front end mirroring SimToolbox/Constraint/BCQPSolver.hpp:37-111.
// The BBPGD / APGD loops (BCQPSolver.cpp:134-389) run on the device (alens_bcqp_solve); there is no host
// implementation, so A must be the device-backed ConstraintOperator and the bounds are the ones
// ConstraintSolver uses (0 for unilateral, -0.1*DBL_MAX for bilateral rows; +DBL_MAX/10 above).
#ifndef ALENS_B200_BCQPSOLVER_HPP_
#define ALENS_B200_BCQPSOLVER_HPP_

#include <array>
#include <deque>
#include <limits>
#include <stdexcept>
#include <vector>

#include "ConstraintOperator.hpp"

using IteHistory = std::deque<std::array<double, 6>>; ///< {ite, 0, 0, step, resPhi, mvCount}

class BCQPSolver {
    Teuchos::RCP<const TOP> ARcp;
    Teuchos::RCP<const TV> bRcp;
    Teuchos::RCP<TV> lbRcp, ubRcp;
    const ConstraintOperator *dev_ = nullptr;

    int run(Teuchos::RCP<TV> &xsolRcp, double tol, int iteMax, IteHistory &history, int choice) const {
        if (!dev_) throw std::invalid_argument("BCQPSolver: A is not a device ConstraintOperator (no CPU fallback)");
        if (!xsolRcp->getMap()->isSameAs(*bRcp->getMap()))
            throw std::invalid_argument("xsolrcp and A operator do not have the same Map."); // BCQPSolver.cpp:136-137
        alens_solve_report rep{};
        if (alens_bcqp_solve(dev_->device(), bRcp->data(), xsolRcp->data(), tol, iteMax, choice, &rep) != ALENS_OK)
            throw std::runtime_error(alens_last_error(dev_->device()));
        std::vector<double> rows(6 * (size_t)std::max(rep.history_rows, 1));
        int n = 0;
        alens_get_history(dev_->device(), rows.data(), rep.history_rows, &n);
        for (int i = 0; i < std::min(n, rep.history_rows); i++)
            history.push_back({rows[6 * i], rows[6 * i + 1], rows[6 * i + 2], rows[6 * i + 3], rows[6 * i + 4],
                               rows[6 * i + 5]});
        return rep.status;
    }

  public:
    BCQPSolver(const Teuchos::RCP<const TOP> &A_, const Teuchos::RCP<const TV> &b_) : ARcp(A_), bRcp(b_) {
        if (!ARcp->getDomainMap()->isSameAs(*bRcp->getMap()))
            throw std::invalid_argument("A (domain) and b do not have the same Map."); // BCQPSolver.cpp:14-15
        dev_ = dynamic_cast<const ConstraintOperator *>(ARcp.get());
        lbRcp = Teuchos::RCP<TV>(std::make_shared<TV>(bRcp->getMap(), false));
        ubRcp = Teuchos::RCP<TV>(std::make_shared<TV>(bRcp->getMap(), false));
        lbRcp->putScalar(-std::numeric_limits<double>::max() / 10); // setDefaultBounds, BCQPSolver.cpp:499-510
        ubRcp->putScalar(std::numeric_limits<double>::max() / 10);
    }
    /// bounds live on the device (derived from the bilateral flag); these views exist for source compatibility
    Teuchos::RCP<TV> getLowerBound() { return lbRcp; }
    Teuchos::RCP<TV> getUpperBound() { return ubRcp; }

    int solveBBPGD(Teuchos::RCP<TV> &xsolRcp, const double tol, const int iteMax, IteHistory &history) const {
        return run(xsolRcp, tol, iteMax, history, ALENS_SOLVER_BBPGD);
    }
    int solveAPGD(Teuchos::RCP<TV> &xsolRcp, const double tol, const int iteMax, IteHistory &history) const {
        return run(xsolRcp, tol, iteMax, history, ALENS_SOLVER_APGD);
    }
};

#endif
