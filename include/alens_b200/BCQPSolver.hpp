// BCQPSolver.hpp -- bound-constrained QP front end mirroring SimToolbox/Constraint/BCQPSolver.hpp:37-111, member for
// member: both constructors, setLowerBound / setUpperBound / getLowerBound / getUpperBound, prepareSolver, solveAPGD,
// solveBBPGD, selfTest.  The loops (BCQPSolver.cpp:134-389) run on the device behind alens_bcqp_* (bcqp.cu); there is no
// host implementation.  Operators the device can apply: the matrix-free ConstraintOperator of a context, and a CSR
// matrix (the shim's TCMAT, uploaded once).  Any other TOP throws std::invalid_argument (no CPU fallback).
#ifndef ALENS_B200_BCQPSOLVER_HPP_
#define ALENS_B200_BCQPSOLVER_HPP_

#include <array>
#include <cmath>
#include <cstdio>
#include <deque>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

#include "ConstraintOperator.hpp"

using IteHistory = std::deque<std::array<double, 6>>; ///< {ite, 0, 0, step, resPhi, mvCount}

class BCQPSolver {
    Teuchos::RCP<const TOP> ARcp;
    Teuchos::RCP<const TV> bRcp;
    Teuchos::RCP<const TMAP> mapRcp;
    Teuchos::RCP<const TCOMM> commRcp;
    Teuchos::RCP<TV> lbRcp, ubRcp;
    bool lbSet = false, ubSet = false;
    alens_ctx *ctx_ = nullptr;
    mutable alens_bcqp *dev_ = nullptr;

    static void ck(alens_ctx *c, int rc) {
        if (rc == ALENS_ERR_PROJECTION) throw std::runtime_error("projection error occured"); // BCQPSolver.cpp:484-494 exits
        if (rc != ALENS_OK) throw std::runtime_error(alens_last_error(c));
    }
    void bind() { // the device object behind (A, b)
        if (const auto *op = dynamic_cast<const ConstraintOperator *>(ARcp.get())) {
            ctx_ = op->device();
            ck(ctx_, alens_bcqp_create_constraint(ctx_, bRcp->data(), &dev_));
        } else if (const auto *mat = dynamic_cast<const TCMAT *>(ARcp.get())) {
            ctx_ = mat->device();
            if (!ctx_) throw std::invalid_argument("BCQPSolver: the CrsMatrix has no device context (TCMAT::setDevice)");
            ck(ctx_, alens_bcqp_create_csr(ctx_, (int)mat->getNodeNumRows(), mat->rowPtr().data(), mat->colInd().data(),
                                           mat->values().data(), bRcp->data(), &dev_));
        } else {
            throw std::invalid_argument("BCQPSolver: this operator type cannot be applied on the device (no CPU fallback)");
        }
    }
    int run(Teuchos::RCP<TV> &xsolRcp, double tol, int iteMax, IteHistory &history, int choice) const {
        if (!mapRcp->isSameAs(*xsolRcp->getMap()))
            throw std::invalid_argument("xsolrcp and A operator do not have the same Map."); // BCQPSolver.cpp:136-137
        // the bounds the caller set (or edited in place through getLowerBound(), as ConstraintSolver.cpp:66-69 does)
        ck(ctx_, alens_bcqp_set_lower_bound(dev_, lbRcp->data()));
        ck(ctx_, alens_bcqp_set_upper_bound(dev_, ubRcp->data()));
        alens_solve_report rep{};
        Teuchos::RCP<TV> out(std::make_shared<TV>(*xsolRcp)); // the reference hands back a fresh vector
        const int rc = alens_bcqp_run(dev_, out->data(), tol, iteMax, choice, &rep);
        std::vector<double> rows(6 * (size_t)std::max(rep.history_rows, 1));
        int n = 0;
        alens_bcqp_history(dev_, rows.data(), rep.history_rows, &n);
        for (int i = 0; i < std::min(n, rep.history_rows); i++)
            history.push_back({rows[6 * i], rows[6 * i + 1], rows[6 * i + 2], rows[6 * i + 3], rows[6 * i + 4],
                               rows[6 * i + 5]});
        ck(ctx_, rc);
        xsolRcp = out;
        return rep.status;
    }
    void setDefaultBounds() { // BCQPSolver.cpp:499-510
        if (!lbSet) {
            Teuchos::RCP<TV> v(std::make_shared<TV>(bRcp->getMap(), false));
            v->putScalar(-std::numeric_limits<double>::max() / 10);
            setLowerBound(v);
        }
        if (!ubSet) {
            Teuchos::RCP<TV> v(std::make_shared<TV>(bRcp->getMap(), false));
            v->putScalar(std::numeric_limits<double>::max() / 10);
            setUpperBound(v);
        }
    }
    void generateRandomBounds(std::mt19937 &gen) { // BCQPSolver.cpp:512-533: lb = min(u, v), ub = max(u, v), u, v ~ U(-1, 1)
        std::uniform_real_distribution<> dis(-1, 1);
        Teuchos::RCP<TV> v1(std::make_shared<TV>(bRcp->getMap(), true)), v2(std::make_shared<TV>(bRcp->getMap(), true));
        for (size_t i = 0; i < v1->getLocalLength(); i++) {
            const double a = dis(gen), b = dis(gen);
            v1->data()[i] = std::min(a, b);
            v2->data()[i] = std::max(a, b);
        }
        setLowerBound(v1);
        setUpperBound(v2);
    }

  public:
    BCQPSolver(const Teuchos::RCP<const TOP> &A_, const Teuchos::RCP<const TV> &b_)
        : ARcp(A_), bRcp(b_), mapRcp(b_->getMap()), commRcp(b_->getMap()->getComm()) {
        if (!ARcp->getDomainMap()->isSameAs(*bRcp->getMap()))
            throw std::invalid_argument("A (domain) and b do not have the same Map."); // BCQPSolver.cpp:14-15
        bind();
        setDefaultBounds();
    }
    /// internal test problem (BCQPSolver.cpp:23-132): A = B^T D B + diagonal I with B ~ U(-1,1)^(n x n), D = diag(10^U(-1,1)),
    /// entries below 1e-7 dropped; b ~ U(-1,1); random bounds.  `ctx` is the device the problem is solved on; `seed` makes
    /// the problem reproducible (the reference seeds from std::random_device).
    BCQPSolver(int localSize, double diagonal, alens_ctx *ctx, unsigned seed = std::random_device{}()) : ctx_(ctx) {
        commRcp = getMPIWORLDTCOMM();
        Teuchos::RCP<TMAP> rowMap = getTMAPFromLocalSize(localSize, commRcp);
        mapRcp = rowMap;
        std::mt19937 gen(seed);
        std::uniform_real_distribution<> dis(-1, 1);
        Teuchos::RCP<TV> btemp(std::make_shared<TV>(mapRcp, false));
        for (int i = 0; i < localSize; i++) btemp->data()[i] = dis(gen);
        bRcp = btemp;
        const size_t n = (size_t)localSize;
        std::vector<double> B(n * n), D(n), A(n * n, 0.0);
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < n; j++) B[i * n + j] = dis(gen);
        for (size_t i = 0; i < n; i++) D[i] = std::pow(10, dis(gen));
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < n; j++) {
                double s = 0;
                for (size_t k = 0; k < n; k++) s += B[k * n + i] * (D[k] * B[k * n + j]);
                A[i * n + j] = s + (i == j ? diagonal : 0.0);
            }
        std::vector<long long> rowPtr(n + 1, 0);
        std::vector<int> col;
        std::vector<double> val;
        for (size_t i = 0; i < n; i++) {
            for (size_t j = 0; j < n; j++)
                if (std::fabs(A[i * n + j]) > 1e-7) {
                    col.push_back((int)j);
                    val.push_back(A[i * n + j]);
                }
            rowPtr[i + 1] = (long long)col.size();
        }
        auto mat = std::make_shared<TCMAT>(mapRcp, rowPtr, col, val);
        mat->setDevice(ctx);
        ARcp = Teuchos::RCP<const TOP>(std::shared_ptr<const TOP>(mat));
        bind();
        generateRandomBounds(gen);
    }
    ~BCQPSolver() {
        if (dev_) alens_bcqp_destroy(dev_);
    }
    BCQPSolver(const BCQPSolver &) = delete;
    BCQPSolver &operator=(const BCQPSolver &) = delete;

    void setLowerBound(const Teuchos::RCP<TV> &lbRcp_) {
        lbSet = true;
        lbRcp = lbRcp_;
    }
    void setUpperBound(const Teuchos::RCP<TV> &ubRcp_) {
        ubSet = true;
        ubRcp = ubRcp_;
    }
    Teuchos::RCP<TV> getLowerBound() { return lbRcp; }
    Teuchos::RCP<TV> getUpperBound() { return ubRcp; }
    void prepareSolver() { setDefaultBounds(); }
    Teuchos::RCP<const TOP> getOperator() const { return ARcp; }
    Teuchos::RCP<const TV> getB() const { return bRcp; }

    int solveBBPGD(Teuchos::RCP<TV> &xsolRcp, const double tol, const int iteMax, IteHistory &history) const {
        return run(xsolRcp, tol, iteMax, history, ALENS_SOLVER_BBPGD);
    }
    int solveAPGD(Teuchos::RCP<TV> &xsolRcp, const double tol, const int iteMax, IteHistory &history) const {
        return run(xsolRcp, tol, iteMax, history, ALENS_SOLVER_APGD);
    }
    /// BCQPSolver.cpp:391-429: zero initial guess, default bounds where none are set, history printed as CSV lines
    int selfTest(double tol, int maxIte, int solverChoice, Teuchos::RCP<TV> *solution = nullptr) {
        IteHistory history;
        Teuchos::RCP<TV> xsolRcp(std::make_shared<TV>(mapRcp, true));
        prepareSolver();
        if (solverChoice == 1) solveAPGD(xsolRcp, tol, maxIte, history);
        else solveBBPGD(xsolRcp, tol, maxIte, history);
        if (commRcp->getRank() == 0)
            for (const auto &record : history) {
                std::printf(solverChoice == 1 ? "APGD_HISTORY," : "BBPGD_HISTORY,");
                for (const auto &v : record) std::printf("%.6g, ", v);
                std::printf("\n");
            }
        if (solution) *solution = xsolRcp;
        return 0;
    }
};

#endif
