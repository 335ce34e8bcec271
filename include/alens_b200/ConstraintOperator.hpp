// ConstraintOperator.hpp -- y = (D^T M D + K^-1) x as a TOP, mirroring
// SimToolbox/Constraint/ConstraintOperator.hpp / .cpp:4-71.  The explicit transpose and the three SpMVs
// are replaced by the device kernels behind alens_operator_apply; this class only stages host vectors.
#ifndef ALENS_B200_CONSTRAINTOPERATOR_HPP_
#define ALENS_B200_CONSTRAINTOPERATOR_HPP_

#include <stdexcept>

#include "../alens_b200.h"
#include "TpetraShim.hpp"

class ConstraintOperator : public TOP {
    alens_ctx *ctx_;
    Teuchos::RCP<const TMAP> gammaMapRcp, mobMapRcp;
    mutable Teuchos::RCP<TV> forceRcp, velRcp; ///< D x and M D x of the last apply (ConstraintOperator.cpp:26-27)

  public:
    ConstraintOperator(alens_ctx *ctx, const Teuchos::RCP<const TMAP> &gammaMap, const Teuchos::RCP<const TMAP> &mobMap)
        : ctx_(ctx), gammaMapRcp(gammaMap), mobMapRcp(mobMap) {
        forceRcp = Teuchos::RCP<TV>(std::make_shared<TV>(mobMapRcp, true));
        velRcp = Teuchos::RCP<TV>(std::make_shared<TV>(mobMapRcp, true));
    }
    alens_ctx *device() const { return ctx_; }

    void apply(const TMV &X, TMV &Y, Teuchos::ETransp /*mode*/ = Teuchos::NO_TRANS, double alpha = 1.0,
               double beta = 0.0) const override {
        if (!X.getMap()->isSameAs(*Y.getMap()))
            throw std::invalid_argument("X and Y do not have the same Map.\n"); // ConstraintOperator.cpp:34-35
        TV tmp(Y.getMap(), true);
        if (alens_operator_apply(ctx_, X.data(), tmp.data(), forceRcp->data(), velRcp->data()) != ALENS_OK)
            throw std::runtime_error(alens_last_error(ctx_));
        Y.update(alpha, tmp, beta);
    }
    Teuchos::RCP<const TMAP> getDomainMap() const override { return gammaMapRcp; }
    Teuchos::RCP<const TMAP> getRangeMap() const override { return gammaMapRcp; }
    Teuchos::RCP<TV> getForce() const { return forceRcp; }
    Teuchos::RCP<TV> getVel() const { return velRcp; }
    void enableTimer() {}
    void disableTimer() {}
};

#endif
