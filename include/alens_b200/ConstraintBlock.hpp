// ConstraintBlock.hpp -- the constraint record shared with host code (boundary / link / protein
// producers push these into the pool; snapshots read them back).
// Same fields, order and 272-byte layout as SimToolbox/Constraint/ConstraintBlock.hpp:30-127 so that
// a pool can be handed to alens_append_constraints / alens_get_constraints without conversion.
#ifndef ALENS_B200_CONSTRAINTBLOCK_HPP_
#define ALENS_B200_CONSTRAINTBLOCK_HPP_

#include <algorithm>
#include <deque>
#include <type_traits>
#include <utility>
#include <vector>

#include "../alens_b200.h"

#ifndef GEO_INVALID_INDEX
#define GEO_INVALID_INDEX (-1) // Util/GeoCommon.h:14
#endif

struct ConstraintBlock {
    double delta0 = 0;  ///< constraint initial value
    double gamma = 0;   ///< force magnitude / initial guess
    double gammaLB = 0; ///< stored, never read by the solver (bounds come from `bilateral`)
    int gidI = GEO_INVALID_INDEX, gidJ = GEO_INVALID_INDEX;
    int globalIndexI = GEO_INVALID_INDEX, globalIndexJ = GEO_INVALID_INDEX;
    bool oneSide = false;   ///< body J does not appear in the mobility matrix
    bool bilateral = false; ///< unbounded gamma
    double kappa = 0;       ///< spring constant, 0 = rigid
    double normI[3] = {0, 0, 0}, normJ[3] = {0, 0, 0};
    double posI[3] = {0, 0, 0}, posJ[3] = {0, 0, 0};
    double labI[3] = {0, 0, 0}, labJ[3] = {0, 0, 0};
    double stress[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; ///< row-major, for unit gamma

    ConstraintBlock() = default;
    ConstraintBlock(double delta0_, double gamma_, int gidI_, int gidJ_, int globalIndexI_, int globalIndexJ_,
                    const double normI_[3], const double normJ_[3], const double posI_[3], const double posJ_[3],
                    const double labI_[3], const double labJ_[3], bool oneSide_, bool bilateral_, double kappa_,
                    double gammaLB_)
        : delta0(delta0_), gamma(gamma_), gammaLB(gammaLB_), gidI(gidI_), gidJ(gidJ_), globalIndexI(globalIndexI_),
          globalIndexJ(globalIndexJ_), oneSide(oneSide_), bilateral(bilateral_), kappa(kappa_) {
        for (int d = 0; d < 3; d++) {
            normI[d] = normI_[d]; normJ[d] = normJ_[d];
            posI[d] = posI_[d];   posJ[d] = posJ_[d];
            labI[d] = labI_[d];   labJ[d] = labJ_[d];
        }
    }

    void setStress(const double *s) { std::copy(s, s + 9, stress); }
    /// any 3x3 type with operator()(i,j), e.g. Eigen::Matrix3d
    template <class Mat3, class = decltype(std::declval<const Mat3 &>()(0, 0))>
    void setStress(const Mat3 &m) {
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) stress[i * 3 + j] = m(i, j);
    }
    const double *getStress() const { return stress; }
    template <class Mat3, class = decltype(std::declval<Mat3 &>()(0, 0))>
    void getStress(Mat3 &m) const {
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) m(i, j) = stress[i * 3 + j];
    }
    void reverseIJ() {
        std::swap(gidI, gidJ);
        std::swap(globalIndexI, globalIndexJ);
        for (int k = 0; k < 3; k++) {
            std::swap(normI[k], normJ[k]);
            std::swap(posI[k], posJ[k]);
            std::swap(labI[k], labJ[k]);
        }
    }
};

static_assert(std::is_trivially_copyable<ConstraintBlock>::value, "");
static_assert(sizeof(ConstraintBlock) == sizeof(alens_constraint_block), "layout must match the C ABI record");

using ConstraintBlockQue = std::deque<ConstraintBlock>;      ///< blocks collected by one thread
using ConstraintBlockPool = std::vector<ConstraintBlockQue>; ///< one queue per OpenMP thread

#endif
