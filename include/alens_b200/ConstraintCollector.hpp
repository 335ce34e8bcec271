// ConstraintCollector.hpp -- owner of the host-side constraint pool, mirroring
// SimToolbox/Constraint/ConstraintCollector.hpp:34-167 / .cpp:6-74,425-461.
//
// Division of labour with the device: pair-collision blocks are produced on the GPU and stay there; the
// host pool only ever holds what host code pushes into it (boundary / link / protein blocks,
// SylinderSystem.cpp:1093-1150,1386-1482, SRC/TubuleSystem.cpp:694-745).  ConstraintSolver::setup appends
// those to the device list; `pullFromDevice` refills the pool with EVERY block (gamma written back, stress
// scaled) for output code such as writeVTP / calcConStress.  The CRS build of
// buildConstraintMatrixVector (.cpp:237-423) has no equivalent: D is never materialised.
#ifndef ALENS_B200_CONSTRAINTCOLLECTOR_HPP_
#define ALENS_B200_CONSTRAINTCOLLECTOR_HPP_

#include <cassert>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ConstraintBlock.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

class ConstraintCollector {
  public:
    std::shared_ptr<ConstraintBlockPool> constraintPoolPtr; ///< all copies of a collector share this pool

    ConstraintCollector() {
        constraintPoolPtr = std::make_shared<ConstraintBlockPool>();
        constraintPoolPtr->resize(maxThreads());
    }
    ConstraintCollector(const ConstraintCollector &) = default;
    ConstraintCollector &operator=(const ConstraintCollector &) = default;

    static int maxThreads() {
#ifdef _OPENMP
        return omp_get_max_threads();
#else
        return 1;
#endif
    }

    bool valid() const { return constraintPoolPtr->empty(); } // sic, ConstraintCollector.cpp:17

    void clear() {
        assert(constraintPoolPtr);
        for (auto &q : *constraintPoolPtr) q.clear();
        constraintPoolPtr->resize(maxThreads());
    }

    int getLocalNumberOfConstraints() const {
        int sum = 0;
        for (auto &q : *constraintPoolPtr) sum += (int)q.size();
        return sum;
    }

    /// row-major 3x3 sums of the unilateral / bilateral block stresses (.cpp:38-74)
    void sumLocalConstraintStress(double uniStress[9], double biStress[9], bool withOneSide = false) const {
        for (int k = 0; k < 9; k++) uniStress[k] = biStress[k] = 0;
        for (auto &q : *constraintPoolPtr)
            for (auto &b : q) {
                if (b.oneSide && !withOneSide) continue;
                double *dst = b.bilateral ? biStress : uniStress;
                for (int k = 0; k < 9; k++) dst[k] += b.stress[k];
            }
    }

    int buildConIndex(std::vector<int> &cQueSize, std::vector<int> &cQueIndex) const {
        const auto &pool = *constraintPoolPtr;
        const int nq = (int)pool.size();
        cQueSize.assign(nq, 0);
        cQueIndex.assign(nq + 1, 0);
        for (int i = 0; i < nq; i++) cQueSize[i] = (int)pool[i].size();
        for (int i = 1; i <= nq; i++) cQueIndex[i] = cQueSize[i - 1] + cQueIndex[i - 1];
        return 0;
    }

    /// flatten the pool in queue order (the row order the reference gives D^T, .cpp:268-278)
    std::vector<ConstraintBlock> flatten() const {
        std::vector<ConstraintBlock> out;
        out.reserve(getLocalNumberOfConstraints());
        for (auto &q : *constraintPoolPtr) out.insert(out.end(), q.begin(), q.end());
        return out;
    }

    /// replace the pool content by every block the device holds (collision blocks first, then the host
    /// blocks in the order they were appended); gamma written back and stress scaled (.cpp:439-461)
    void pullFromDevice(alens_ctx *ctx, bool withStress = true, bool writeBack = true) {
        long long n = 0;
        if (alens_num_constraints(ctx, &n) != ALENS_OK) throw std::runtime_error(alens_last_error(ctx));
        std::vector<ConstraintBlock> buf((size_t)n);
        if (n > 0 && alens_get_constraints(ctx, reinterpret_cast<alens_constraint_block *>(buf.data()), n,
                                           withStress ? 1 : 0, writeBack ? 1 : 0) != ALENS_OK)
            throw std::runtime_error(alens_last_error(ctx));
        clear();
        auto &pool = *constraintPoolPtr;
        const size_t nq = pool.size();
        for (size_t i = 0; i < buf.size(); i++) pool[i * nq / std::max<size_t>(buf.size(), 1)].push_back(buf[i]);
    }
};

#endif
