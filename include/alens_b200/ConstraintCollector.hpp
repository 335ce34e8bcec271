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
#include "VtkPolyWriter.hpp"

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

    /// ConBlock_r<rank>_<postfix>.vtp: one line per block from labI to labJ (ConstraintCollector.cpp:103-224)
    void writeVTP(const std::string &folder, const std::string &prefix, const std::string &postfix, int rank) const {
        using alens_vtk::Column;
        const std::vector<ConstraintBlock> b = flatten();
        const size_t n = b.size();
        std::vector<double> ends(6 * n);
        std::vector<int32_t> gid(2 * n), gidx(2 * n), one(n), bi(n);
        std::vector<float> posIJ(6 * n), normIJ(6 * n), d0(n), gam(n), kap(n), stress(9 * n);
        for (size_t i = 0; i < n; i++) {
            for (int k = 0; k < 3; k++) {
                ends[6 * i + k] = b[i].labI[k];         ends[6 * i + 3 + k] = b[i].labJ[k];
                posIJ[6 * i + k] = (float)b[i].posI[k];   posIJ[6 * i + 3 + k] = (float)b[i].posJ[k];
                normIJ[6 * i + k] = (float)b[i].normI[k]; normIJ[6 * i + 3 + k] = (float)b[i].normJ[k];
            }
            gid[2 * i] = b[i].gidI; gid[2 * i + 1] = b[i].gidJ;
            gidx[2 * i] = b[i].globalIndexI; gidx[2 * i + 1] = b[i].globalIndexJ;
            one[i] = b[i].oneSide ? 1 : 0;
            bi[i] = b[i].bilateral ? 1 : 0;
            d0[i] = (float)b[i].delta0; gam[i] = (float)b[i].gamma; kap[i] = (float)b[i].kappa;
            for (int k = 0; k < 9; k++) stress[9 * i + k] = (float)b[i].stress[k];
        }
        alens_vtk::writeLinePiece(folder + '/' + prefix + "ConBlock_r" + std::to_string(rank) + "_" + postfix + ".vtp", (int)n, ends,
                                  {Column::of("gid", 1, gid), Column::of("globalIndex", 1, gidx), Column::of("posIJ", 3, posIJ),
                                   Column::of("normIJ", 3, normIJ)},
                                  {Column::of("oneSide", 1, one), Column::of("bilateral", 1, bi), Column::of("delta0", 1, d0),
                                   Column::of("gamma", 1, gam), Column::of("kappa", 1, kap), Column::of("Stress", 9, stress)});
    }
    /// ConBlock_<postfix>.pvtp (ConstraintCollector.cpp:76-101)
    void writePVTP(const std::string &folder, const std::string &prefix, const std::string &postfix, const int nProcs) const {
        std::vector<std::string> pieces;
        for (int i = 0; i < nProcs; i++) pieces.push_back(prefix + "ConBlock_r" + std::to_string(i) + "_" + postfix + ".vtp");
        alens_vtk::writeParallelIndex(folder + "/" + prefix + "ConBlock_" + postfix + ".pvtp",
                                      {{"gid", "Int32", 1}, {"globalIndex", "Int32", 1}, {"posIJ", "Float32", 3}, {"normIJ", "Float32", 3}},
                                      {{"oneSide", "Int32", 1}, {"bilateral", "Int32", 1}, {"delta0", "Float32", 1},
                                       {"gamma", "Float32", 1}, {"kappa", "Float32", 1}, {"Stress", "Float32", 9}},
                                      pieces);
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
