// TpetraShim.hpp -- the small subset of Teuchos/Tpetra names that leaks through the reference's public
// signatures (SimToolbox/Trilinos/TpetraUtil.hpp:45-50: TCOMM/TMAP/TOP/TMV/TV; Teuchos::RCP), as plain
// host-side containers.  No Trilinos: vectors handed across these types are staging buffers for the C ABI,
// the arithmetic of the path itself happens on the device.
#ifndef ALENS_B200_TPETRASHIM_HPP_
#define ALENS_B200_TPETRASHIM_HPP_

#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>

namespace Teuchos {
// std::shared_ptr with the few RCP spellings the callers use
template <class T>
class RCP : public std::shared_ptr<T> {
  public:
    using std::shared_ptr<T>::shared_ptr;
    RCP() = default;
    RCP(const std::shared_ptr<T> &p) : std::shared_ptr<T>(p) {}
    template <class U>
    RCP(const RCP<U> &p) : std::shared_ptr<T>(static_cast<const std::shared_ptr<U> &>(p)) {}
    bool is_null() const { return !this->get(); }
    bool is_valid_ptr() const { return this->get() != nullptr; }
    RCP<const T> getConst() const { return RCP<const T>(std::shared_ptr<const T>(*this)); }
};
template <class T>
RCP<T> rcp(T *p) {
    return RCP<T>(std::shared_ptr<T>(p));
}
enum ETransp { NO_TRANS, TRANS, CONJ_TRANS };
} // namespace Teuchos

namespace alens_shim {

// one process per GPU; rank/size come from the launcher (torchrun / mpirun environment), no MPI calls here
class Comm {
    int rank_ = 0, size_ = 1;

  public:
    Comm() = default;
    Comm(int rank, int size) : rank_(rank), size_(size) {}
    int getRank() const { return rank_; }
    int getSize() const { return size_; }
    void barrier() const {}
};

// contiguous map: global index = offset + local (TpetraUtil.cpp:39-41 getTMAPFromLocalSize)
class Map {
    int nLocal_ = 0, offset_ = 0, nGlobal_ = 0;
    Teuchos::RCP<const Comm> comm_;

  public:
    Map(int nLocal, int offset, int nGlobal, Teuchos::RCP<const Comm> comm)
        : nLocal_(nLocal), offset_(offset), nGlobal_(nGlobal), comm_(comm) {}
    int getNodeNumElements() const { return nLocal_; }
    int getGlobalNumElements() const { return nGlobal_; }
    int getMinGlobalIndex() const { return offset_; }
    int getMaxGlobalIndex() const { return offset_ + nLocal_ - 1; }
    bool isSameAs(const Map &o) const { return nLocal_ == o.nLocal_ && offset_ == o.offset_ && nGlobal_ == o.nGlobal_; }
    Teuchos::RCP<const Comm> getComm() const { return comm_; }
};

// single-column vector with Tpetra's update/scale/dot/norm spellings.  zeroOut = false leaves the storage
// uninitialised, as Tpetra's constructor does (vectors that a download fills completely); the loops run on all cores when the
// host program is compiled with OpenMP -- at 1M rods these vectors are 48 MB each
#ifdef _OPENMP
#define ALENS_SHIM_OMP_FOR _Pragma("omp parallel for schedule(static)")
#define ALENS_SHIM_OMP_SUM _Pragma("omp parallel for schedule(static) reduction(+ : s)")
#define ALENS_SHIM_OMP_MAX _Pragma("omp parallel for schedule(static) reduction(max : m)")
#else
#define ALENS_SHIM_OMP_FOR
#define ALENS_SHIM_OMP_SUM
#define ALENS_SHIM_OMP_MAX
#endif
class Vector {
    Teuchos::RCP<const Map> map_;
    std::unique_ptr<double[]> v_;
    long long n_ = 0;

  public:
    struct View {
        double *p;
        size_t n;
        double &operator()(size_t i, size_t) const { return p[i]; }
        size_t dimension_0() const { return n; }
        size_t dimension_1() const { return 1; }
        size_t extent(int d) const { return d == 0 ? n : 1; }
    };
    struct ConstView {
        const double *p;
        size_t n;
        const double &operator()(size_t i, size_t) const { return p[i]; }
        size_t dimension_0() const { return n; }
        size_t dimension_1() const { return 1; }
    };
    Vector(const Teuchos::RCP<const Map> &map, bool zeroOut = true)
        : map_(map), v_(new double[(map ? map->getNodeNumElements() : 0) + 1]), n_(map ? map->getNodeNumElements() : 0) {
        if (zeroOut) putScalar(0.0);
    }
    Vector(const Vector &o) : map_(o.map_), v_(new double[o.n_ + 1]), n_(o.n_) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] = o.v_[i];
    }
    Vector &operator=(const Vector &o) {
        if (this != &o) {
            Vector t(o);
            std::swap(map_, t.map_);
            std::swap(v_, t.v_);
            std::swap(n_, t.n_);
        }
        return *this;
    }
    const Teuchos::RCP<const Map> &getMap() const { return map_; }
    size_t getLocalLength() const { return (size_t)n_; }
    size_t getNumVectors() const { return 1; }
    double *data() { return v_.get(); }
    const double *data() const { return v_.get(); }
    template <class Space = void>
    View getLocalView() { return View{v_.get(), (size_t)n_}; }
    template <class Space = void>
    ConstView getLocalView() const { return ConstView{v_.get(), (size_t)n_}; }
    template <class Space = void>
    void modify() {}
    void putScalar(double a) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] = a;
    }
    void scale(double a) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] *= a;
    }
    void scale(double a, const Vector &A) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] = a * A.v_[i];
    }
    void update(double a, const Vector &A, double b) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] = b * v_[i] + a * A.v_[i];
    }
    void update(double a, const Vector &A, double b, const Vector &B, double g) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] = g * v_[i] + a * A.v_[i] + b * B.v_[i];
    }
    void elementWiseMultiply(double s, const Vector &A, const Vector &B, double t) {
        ALENS_SHIM_OMP_FOR
        for (long long i = 0; i < n_; i++) v_[i] = t * v_[i] + s * A.v_[i] * B.v_[i];
    }
    double dot(const Vector &o) const {
        double s = 0;
        ALENS_SHIM_OMP_SUM
        for (long long i = 0; i < n_; i++) s += v_[i] * o.v_[i];
        return s;
    }
    double norm2() const { return std::sqrt(dot(*this)); }
    double normInf() const {
        double m = 0;
        ALENS_SHIM_OMP_MAX
        for (long long i = 0; i < n_; i++) m = std::max(m, std::fabs(v_[i]));
        return m;
    }
};

class Operator {
  public:
    virtual ~Operator() = default;
    virtual Teuchos::RCP<const Map> getDomainMap() const = 0;
    virtual Teuchos::RCP<const Map> getRangeMap() const = 0;
    virtual void apply(const Vector &X, Vector &Y, Teuchos::ETransp mode = Teuchos::NO_TRANS, double alpha = 1.0,
                       double beta = 0.0) const = 0;
    virtual bool hasTransposeApply() const { return false; }
};

// CSR matrix as a TOP (the reference's TCMAT, Tpetra::CrsMatrix<double,int,int>): host arrays that the device-backed
// BCQPSolver uploads once; apply() on host vectors is the caller's convenience (tests), not a solver path
class CrsMatrix : public Operator {
    Teuchos::RCP<const Map> map_;
    std::vector<long long> ptr_;
    std::vector<int> col_;
    std::vector<double> val_;
    struct alens_ctx *dev_ = nullptr;

  public:
    CrsMatrix(const Teuchos::RCP<const Map> &rowMap, std::vector<long long> rowPtr, std::vector<int> colInd,
              std::vector<double> values)
        : map_(rowMap), ptr_(std::move(rowPtr)), col_(std::move(colInd)), val_(std::move(values)) {}
    void setDevice(struct alens_ctx *c) { dev_ = c; }
    struct alens_ctx *device() const { return dev_; }
    size_t getNodeNumRows() const { return ptr_.size() - 1; }
    size_t getNodeNumEntries() const { return col_.size(); }
    const std::vector<long long> &rowPtr() const { return ptr_; }
    const std::vector<int> &colInd() const { return col_; }
    const std::vector<double> &values() const { return val_; }
    Teuchos::RCP<const Map> getDomainMap() const override { return map_; }
    Teuchos::RCP<const Map> getRangeMap() const override { return map_; }
    void apply(const Vector &X, Vector &Y, Teuchos::ETransp = Teuchos::NO_TRANS, double alpha = 1.0,
               double beta = 0.0) const override {
        for (size_t i = 0; i + 1 < ptr_.size(); i++) {
            double s = 0;
            for (long long k = ptr_[i]; k < ptr_[i + 1]; k++) s += val_[k] * X.data()[col_[k]];
            Y.data()[i] = (beta == 0.0 ? 0.0 : beta * Y.data()[i]) + alpha * s;
        }
    }
};

} // namespace alens_shim

using TCMAT = alens_shim::CrsMatrix;
using TCOMM = alens_shim::Comm;
using TMAP = alens_shim::Map;
using TV = alens_shim::Vector;
using TMV = alens_shim::Vector;
using TOP = alens_shim::Operator;

inline Teuchos::RCP<const TCOMM> getMPIWORLDTCOMM(int rank = 0, int size = 1) {
    return Teuchos::RCP<const TCOMM>(std::make_shared<const TCOMM>(rank, size));
}
/// contiguous map from the local size (TpetraUtil.cpp:37-41); offsets of the other ranks are supplied by the
/// caller on multi-GPU runs (exclusive scan over slab counts)
inline Teuchos::RCP<TMAP> getTMAPFromLocalSize(int localSize, const Teuchos::RCP<const TCOMM> &comm, int offset = 0,
                                               int globalSize = -1) {
    return Teuchos::RCP<TMAP>(std::make_shared<TMAP>(localSize, offset, globalSize < 0 ? localSize : globalSize, comm));
}
inline Teuchos::RCP<TV> getTVFromVector(const std::vector<double> &in, const Teuchos::RCP<const TCOMM> &comm) {
    auto map = getTMAPFromLocalSize((int)in.size(), comm);
    Teuchos::RCP<TV> v(std::make_shared<TV>(Teuchos::RCP<const TMAP>(map), false));
    std::copy(in.begin(), in.end(), v->data());
    return v;
}

#endif
