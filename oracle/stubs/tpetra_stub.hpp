// oracle/stubs/tpetra_stub.hpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A single-process, host-only stand-in for the subset of Trilinos 12.18.1 (Teuchos / Kokkos / Tpetra / Belos
// names) that the reference's constraint path uses, so that the reference's OWN sources
//     SimToolbox/Constraint/{BCQPSolver,ConstraintOperator,ConstraintSolver,ConstraintCollector}.cpp
//     SimToolbox/Trilinos/TpetraUtil.cpp
// compile UNMODIFIED, in place, into oracle/_ref/libalens_refsolver.so (recipe: oracle/Makefile `refsolver`).
// Trilinos itself is absent from this image.  Semantics follow SURVEY.md Appendix A:
//   Y.update(a,A,b)          Y = b*Y + a*A                    (b == 0: Y is overwritten, Tpetra's rule)
//   Y.update(a,A,b,B,g)      Y = g*Y + a*A + b*B              (g == 0: Y = a*A + b*B, KokkosBlas::update)
//   dot / norm2              chunks of 4096 entries summed left to right, then the chunk sums in order (Kokkos' order
//                            is thread-count dependent: the reference itself is reproducible only to rounding)
//   CrsMatrix::apply         y_i = beta*y_i + alpha*sum_j a_ij x_j, row entries summed in storage order
//   RowMatrixTransposer      explicit transpose, rows sorted by column index (Tpetra's default sort = true)
//   Map(INVALID, n, 0, comm) contiguous, global index = local index on the single rank
// One rank only: every communication call is the identity.
#pragma once
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <mpi.h>

// ---------------------------------------------------------------------------------------------- Teuchos
#define TEUCHOS_TEST_FOR_EXCEPTION(cond, Exc, msg)                                                                  \
    do {                                                                                                            \
        if (cond) {                                                                                                 \
            std::ostringstream os__;                                                                                \
            os__ << msg;                                                                                            \
            throw Exc(os__.str());                                                                                  \
        }                                                                                                           \
    } while (0);
#define TEUCHOS_ASSERT(cond) TEUCHOS_TEST_FOR_EXCEPTION(!(cond), std::logic_error, "TEUCHOS_ASSERT(" #cond ") failed")

namespace Teuchos {

enum ENull { null };
enum DataAccess { Copy, View };
enum ETransp { NO_TRANS, TRANS, CONJ_TRANS };

template <class T>
class RCP {
  public:
    std::shared_ptr<T> p;
    RCP() {}
    RCP(ENull) {}
    explicit RCP(T *raw) : p(raw) {}
    RCP(const std::shared_ptr<T> &sp) : p(sp) {}
    template <class U, class = typename std::enable_if<std::is_convertible<U *, T *>::value>::type>
    RCP(const RCP<U> &o) : p(o.p) {}
    T *get() const { return p.get(); }
    T *getRawPtr() const { return p.get(); }
    T *operator->() const { return p.get(); }
    T &operator*() const { return *p; }
    bool is_null() const { return !p; }
    bool is_valid_ptr() const { return (bool)p; }
    void reset() { p.reset(); }
    void swap(RCP<T> &o) { p.swap(o.p); }
    RCP<const T> getConst() const { return RCP<const T>(std::shared_ptr<const T>(p)); }
    RCP<T> &operator=(ENull) {
        p.reset();
        return *this;
    }
    bool operator==(ENull) const { return !p; }
    bool operator!=(ENull) const { return (bool)p; }
};
template <class T>
RCP<T> rcp(T *raw) {
    return RCP<T>(raw);
}
template <class T, class U>
RCP<T> rcp_dynamic_cast(const RCP<U> &o, bool throwOnFail = false) {
    std::shared_ptr<T> q = std::dynamic_pointer_cast<T>(o.p);
    if (!q && throwOnFail) throw std::bad_cast();
    return RCP<T>(q);
}
template <class T, class U>
RCP<T> rcp_const_cast(const RCP<U> &o) {
    return RCP<T>(std::const_pointer_cast<T>(o.p));
}

template <class T>
struct ScalarTraits {
    static T one() { return T(1); }
    static T zero() { return T(0); }
    static T eps() { return std::numeric_limits<T>::epsilon(); }
};
template <class T>
struct OrdinalTraits {
    static T invalid() { return std::numeric_limits<T>::max(); }
};
template <>
struct OrdinalTraits<int> {
    static int invalid() { return -1; }
};

template <class Ordinal>
class Comm {
  public:
    virtual ~Comm() {}
    virtual int getRank() const { return 0; }
    virtual int getSize() const { return 1; }
    virtual void barrier() const {}
};
template <class Ordinal>
class MpiComm : public Comm<Ordinal> {
  public:
    explicit MpiComm(MPI_Comm) {}
};
template <class Ordinal>
class SerialComm : public Comm<Ordinal> {};

template <class Ordinal, class T>
struct SumValueReductionOp {};
template <class Ordinal, class T>
struct MaxValueReductionOp {};
template <class Ordinal, class T>
struct MinValueReductionOp {};
enum EReductionType { REDUCE_SUM, REDUCE_MIN, REDUCE_MAX };
template <class Ordinal, class Op, class T>
void reduceAll(const Comm<Ordinal> &, const Op &, int n, const T *in, T *out) {
    for (int i = 0; i < n; i++) out[i] = in[i];
}
template <class Ordinal, class T>
void reduceAll(const Comm<Ordinal> &, EReductionType, int n, const T *in, T *out) {
    for (int i = 0; i < n; i++) out[i] = in[i];
}

class Time {
  public:
    explicit Time(const std::string &n) : name_(n) {}
    void enable() { enabled_ = true; }
    void disable() { enabled_ = false; }
    void start() { t0_ = std::chrono::steady_clock::now(); }
    void stop() {
        if (enabled_) total_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
        calls_++;
    }
    double totalElapsedTime() const { return total_; }
    void reset() { total_ = 0; calls_ = 0; }
    const std::string &name() const { return name_; }
    int numCalls() const { return calls_; }

  private:
    std::string name_;
    bool enabled_ = true;
    double total_ = 0;
    int calls_ = 0;
    std::chrono::steady_clock::time_point t0_;
};
class TimeMonitor {
  public:
    explicit TimeMonitor(Time &t) : t_(t) { t_.start(); }
    ~TimeMonitor() { t_.stop(); }
    static std::map<std::string, RCP<Time>> &table() {
        static std::map<std::string, RCP<Time>> tbl;
        return tbl;
    }
    static RCP<Time> getNewCounter(const std::string &name) {
        auto &tbl = table();
        auto it = tbl.find(name);
        if (it != tbl.end()) return it->second;
        RCP<Time> t = rcp(new Time(name));
        tbl[name] = t;
        return t;
    }
    static RCP<Time> getNewTimer(const std::string &name) { return getNewCounter(name); }
    static void summarize(std::ostream &os = std::cout) {
        for (auto &kv : table())
            os << std::left << std::setw(48) << kv.first << " " << kv.second->totalElapsedTime() << " s ("
               << kv.second->numCalls() << ")\n";
    }
    static void zeroOutTimers() {
        for (auto &kv : table()) kv.second->reset();
    }

  private:
    Time &t_;
};

template <class Ordinal, class Scalar>
class SerialDenseMatrix { // column-major like the original; only what BCQPSolver's random-problem ctor uses
  public:
    SerialDenseMatrix(Ordinal r, Ordinal c, bool zeroOut = true) : r_(r), c_(c), v_((size_t)r * c, Scalar(0)) { (void)zeroOut; }
    Scalar &operator()(Ordinal i, Ordinal j) { return v_[(size_t)j * r_ + i]; }
    const Scalar &operator()(Ordinal i, Ordinal j) const { return v_[(size_t)j * r_ + i]; }
    Ordinal numRows() const { return r_; }
    Ordinal numCols() const { return c_; }
    // this = alpha * op(A) * op(B) + beta * this   (reference BLAS dgemm loop order: j, l, i)
    int multiply(ETransp ta, ETransp tb, Scalar alpha, const SerialDenseMatrix &A, const SerialDenseMatrix &B, Scalar beta) {
        const Ordinal m = r_, n = c_, k = (ta == NO_TRANS) ? A.c_ : A.r_;
        for (Ordinal j = 0; j < n; j++)
            for (Ordinal i = 0; i < m; i++) {
                Scalar s = 0;
                for (Ordinal l = 0; l < k; l++) {
                    const Scalar a = (ta == NO_TRANS) ? A(i, l) : A(l, i);
                    const Scalar b = (tb == NO_TRANS) ? B(l, j) : B(j, l);
                    s += a * b;
                }
                (*this)(i, j) = (beta == Scalar(0) ? Scalar(0) : beta * (*this)(i, j)) + alpha * s;
            }
        return 0;
    }

  private:
    Ordinal r_, c_;
    std::vector<Scalar> v_;
};

template <class T>
class ArrayView {
  public:
    ArrayView(const T *p, size_t n) : p_(p), n_(n) {}
    const T &operator[](size_t i) const { return p_[i]; }
    size_t size() const { return n_; }

  private:
    const T *p_;
    size_t n_;
};

class GlobalMPISession {
  public:
    GlobalMPISession(int *, char ***, std::ostream * = nullptr) {}
};
class oblackholestream : public std::ostream {
  public:
    oblackholestream() : std::ostream(nullptr) {}
};

} // namespace Teuchos

// ---------------------------------------------------------------------------------------------- Kokkos
namespace Kokkos {
struct HostSpace {};
struct LayoutLeft {};

template <class DataType, class... Props>
class View;

template <class T, class... Props>
class View<T *, Props...> { // rank-1, reference counted like the original
  public:
    View() {}
    View(const std::string &, size_t n) : d_(std::make_shared<std::vector<T>>(n, T())) {}
    T &operator[](size_t i) const { return (*d_)[i]; }
    T &operator()(size_t i) const { return (*d_)[i]; }
    size_t dimension_0() const { return d_ ? d_->size() : 0; }
    size_t extent(int) const { return dimension_0(); }
    size_t size() const { return dimension_0(); }
    T *data() const { return d_ ? d_->data() : nullptr; }

  private:
    std::shared_ptr<std::vector<T>> d_;
};

template <class T>
class View2D { // what MultiVector::getLocalView<HostSpace>() returns: (row, column), one column here
  public:
    View2D(T *p, size_t n) : p_(p), n_(n) {}
    T &operator()(size_t i, size_t) const { return p_[i]; }
    size_t dimension_0() const { return n_; }
    size_t dimension_1() const { return 1; }
    size_t extent(int k) const { return k == 0 ? n_ : 1; }
    T *data() const { return p_; }

  private:
    T *p_;
    size_t n_;
};
} // namespace Kokkos

// ---------------------------------------------------------------------------------------------- Tpetra
namespace Tpetra {
typedef size_t global_size_t;
enum LocalGlobal { LocallyReplicated, GloballyDistributed };
enum CombineMode { ADD, INSERT, REPLACE, ABSMAX, ZERO };
inline std::string version() { return "Tpetra stub (oracle/stubs/tpetra_stub.hpp), single rank"; }

template <class LO = int, class GO = int, class Node = void>
class Map {
  public:
    typedef Teuchos::RCP<const Teuchos::Comm<int>> CommRcp;
    // contiguous: Map(INVALID or global size, localSize, indexBase, comm)
    Map(global_size_t, size_t localSize, GO indexBase, const CommRcp &comm)
        : comm_(comm), contiguous_(true), n_(localSize), base_(indexBase) {}
    // arbitrary list of global indices
    Map(global_size_t, const GO *ids, size_t n, GO indexBase, const CommRcp &comm)
        : comm_(comm), contiguous_(false), n_(n), base_(indexBase), ids_(ids, ids + n) {
        contiguous_ = true;
        for (size_t i = 0; i < n; i++) {
            if (ids_[i] != ids_[0] + (GO)i) contiguous_ = false;
            lid_[ids_[i]] = (LO)i;
        }
        if (contiguous_ && n > 0) base_ = ids_[0];
    }
    // uniform contiguous of a global size (single rank: everything local)
    Map(global_size_t globalSize, GO indexBase, const CommRcp &comm, LocalGlobal = GloballyDistributed)
        : comm_(comm), contiguous_(true), n_(globalSize), base_(indexBase) {}

    size_t getNodeNumElements() const { return n_; }
    global_size_t getGlobalNumElements() const { return n_; }
    GO getMinGlobalIndex() const { return n_ == 0 ? std::numeric_limits<GO>::max() : (ids_.empty() ? base_ : *std::min_element(ids_.begin(), ids_.end())); }
    GO getMaxGlobalIndex() const { return n_ == 0 ? std::numeric_limits<GO>::lowest() : (ids_.empty() ? base_ + (GO)n_ - 1 : *std::max_element(ids_.begin(), ids_.end())); }
    GO getMinAllGlobalIndex() const { return getMinGlobalIndex(); }
    GO getMaxAllGlobalIndex() const { return getMaxGlobalIndex(); }
    GO getGlobalElement(LO l) const { return ids_.empty() ? base_ + (GO)l : ids_[l]; }
    LO getLocalElement(GO g) const {
        if (ids_.empty()) return (g >= base_ && g < base_ + (GO)n_) ? (LO)(g - base_) : Teuchos::OrdinalTraits<LO>::invalid();
        auto it = lid_.find(g);
        return it == lid_.end() ? Teuchos::OrdinalTraits<LO>::invalid() : it->second;
    }
    bool isContiguous() const { return contiguous_; }
    bool isSameAs(const Map &o) const {
        if (n_ != o.n_) return false;
        for (size_t i = 0; i < n_; i++)
            if (getGlobalElement((LO)i) != o.getGlobalElement((LO)i)) return false;
        return true;
    }
    CommRcp getComm() const { return comm_; }
    struct IdList {
        const Map *m;
        GO operator[](size_t i) const { return m->getGlobalElement((LO)i); }
        size_t size() const { return m->n_; }
    };
    IdList getMyGlobalIndices() const { return IdList{this}; }
    std::string description() const {
        std::ostringstream os;
        os << "Tpetra::Map(stub){local = global = " << n_ << ", contiguous = " << contiguous_ << "}";
        return os.str();
    }

  private:
    CommRcp comm_;
    bool contiguous_;
    size_t n_;
    GO base_;
    std::vector<GO> ids_;
    std::map<GO, LO> lid_;
};

template <class LO, class GO, class Node = void>
class Import {
  public:
    template <class A, class B>
    Import(const A &, const B &) {}
};

template <class S, class LO, class GO, class Node>
class Vector;

template <class S = double, class LO = int, class GO = int, class Node = void>
class MultiVector {
  public:
    typedef S scalar_type;
    typedef LO local_ordinal_type;
    typedef GO global_ordinal_type;
    typedef Node node_type;
    typedef Map<LO, GO, Node> map_type;
    typedef Vector<S, LO, GO, Node> vec_type;

    MultiVector(const Teuchos::RCP<const map_type> &map, bool zeroOut = true)
        : map_(map), store_(std::make_shared<std::vector<S>>(map->getNodeNumElements(), S(0))), off_(0),
          n_(map->getNodeNumElements()) {
        (void)zeroOut;
    }
    MultiVector(const MultiVector &src, Teuchos::DataAccess acc) : map_(src.map_), off_(0), n_(src.n_) {
        if (acc == Teuchos::Copy) store_ = std::make_shared<std::vector<S>>(src.ptr(), src.ptr() + src.n_);
        else { store_ = src.store_; off_ = src.off_; }
    }
    // view of a range of another vector's storage (offsetViewNonConst)
    MultiVector(const Teuchos::RCP<const map_type> &map, const std::shared_ptr<std::vector<S>> &store, size_t off)
        : map_(map), store_(store), off_(off), n_(map->getNodeNumElements()) {}
    virtual ~MultiVector() {}

    Teuchos::RCP<const map_type> getMap() const { return map_; }
    size_t getLocalLength() const { return n_; }
    global_size_t getGlobalLength() const { return n_; }
    size_t getNumVectors() const { return 1; }
    S *ptr() const { return store_->data() + off_; }

    Teuchos::RCP<const vec_type> getVector(size_t) const;
    Teuchos::RCP<vec_type> getVectorNonConst(size_t);
    Teuchos::RCP<vec_type> offsetViewNonConst(const Teuchos::RCP<const map_type> &sub, size_t offset);

    template <class Space>
    Kokkos::View2D<S> getLocalView() const { return Kokkos::View2D<S>(ptr(), n_); }
    template <class Space>
    void modify() {}
    template <class Space>
    void sync() {}
    template <class Space>
    bool need_sync() const { return false; }

    // elementwise kernels run on all OpenMP threads, as Kokkos' OpenMP backend would (results do not depend on it)
    void putScalar(S a) {
        S *y = ptr();
        const long long n = (long long)n_;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) y[i] = a;
    }
    void randomize() { randomize(S(-1), S(1)); }
    void randomize(S lo, S hi) {
        static std::mt19937_64 gen(20211011ull);
        std::uniform_real_distribution<S> d(lo, hi);
        for (size_t i = 0; i < n_; i++) ptr()[i] = d(gen);
    }
    void scale(S a) {
        S *y = ptr();
        const long long n = (long long)n_;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) y[i] = a * y[i];
    }
    void scale(S a, const MultiVector &A) {
        S *y = ptr();
        const S *x = A.ptr();
        const long long n = (long long)n_;
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) y[i] = a * x[i];
    }
    void update(S a, const MultiVector &A, S b) {
        S *y = ptr();
        const S *x = A.ptr();
        const long long n = (long long)n_;
        if (b == S(0)) {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) y[i] = a * x[i];
        } else {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) y[i] = a * x[i] + b * y[i];
        }
    }
    void update(S a, const MultiVector &A, S b, const MultiVector &B, S g) {
        S *z = ptr();
        const S *x = A.ptr(), *y = B.ptr();
        const long long n = (long long)n_;
        if (g == S(0)) {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) z[i] = a * x[i] + b * y[i];
        } else {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) z[i] = a * x[i] + b * y[i] + g * z[i];
        }
    }
    // this = t*this + s*A(i)*B(i)
    void elementWiseMultiply(S s, const MultiVector &A, const MultiVector &B, S t) {
        S *c = ptr();
        const S *a = A.ptr(), *b = B.ptr();
        const long long n = (long long)n_;
        if (t == S(0)) {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) c[i] = s * a[i] * b[i];
        } else {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) c[i] = t * c[i] + s * a[i] * b[i];
        }
    }
    // fixed chunks of 4096 entries, each summed left to right, partial sums added in chunk order: independent of the
    // thread count (Kokkos' own order depends on it) and the same order oracle/alens_oracle.c uses
    S dot(const MultiVector &A) const {
        const S *x = ptr(), *y = A.ptr();
        const long long n = (long long)n_, nch = (n + 4095) / 4096;
        std::vector<S> part((size_t)std::max<long long>(nch, 1), S(0));
#pragma omp parallel for schedule(static)
        for (long long c = 0; c < nch; c++) {
            const long long e = std::min<long long>((c + 1) * 4096, n);
            S s = 0;
            for (long long i = c * 4096; i < e; i++) s += x[i] * y[i];
            part[(size_t)c] = s;
        }
        S s = 0;
        for (long long c = 0; c < nch; c++) s += part[(size_t)c];
        return s;
    }
    S norm2() const { return std::sqrt(dot(*this)); }
    S normInf() const {
        S m = 0;
        const S *x = ptr();
        const long long n = (long long)n_;
#pragma omp parallel for schedule(static) reduction(max : m)
        for (long long i = 0; i < n; i++) m = std::max(m, std::fabs(x[i]));
        return m;
    }
    S norm1() const {
        S m = 0;
        for (size_t i = 0; i < n_; i++) m += std::fabs(ptr()[i]);
        return m;
    }
    template <class Imp>
    void doImport(const MultiVector &src, const Imp &, CombineMode) {
        const size_t m = std::min(n_, src.n_);
        for (size_t i = 0; i < m; i++) ptr()[i] = src.ptr()[i];
    }
    std::string description() const { return "Tpetra::Vector(stub){length = " + std::to_string(n_) + "}"; }

  protected:
    Teuchos::RCP<const map_type> map_;
    std::shared_ptr<std::vector<S>> store_;
    size_t off_, n_;
};

template <class S = double, class LO = int, class GO = int, class Node = void>
class Vector : public MultiVector<S, LO, GO, Node> {
  public:
    typedef MultiVector<S, LO, GO, Node> base;
    using base::base;
    Vector(const base &b) : base(b, Teuchos::View) {}
    Vector(const Vector &src, Teuchos::DataAccess acc) : base(src, acc) {}
};

template <class S, class LO, class GO, class Node>
Teuchos::RCP<const Vector<S, LO, GO, Node>> MultiVector<S, LO, GO, Node>::getVector(size_t) const {
    return Teuchos::rcp(new const vec_type(*this)); // shares the storage
}
template <class S, class LO, class GO, class Node>
Teuchos::RCP<Vector<S, LO, GO, Node>> MultiVector<S, LO, GO, Node>::getVectorNonConst(size_t) {
    return Teuchos::rcp(new vec_type(*this));
}
template <class S, class LO, class GO, class Node>
Teuchos::RCP<Vector<S, LO, GO, Node>> MultiVector<S, LO, GO, Node>::offsetViewNonConst(const Teuchos::RCP<const map_type> &sub,
                                                                                      size_t offset) {
    return Teuchos::rcp(new vec_type(sub, store_, off_ + offset));
}

template <class S = double, class LO = int, class GO = int, class Node = void>
class Operator {
  public:
    typedef S scalar_type;
    typedef MultiVector<S, LO, GO, Node> mv_type;
    typedef Map<LO, GO, Node> map_type;
    virtual ~Operator() {}
    virtual void apply(const mv_type &X, mv_type &Y, Teuchos::ETransp mode = Teuchos::NO_TRANS,
                       S alpha = Teuchos::ScalarTraits<S>::one(), S beta = Teuchos::ScalarTraits<S>::zero()) const = 0;
    virtual Teuchos::RCP<const map_type> getDomainMap() const = 0;
    virtual Teuchos::RCP<const map_type> getRangeMap() const = 0;
    virtual bool hasTransposeApply() const { return false; }
    virtual std::string description() const { return "Tpetra::Operator(stub)"; }
};

template <class S = double, class LO = int, class GO = int, class Node = void>
class CrsMatrix : public Operator<S, LO, GO, Node> {
  public:
    typedef Map<LO, GO, Node> map_type;
    typedef MultiVector<S, LO, GO, Node> mv_type;
    // column indices are LOCAL indices into colMap
    template <class RP, class CI, class VA>
    CrsMatrix(const Teuchos::RCP<const map_type> &rowMap, const Teuchos::RCP<const map_type> &colMap, const RP &rowPtr,
              const CI &colInd, const VA &values)
        : rowMap_(rowMap), colMap_(colMap) {
        const size_t nr = rowMap->getNodeNumElements();
        ptr_.resize(nr + 1);
        for (size_t i = 0; i <= nr; i++) ptr_[i] = (size_t)rowPtr[i];
        ind_.resize(ptr_[nr]);
        val_.resize(ptr_[nr]);
        for (size_t k = 0; k < ptr_[nr]; k++) {
            ind_[k] = (LO)colInd[k];
            val_[k] = values[k];
        }
    }
    void fillComplete(const Teuchos::RCP<const map_type> &domainMap, const Teuchos::RCP<const map_type> &rangeMap) {
        domainMap_ = domainMap;
        rangeMap_ = rangeMap;
        // single rank: the Import from the domain map into the column map is a local permutation
        colToDom_.resize(colMap_->getNodeNumElements());
        for (size_t c = 0; c < colToDom_.size(); c++) {
            colToDom_[c] = domainMap_->getLocalElement(colMap_->getGlobalElement((LO)c));
            if (colToDom_[c] == Teuchos::OrdinalTraits<LO>::invalid())
                throw std::runtime_error("tpetra stub: a matrix column is not owned by this (single) rank");
        }
        filled_ = true;
    }
    void fillComplete() { fillComplete(rowMap_, rowMap_); }
    bool isFillComplete() const { return filled_; }
    void apply(const mv_type &X, mv_type &Y, Teuchos::ETransp mode = Teuchos::NO_TRANS, S alpha = S(1),
               S beta = S(0)) const override {
        if (mode != Teuchos::NO_TRANS) throw std::runtime_error("tpetra stub: transposed CrsMatrix::apply is not needed by the path");
        const S *x = X.ptr();
        S *y = Y.ptr();
        const size_t nr = ptr_.size() - 1;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < nr; i++) {
            S s = 0;
            for (size_t k = ptr_[i]; k < ptr_[i + 1]; k++) s += val_[k] * x[colToDom_[ind_[k]]];
            if (beta == S(0)) y[i] = alpha * s;
            else y[i] = beta * y[i] + alpha * s;
        }
    }
    Teuchos::RCP<const map_type> getDomainMap() const override { return domainMap_; }
    Teuchos::RCP<const map_type> getRangeMap() const override { return rangeMap_; }
    Teuchos::RCP<const map_type> getRowMap() const { return rowMap_; }
    Teuchos::RCP<const map_type> getColMap() const { return colMap_; }
    size_t getNodeNumRows() const { return ptr_.size() - 1; }
    size_t getNodeNumEntries() const { return ind_.size(); }
    global_size_t getGlobalNumRows() const { return getNodeNumRows(); }
    global_size_t getGlobalNumEntries() const { return getNodeNumEntries(); }
    std::string description() const override {
        std::ostringstream os;
        os << "Tpetra::CrsMatrix(stub){rows = " << getNodeNumRows() << ", cols = " << domainMap_->getNodeNumElements()
           << ", nnz = " << ind_.size() << "}";
        return os.str();
    }
    // raw access for the transposer / writers / the test driver
    const std::vector<size_t> &rowPtr() const { return ptr_; }
    const std::vector<LO> &colInd() const { return ind_; }
    const std::vector<S> &values() const { return val_; }
    LO domainIndexOfColumn(LO c) const { return colToDom_[c]; }

  private:
    Teuchos::RCP<const map_type> rowMap_, colMap_, domainMap_, rangeMap_;
    std::vector<size_t> ptr_;
    std::vector<LO> ind_;
    std::vector<S> val_;
    std::vector<LO> colToDom_;
    bool filled_ = false;
};

template <class S = double, class LO = int, class GO = int, class Node = void>
class RowMatrixTransposer {
  public:
    typedef CrsMatrix<S, LO, GO, Node> crs_type;
    typedef Map<LO, GO, Node> map_type;
    explicit RowMatrixTransposer(const Teuchos::RCP<const crs_type> &A) : A_(A) {}
    // rows of the transpose = domain map of A, entries of a row sorted by column (= A's row index): counting sort,
    // which visits A's rows in ascending order and therefore produces sorted rows
    Teuchos::RCP<crs_type> createTranspose() {
        const size_t nr = A_->getNodeNumRows(), ncT = nr;
        const size_t nrT = A_->getDomainMap()->getNodeNumElements();
        std::vector<size_t> ptr(nrT + 1, 0);
        const auto &ap = A_->rowPtr();
        const auto &ai = A_->colInd();
        const auto &av = A_->values();
        for (size_t k = 0; k < ai.size(); k++) ptr[A_->domainIndexOfColumn(ai[k]) + 1]++;
        for (size_t r = 0; r < nrT; r++) ptr[r + 1] += ptr[r];
        std::vector<size_t> fill(ptr.begin(), ptr.end() - 1);
        std::vector<LO> ind(ai.size());
        std::vector<S> val(ai.size());
        for (size_t i = 0; i < nr; i++)
            for (size_t k = ap[i]; k < ap[i + 1]; k++) {
                const size_t p = fill[A_->domainIndexOfColumn(ai[k])]++;
                ind[p] = (LO)i;
                val[p] = av[k];
            }
        (void)ncT;
        // column map of the transpose = range (row) map of A: local index = A's local row
        Teuchos::RCP<const map_type> rowMapT = A_->getDomainMap(), colMapT = A_->getRangeMap();
        Teuchos::RCP<crs_type> T = Teuchos::rcp(new crs_type(rowMapT, colMapT, ptr, ind, val));
        T->fillComplete(A_->getRangeMap(), A_->getDomainMap());
        return T;
    }

  private:
    Teuchos::RCP<const crs_type> A_;
};

namespace MatrixMarket {
// writers: real MatrixMarket files, so that the reference's self-test problem (BCQPSolver::selfTest dumps A, b, lb,
// ub and the solution) can be read back by the parity tests
template <class T>
class Writer {
  public:
    template <class Mat>
    static void writeSparseFile(const std::string &fn, const Teuchos::RCP<Mat> &A, const std::string & = "",
                                const std::string & = "", bool = false) {
        std::ofstream f(fn);
        f << "%%MatrixMarket matrix coordinate real general\n";
        f << A->getNodeNumRows() << " " << A->getDomainMap()->getNodeNumElements() << " " << A->getNodeNumEntries() << "\n";
        f << std::setprecision(17);
        const auto &p = A->rowPtr();
        for (size_t i = 0; i + 1 < p.size(); i++)
            for (size_t k = p[i]; k < p[i + 1]; k++)
                f << i + 1 << " " << A->domainIndexOfColumn(A->colInd()[k]) + 1 << " " << A->values()[k] << "\n";
    }
    template <class Vec>
    static void writeDenseFile(const std::string &fn, const Teuchos::RCP<Vec> &B, const std::string & = "",
                               const std::string & = "") {
        std::ofstream f(fn);
        f << "%%MatrixMarket matrix array real general\n" << B->getLocalLength() << " 1\n" << std::setprecision(17);
        for (size_t i = 0; i < B->getLocalLength(); i++) f << B->ptr()[i] << "\n";
    }
    template <class M>
    static void writeMapFile(const std::string &fn, const M &map) {
        std::ofstream f(fn);
        f << "%%MatrixMarket matrix array integer general\n" << map.getNodeNumElements() << " 1\n";
        for (size_t i = 0; i < map.getNodeNumElements(); i++) f << map.getGlobalElement((int)i) << "\n";
    }
};
} // namespace MatrixMarket

} // namespace Tpetra

// ---------------------------------------------------------------------------------------------- Belos
namespace Belos {
enum ETrans { NOTRANS, TRANS, CONJTRANS };
template <class Scalar, class MV, class OP>
class OperatorTraits {};
} // namespace Belos
