// bcqp.cu -- BCQPSolver as the reference exposes it to ANY caller (SimToolbox/Constraint/BCQPSolver.hpp:37-111):
//   min 1/2 x^T A x + b^T x   s.t.  lb <= x <= ub
// with a caller-supplied b, caller-supplied bounds (setLowerBound / setUpperBound, default -+DBL_MAX/10) and an operator
// A that is either a CSR matrix uploaded by the caller (the reference's TCMAT, e.g. its own self-test problem
// BCQPSolver(int, double), BCQPSolver.cpp:38-132) or the matrix-free constraint operator of the context's last setup
// (ConstraintOperator.cpp:30-71).  solveBBPGD (BCQPSolver.cpp:134-247) and solveAPGD (:249-389) run as vector kernels on
// the device with the scalar control flow on the host -- the structure of the reference, one kernel per Tpetra call.
// (The fused two-kernel BBPGD loop of solver.cu is the fast path of ConstraintSolver, whose bounds follow the bilateral
// flag; this file is the general front end.)  No CPU fallback: every vector operation below is a CUDA kernel.
#include "context.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace alens {

void operatorApplyDevice(Context &c, const double *x, double *y); // solver.cu: y = (D^T M D + K^-1/dt) x, device vectors

static constexpr int kB = 256;

// one thread per row, entries summed in storage order (the order of the reference's CrsMatrix::apply on a host backend)
__global__ void k_csr_spmv(int n, const long long *__restrict__ rowPtr, const int *__restrict__ col,
                           const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s = 0;
    for (long long p = rowPtr[r]; p < rowPtr[r + 1]; p++) s += val[p] * x[col[p]];
    y[r] = 1.0 * s;
}
// z = a*A + b*B (Tpetra update with gamma = 0) / z += ... variants are not needed by the two loops
__global__ void k_axpby(long long n, double *z, double a, const double *A, double b, const double *B) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = a * A[i] + b * B[i];
}
__global__ void k_add_inplace(long long n, double *y, const double *b) { // y = 1.0*b + 1.0*y (BCQPSolver.cpp:156,197)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = 1.0 * b[i] + 1.0 * y[i];
}
__global__ void k_set(long long n, double *y, double v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = v;
}
// boundProjection (BCQPSolver.cpp:431-459): max with lb, then min with ub
__global__ void k_clamp(long long n, double *x, const double *lb, const double *ub) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = x[i];
    v = fmax(v, lb[i]);
    v = fmin(v, ub[i]);
    x[i] = v;
}
// one pass: up to three dot products and the projected-gradient residual of checkProjectionResidual
// (BCQPSolver.cpp:461-497, Dai & Fletcher 2005 eq. 2.2); out = {a.b, c.d, e.f, max |q|} (+inf: projection error)
struct Red {
    const double *a, *b, *c, *d, *e, *f;
    const double *x, *g, *lb, *ub;
};
__global__ void __launch_bounds__(kB) k_reduce(long long n, Red p, double *partial, unsigned *ticket, double *result) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0, 0, 0, 0};
    if (i < n) {
        if (p.a) v[0] = p.a[i] * p.b[i];
        if (p.c) v[1] = p.c[i] * p.d[i];
        if (p.e) v[2] = p.e[i] * p.f[i];
        if (p.x) {
            const double eps = DBL_EPSILON * 100, x = p.x[i], g = p.g[i], lb = p.lb[i], ub = p.ub[i];
            double q;
            if (x < lb + eps) q = fmin(g, 0.0);
            else if (x > ub - eps) q = fmax(g, 0.0);
            else if (x > lb && x < ub) q = g;
            else q = INFINITY; // "projection error occured"
            v[3] = fabs(q);
            if (q != q) v[3] = INFINITY;
        }
    }
    __shared__ double sh[4][kB / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        v[1] += __shfl_xor_sync(0xffffffffu, v[1], o);
        v[2] += __shfl_xor_sync(0xffffffffu, v[2], o);
        v[3] = fmax(v[3], __shfl_xor_sync(0xffffffffu, v[3], o));
    }
    if (lane == 0)
        for (int k = 0; k < 4; k++) sh[k][w] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 1; j < kB / 32; j++) {
            sh[0][0] += sh[0][j]; sh[1][0] += sh[1][j]; sh[2][0] += sh[2][j];
            sh[3][0] = fmax(sh[3][0], sh[3][j]);
        }
        double *dst = partial + 4 * (size_t)blockIdx.x;
        for (int k = 0; k < 4; k++) dst[k] = sh[k][0];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    double s[4] = {0, 0, 0, 0}; // fixed order: run-to-run reproducible
    const volatile double *pp = partial;
    for (unsigned j = 0; j < gridDim.x; j++) {
        s[0] += pp[4 * j]; s[1] += pp[4 * j + 1]; s[2] += pp[4 * j + 2];
        s[3] = fmax(s[3], pp[4 * j + 3]);
    }
    for (int k = 0; k < 4; k++) result[k] = s[k];
    *ticket = 0;
}

struct Bcqp {
    Context *c = nullptr;
    int n = 0;
    int mode = 0; // 0: CSR matrix, 1: the constraint operator of the context's last setup
    DevBuf<long long> rowPtr;
    DevBuf<int> col;
    DevBuf<double> val, b, lb, ub;
    DevBuf<double> partial, result;
    DevBuf<unsigned> ticket;
    std::vector<double> hist;
    alens_solve_report rep{};
    std::string err;
};

static void apply(Bcqp &q, const double *x, double *y) {
    Context &c = *q.c;
    if (q.mode == 1) {
        operatorApplyDevice(c, x, y);
        return;
    }
    k_csr_spmv<<<gridFor(q.n, kB), kB, 0, c.stream>>>(q.n, q.rowPtr.p, q.col.p, q.val.p, x, y);
    c.launches++;
}
static void reduce(Bcqp &q, const Red &r, double out[4]) {
    Context &c = *q.c;
    const int grid = gridFor(std::max(q.n, 1), kB);
    q.partial.reserve(4 * (size_t)grid);
    k_reduce<<<grid, kB, 0, c.stream>>>(q.n, r, q.partial.p, q.ticket.p, q.result.p);
    c.launches++;
    ALENS_CUDA(cudaMemcpyAsync(out, q.result.p, 32, cudaMemcpyDeviceToHost, c.stream));
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
}
static void setDefaultBounds(Bcqp &q, bool lower, bool upper) { // BCQPSolver.cpp:499-510
    Context &c = *q.c;
    const int grid = gridFor(std::max(q.n, 1), kB);
    if (lower) k_set<<<grid, kB, 0, c.stream>>>(q.n, q.lb.p, -DBL_MAX / 10);
    if (upper) k_set<<<grid, kB, 0, c.stream>>>(q.n, q.ub.p, DBL_MAX / 10);
}

Bcqp *bcqpCreate(Context &c, int n, const long long *rowPtr, const int *col, const double *val, const double *b) {
    Bcqp *q = new Bcqp();
    q->c = &c;
    cudaStream_t st = c.stream;
    if (rowPtr) {
        if (n < 0 || !col || !val || !b) throw ArgError{ALENS_ERR_ARG, "alens_bcqp_create: null input"};
        q->mode = 0;
        q->n = n;
        const size_t nnz = (size_t)rowPtr[n];
        q->rowPtr.reserve((size_t)n + 1); q->col.reserve(nnz + 1); q->val.reserve(nnz + 1);
        ALENS_CUDA(cudaMemcpyAsync(q->rowPtr.p, rowPtr, 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
        if (nnz) {
            ALENS_CUDA(cudaMemcpyAsync(q->col.p, col, 4 * nnz, cudaMemcpyHostToDevice, st));
            ALENS_CUDA(cudaMemcpyAsync(q->val.p, val, 8 * nnz, cudaMemcpyHostToDevice, st));
        }
    } else {
        if (!c.haveSetup) throw ArgError{ALENS_ERR_STATE, "alens_bcqp_create: the constraint operator needs alens_setup_constraints"};
        if (c.comm.active) throw ArgError{ALENS_ERR_UNSUPPORTED, "alens_bcqp_create: single-rank front end"};
        q->mode = 1;
        q->n = (int)c.nCon;
    }
    const size_t N = (size_t)q->n + 1;
    q->b.reserve(N); q->lb.reserve(N); q->ub.reserve(N);
    q->result.reserve(4); q->ticket.reserve(1);
    ALENS_CUDA(cudaMemsetAsync(q->ticket.p, 0, sizeof(unsigned), st));
    if (b) ALENS_CUDA(cudaMemcpyAsync(q->b.p, b, 8 * (size_t)q->n, cudaMemcpyHostToDevice, st));
    else if (q->mode == 1 && q->n > 0) ALENS_CUDA(cudaMemcpyAsync(q->b.p, c.vB.p, 8 * (size_t)q->n, cudaMemcpyDeviceToDevice, st));
    setDefaultBounds(*q, true, true);
    ALENS_CUDA(cudaStreamSynchronize(st));
    return q;
}

void bcqpSetBounds(Bcqp &q, const double *lb, const double *ub, int which) {
    cudaStream_t st = q.c->stream;
    if (which & 1) {
        if (lb) ALENS_CUDA(cudaMemcpyAsync(q.lb.p, lb, 8 * (size_t)q.n, cudaMemcpyHostToDevice, st));
        else setDefaultBounds(q, true, false);
    }
    if (which & 2) {
        if (ub) ALENS_CUDA(cudaMemcpyAsync(q.ub.p, ub, 8 * (size_t)q.n, cudaMemcpyHostToDevice, st));
        else setDefaultBounds(q, false, true);
    }
    ALENS_CUDA(cudaStreamSynchronize(st));
}
void bcqpGetBounds(Bcqp &q, double *lb, double *ub) {
    cudaStream_t st = q.c->stream;
    if (lb) ALENS_CUDA(cudaMemcpyAsync(lb, q.lb.p, 8 * (size_t)q.n, cudaMemcpyDeviceToHost, st));
    if (ub) ALENS_CUDA(cudaMemcpyAsync(ub, q.ub.p, 8 * (size_t)q.n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
}

// BCQPSolver::solveBBPGD (BCQPSolver.cpp:134-247), one kernel per vector operation of the reference
static int runBBPGD(Bcqp &q, double *x, double tol, int iteMax) {
    Context &c = *q.c;
    cudaStream_t st = c.stream;
    const long long n = q.n;
    const int grid = gridFor(std::max<long long>(n, 1), kB);
    const size_t N = (size_t)n + 1;
    DevBuf<double> bxk, bxkm1, bgk, bgkm1, bxd, bgd;
    bxk.reserve(N); bxkm1.reserve(N); bgk.reserve(N); bgkm1.reserve(N); bxd.reserve(N); bgd.reserve(N);
    double *xk = bxk.p, *xkm1 = bxkm1.p, *gk = bgk.p, *gkm1 = bgkm1.p, *xd = bxd.p, *gd = bgd.p;
    auto &H = q.hist;
    H.clear();
    int mv = 0, ite = 0;
    ALENS_CUDA(cudaMemcpyAsync(xk, x, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    ALENS_CUDA(cudaMemcpyAsync(xkm1, xk, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    apply(q, xkm1, gkm1);
    mv++;
    k_add_inplace<<<grid, kB, 0, st>>>(n, gkm1, q.b.p);
    double r[4];
    reduce(q, Red{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, xkm1, gkm1, q.lb.p, q.ub.p}, r);
    double resPhi = r[3];
    H.insert(H.end(), {0.0, 0, 0, 0, resPhi, 1.0 * mv});
    bool stag = false, perr = std::isinf(resPhi) || std::isnan(resPhi);
    double alpha = 0;
    double *ret = xkm1;
    if (!perr && !(std::fabs(resPhi) < tol)) {
        alpha = 1.0 / resPhi; // 1 / ||projected gradient||_inf (Dai & Fletcher 2005, section 5)
        while (ite < iteMax) {
            ite++;
            k_axpby<<<grid, kB, 0, st>>>(n, xk, -alpha, gkm1, 1.0, xkm1);
            k_clamp<<<grid, kB, 0, st>>>(n, xk, q.lb.p, q.ub.p);
            apply(q, xk, gk);
            mv++;
            k_add_inplace<<<grid, kB, 0, st>>>(n, gk, q.b.p);
            k_axpby<<<grid, kB, 0, st>>>(n, xd, 1.0, xk, -1.0, xkm1);
            k_axpby<<<grid, kB, 0, st>>>(n, gd, 1.0, gk, -1.0, gkm1);
            c.launches += 5;
            reduce(q, Red{xd, xd, xd, gd, gd, gd, xk, gk, q.lb.p, q.ub.p}, r);
            resPhi = r[3];
            H.insert(H.end(), {1.0 * ite, 0, 0, alpha, resPhi, 1.0 * mv});
            if (std::isinf(resPhi) || std::isnan(resPhi)) { perr = true; break; }
            if (std::fabs(resPhi) < tol) break;
            double a, b;
            if (ite % 2 == 0) { a = r[0]; b = r[1]; } // BB1: |dx|^2 / dx.dg
            else { a = r[1]; b = r[2]; }              // BB2: dx.dg / |dg|^2
            if (std::fabs(b) < 10 * DBL_EPSILON) b += 10 * DBL_EPSILON;
            alpha = a / b;
            if (alpha < DBL_EPSILON * 10) { stag = true; break; }
            std::swap(xkm1, xk);
            std::swap(gkm1, gk);
        }
        ret = xk; // after an iteMax exit the swap has happened: the OLDER iterate, as in the reference (BCQPSolver.cpp:237-241)
    }
    ALENS_CUDA(cudaMemcpyAsync(x, ret, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    q.rep = alens_solve_report{stag ? 1 : 0, ite, mv, (int)(H.size() / 6), resPhi, alpha, n, c.nLocal};
    if (perr) throw ArgError{ALENS_ERR_PROJECTION, "BCQP: projection error (an iterate left [lb, ub] or became NaN)"};
    return stag ? 1 : 0;
}

// BCQPSolver::solveAPGD (BCQPSolver.cpp:249-389)
static int runAPGD(Bcqp &q, double *x, double tol, int iteMax) {
    Context &c = *q.c;
    cudaStream_t st = c.stream;
    const long long n = q.n;
    const int grid = gridFor(std::max<long long>(n, 1), kB);
    const size_t N = (size_t)n + 1;
    DevBuf<double> v[11];
    for (auto &b : v) b.reserve(N);
    double *xk = v[0].p, *yk = v[1].p, *xkp1 = v[2].p, *ykp1 = v[3].p, *gVec = v[4].p, *tempVec = v[5].p, *xhatk = v[6].p,
           *xkdiff = v[7].p, *Axb = v[8].p, *Axbkp1 = v[9].p;
    const double *b = q.b.p;
    auto upd = [&](double *z, double a, const double *A, double bb, const double *B) {
        k_axpby<<<grid, kB, 0, st>>>(n, z, a, A, bb, B);
        c.launches++;
    };
    auto &H = q.hist;
    H.clear();
    int mv = 0;
    ALENS_CUDA(cudaMemcpyAsync(xk, x, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    ALENS_CUDA(cudaMemcpyAsync(yk, xk, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    k_set<<<grid, kB, 0, st>>>(n, xhatk, 1.0);
    upd(xkdiff, -1.0, xhatk, 1.0, xk);
    apply(q, xkdiff, tempVec);
    mv++;
    double r[4];
    reduce(q, Red{tempVec, tempVec, xkdiff, xkdiff, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, r);
    double Lk = std::sqrt(r[0]) / std::sqrt(r[1]);
    double tk = 1.0 / Lk;
    H.insert(H.end(), {0.0, 0, 0, tk, 0, 1.0 * mv});
    int ite = 0;
    bool stag = false, perr = false;
    double thetak = 1, thetakp1 = 1, resmin = DBL_MAX, resPhi = 0;
    while (ite < iteMax) {
        ite++;
        apply(q, yk, Axb);
        mv++;
        upd(gVec, 1.0, b, 1.0, Axb);
        upd(xkp1, 1.0, yk, -tk, gVec);
        k_clamp<<<grid, kB, 0, st>>>(n, xkp1, q.lb.p, q.ub.p);
        reduce(q, Red{yk, Axb, yk, b, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, r);
        const double right1 = r[0] * 0.5, right2 = r[1];
        while (true) {
            upd(xkdiff, 1.0, xkp1, -1.0, yk);
            apply(q, xkp1, Axbkp1);
            mv++;
            reduce(q, Red{xkp1, Axbkp1, xkp1, b, gVec, xkdiff, nullptr, nullptr, nullptr, nullptr}, r);
            const double left1 = r[0] * 0.5, left2 = r[1], right3 = r[2];
            double r2[4];
            reduce(q, Red{xkdiff, xkdiff, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, r2);
            const double nrm = std::sqrt(r2[0]);
            const double right4 = 0.5 * Lk * (nrm * nrm);
            if ((left1 + left2) <= (right1 + right2 + right3 + right4)) break;
            Lk *= 2;
            tk = 1 / Lk;
            upd(xkp1, 1.0, yk, -tk, gVec);
            k_clamp<<<grid, kB, 0, st>>>(n, xkp1, q.lb.p, q.ub.p);
        }
        if (tk < DBL_EPSILON * 10) { stag = true; break; }
        thetakp1 = (-thetak * thetak + thetak * std::sqrt(4 + thetak * thetak)) / 2;
        const double betakp1 = thetak * (1 - thetak) / (thetak * thetak + thetakp1);
        upd(ykp1, (1 + betakp1), xkp1, -betakp1, xk);
        k_add_inplace<<<grid, kB, 0, st>>>(n, Axbkp1, b);
        reduce(q, Red{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, xkp1, Axbkp1, q.lb.p, q.ub.p}, r);
        resPhi = std::fabs(r[3]);
        if (std::isinf(resPhi) || std::isnan(resPhi)) { perr = true; break; }
        if (resPhi < resmin) {
            resmin = resPhi;
            ALENS_CUDA(cudaMemcpyAsync(xhatk, xkp1, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
        }
        H.insert(H.end(), {1.0 * ite, 0, 0, tk, resPhi, 1.0 * mv});
        if (resPhi < tol) break;
        upd(tempVec, 1.0, xkp1, -1.0, xk);
        reduce(q, Red{gVec, tempVec, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, r);
        if (r[0] > 0) {
            ALENS_CUDA(cudaMemcpyAsync(ykp1, xkp1, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
            thetakp1 = 1;
        }
        Lk *= 0.9;
        tk = 1 / Lk;
        std::swap(yk, ykp1);
        std::swap(xk, xkp1);
        thetak = thetakp1;
    }
    ALENS_CUDA(cudaMemcpyAsync(x, xhatk, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    q.rep = alens_solve_report{stag ? 1 : 0, ite, mv, (int)(H.size() / 6), resPhi, tk, n, c.nLocal};
    if (perr) throw ArgError{ALENS_ERR_PROJECTION, "BCQP: projection error (an iterate left [lb, ub] or became NaN)"};
    return stag ? 1 : 0;
}

void bcqpSolve(Bcqp &q, double *x, double tol, int iteMax, int choice, alens_solve_report *rep) {
    if (!x && q.n > 0) throw ArgError{ALENS_ERR_ARG, "alens_bcqp_run: x is the initial guess and the result"};
    if (q.mode == 1 && (!q.c->haveSetup || q.c->nCon != q.n))
        throw ArgError{ALENS_ERR_STATE, "alens_bcqp_run: the constraint operator this problem was created on is gone"};
    if (choice == ALENS_SOLVER_APGD) runAPGD(q, x, tol, iteMax);
    else runBBPGD(q, x, tol, iteMax);
    ALENS_CUDA(cudaGetLastError());
    if (rep) *rep = q.rep;
}
int bcqpHistory(Bcqp &q, double *rows6, int cap) {
    const int n = (int)(q.hist.size() / 6);
    if (rows6 && cap > 0) memcpy(rows6, q.hist.data(), 48 * (size_t)std::min(n, cap));
    return n;
}
Context *bcqpContext(Bcqp &q) { return q.c; }
void bcqpDestroy(Bcqp *q) { delete q; }
int bcqpSize(Bcqp &q) { return q.n; }

} // namespace alens
