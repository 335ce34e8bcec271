// collide.cu -- rod packing, uniform cell list and the pair-collision kernels (broad + narrow phase).
//
// Replaces, for the collision path, SylinderSystem::prepareStep's per-rod loop
// (SimToolbox/Sylinder/SylinderSystem.cpp:897-905), SylinderNearEP::copyFromFP (SylinderNear.hpp:74-90),
// the FDPS tree build / neighbour walk behind TreeSylinderNear::calcForceAll
// (FDPS/tree_for_force.hpp:759-842) and CalcSylinderNearForce::operator() (SylinderNear.hpp:197-414).
//
// COMPILED WITH -fmad=false (see geometry.cuh): the pair list is an integer result and must be
// reproducible bit for bit against the CPU path.
//
// Pipeline (all on ctx.stream):
//   k_rod_pack      wrap into box, cell id, per-cell histogram           (1 thread / rod)
//   k_scan_int      exclusive scan of the histogram                      (single CTA)
//   k_cell_scatter  counting-sort scatter                                (1 thread / rod)
//   k_cell_order    per-cell sort by user index (determinism) + gather of the sorted SoA (1 warp / cell)
//   k_pairs<false>  count pass: one warp per cell, half stencil (14 cells); centre-distance broad phase
//                   with warp ballot compaction into a shared-memory queue, dense narrow-phase batches
//   k_scan_int      exclusive scan of per-cell hit counts
//   k_pairs<true>   fill pass: same traversal, writes the constraint SoA at deterministic offsets
#include "context.hpp"
#include "geometry.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace alens {

static constexpr int kWarpsPerCta = 4;
static constexpr int kITile = 64;  // target rods staged per warp
static constexpr int kQueue = 64;  // compaction queue entries per warp

// ------------------------------------------------------------------------------------------------
// rod_pack: applyBoxBC (FDPS/particle_system.hpp:798-843) + cell id + histogram
__global__ void k_rod_pack(int n, double *__restrict__ pos, Box box, CellGrid g, int wrap, int *__restrict__ cellOf,
                           int *__restrict__ cellCount) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    int c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double x = p[k];
        if (wrap) {
            const double len = box.len[k];
            if (len > 0 && isfinite(x)) {
                if (fabs(x - box.lo[k]) > 64.0 * len) x = box.lo[k] + fmod(x - box.lo[k], len); // far-away guard
                while (x < box.lo[k]) x += len;
                while (x >= box.hi[k]) x -= len;
                if (x == box.hi[k]) x = box.lo[k];
            }
            p[k] = x;
        }
        int ci = (int)floor((x - box.lo[k]) * g.inv[k]);
        ci = ci < 0 ? 0 : (ci >= g.n[k] ? g.n[k] - 1 : ci);
        c[k] = ci;
    }
    if (wrap) {
        pos[3 * i] = p[0];
        pos[3 * i + 1] = p[1];
        pos[3 * i + 2] = p[2];
    }
    const int cell = (c[2] * g.n[1] + c[1]) * g.n[0] + c[0];
    cellOf[i] = cell;
    atomicAdd(&cellCount[cell], 1);
}

// single-CTA exclusive scan; out has n+1 entries (out[n] = total).  n up to a few million.
__global__ void k_scan_int(const int *__restrict__ in, int *__restrict__ out, int n) {
    __shared__ int sPart[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int chunk = (n + T - 1) / T;
    const int b = t * chunk, e = min(n, b + chunk);
    int s = 0;
    for (int i = b; i < e; i++) s += in[i];
    sPart[t] = s;
    __syncthreads();
    // inclusive scan of partials (Hillis-Steele)
    for (int off = 1; off < T; off <<= 1) {
        int v = (t >= off) ? sPart[t - off] : 0;
        __syncthreads();
        sPart[t] += v;
        __syncthreads();
    }
    int run = (t == 0) ? 0 : sPart[t - 1];
    for (int i = b; i < e; i++) {
        const int v = in[i];
        out[i] = run;
        run += v;
    }
    if (t == T - 1) out[n] = sPart[T - 1];
}

__global__ void k_cell_scatter(int n, const int *__restrict__ cellOf, const int *__restrict__ cellStart,
                               int *__restrict__ cellFill, int *__restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cellOf[i];
    const int slot = cellStart[c] + atomicAdd(&cellFill[c], 1);
    order[slot] = i;
}

struct RodArrays {
    // user order inputs
    const int *uGid;
    const double *uPos, *uQuat, *uLen, *uRad;
    const unsigned char *uImm;
    // sorted outputs
    int *sUser, *sGid, *userToSorted;
    double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLc, *sRc, *sLen, *sRad;
    unsigned char *sImm;
};

// one warp per cell: rank-sort the cell's rods by user index, then gather/compute the sorted SoA.
// direction = q * (0,0,1) as Eigen evaluates it (SylinderNear.hpp:86): uv = q.vec x v; uv += uv;
// v + w*uv + q.vec x uv.
__global__ void k_cell_order(int ncell, const int *__restrict__ cellStart, const int *__restrict__ order,
                             RodArrays a, double dRatio, double lRatio) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ncell) return;
    const int b = cellStart[warp], e = cellStart[warp + 1];
    const int n = e - b;
    for (int m = lane; m < n; m += 32) {
        const int u = order[b + m];
        int rank = 0;
        for (int k = 0; k < n; k++) rank += (order[b + k] < u) ? 1 : 0;
        const int s = b + rank;
        a.sUser[s] = u;
        a.userToSorted[u] = s;
        a.sGid[s] = a.uGid[u];
        a.sX[s] = a.uPos[3 * u];
        a.sY[s] = a.uPos[3 * u + 1];
        a.sZ[s] = a.uPos[3 * u + 2];
        const double qx = a.uQuat[4 * u], qy = a.uQuat[4 * u + 1], qz = a.uQuat[4 * u + 2], qw = a.uQuat[4 * u + 3];
        const Vec3 qv = v3(qx, qy, qz), ez = v3(0, 0, 1);
        Vec3 uv = cross(qv, ez);
        uv = v3(uv.x + uv.x, uv.y + uv.y, uv.z + uv.z);
        const Vec3 c2 = cross(qv, uv);
        a.sDx[s] = (ez.x + qw * uv.x) + c2.x;
        a.sDy[s] = (ez.y + qw * uv.y) + c2.y;
        a.sDz[s] = (ez.z + qw * uv.z) + c2.z;
        const double len = a.uLen[u], rad = a.uRad[u];
        a.sLen[s] = len;
        a.sRad[s] = rad;
        a.sLc[s] = len * lRatio;
        a.sRc[s] = rad * dRatio;
        a.sImm[s] = a.uImm ? a.uImm[u] : 0;
    }
}

// ------------------------------------------------------------------------------------------------
struct PairIn {
    const int *cellStart;
    const int *sGid;
    const double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLc, *sRc;
};
struct PairOut {
    int *idxI, *idxJ, *gidI, *gidJ;
    signed char *shift;
    double *delta0, *gamma0;
    double *n, *pI, *pJ, *labI, *labJ; // [3][stride]
    size_t stride;
};

__device__ __forceinline__ RodGeom loadRod(const PairIn &in, int s) {
    RodGeom r;
    r.c = v3(in.sX[s], in.sY[s], in.sZ[s]);
    r.d = v3(in.sDx[s], in.sDy[s], in.sDz[s]);
    r.lc = in.sLc[s];
    r.rc = in.sRc[s];
    return r;
}

// Narrow phase for up to 32 queued candidates (one per lane); returns the number of hits.
// Canonical roles (reference: gid filter SylinderNear.hpp:210,225 + FDPS image rule
// FDPS/tree_for_force_utils.hpp:256-262): I = lower gid at its own position, J = higher gid at
// pos + k*boxLen where k is J's image relative to I.
template <bool FILL>
__device__ __forceinline__ int narrowBatch(const PairIn &in, const PairOut &out, const Box &box, double colBuf,
                                           int cnt, const int *qi, const int *qj, const int *qs, int lane,
                                           int outBase) {
    bool hit = false;
    Contact ct;
    int si = 0, sj = 0, code = 13;
    if (lane < cnt) {
        si = qi[lane];
        sj = qj[lane];
        code = qs[lane]; // image of sj relative to si
        int kx = code % 3 - 1, ky = (code / 3) % 3 - 1, kz = code / 9 - 1;
        if (in.sGid[si] > in.sGid[sj]) { // swap roles; relative image flips sign
            const int t = si; si = sj; sj = t;
            kx = -kx; ky = -ky; kz = -kz;
            code = (kx + 1) + 3 * (ky + 1) + 9 * (kz + 1);
        }
        RodGeom a = loadRod(in, si), b = loadRod(in, sj);
        b.c = v3(b.c.x + kx * box.len[0], b.c.y + ky * box.len[1], b.c.z + kz * box.len[2]);
        hit = pairContact(a, b, colBuf, ct);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (FILL && hit) {
        const size_t k = (size_t)outBase + __popc(m & ((1u << lane) - 1));
        const size_t S = out.stride;
        out.idxI[k] = si;
        out.idxJ[k] = sj;
        out.gidI[k] = in.sGid[si];
        out.gidJ[k] = in.sGid[sj];
        out.shift[k] = (signed char)code;
        out.delta0[k] = ct.sep;
        out.gamma0[k] = ct.sep < 0 ? -ct.sep : 0;
        out.n[k] = ct.normI.x; out.n[k + S] = ct.normI.y; out.n[k + 2 * S] = ct.normI.z;
        out.pI[k] = ct.posI.x; out.pI[k + S] = ct.posI.y; out.pI[k + 2 * S] = ct.posI.z;
        out.pJ[k] = ct.posJ.x; out.pJ[k + S] = ct.posJ.y; out.pJ[k + 2 * S] = ct.posJ.z;
        out.labI[k] = ct.labI.x; out.labI[k + S] = ct.labI.y; out.labI[k + 2 * S] = ct.labI.z;
        out.labJ[k] = ct.labJ.x; out.labJ[k + S] = ct.labJ.y; out.labJ[k + 2 * S] = ct.labJ.z;
    }
    return __popc(m);
}

template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_pairs(PairIn in, PairOut out, Box box, CellGrid g, double colBuf, int *__restrict__ cellHits,
        const int *__restrict__ cellHitStart, unsigned long long *__restrict__ counters) {
    __shared__ double sI[kWarpsPerCta][4][kITile]; // x, y, z, R of the staged target rods
    __shared__ int sQ[kWarpsPerCta][3][kQueue];    // queue: i, j, image code
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * kWarpsPerCta + w;
    if (cell >= g.ncell) return;
    const int ib = in.cellStart[cell], ie = in.cellStart[cell + 1];
    if (ib == ie) {
        if (!FILL && lane == 0) cellHits[cell] = 0;
        return;
    }
    const int cx = cell % g.n[0], cy = (cell / g.n[0]) % g.n[1], cz = cell / (g.n[0] * g.n[1]);
    int *qi = sQ[w][0], *qj = sQ[w][1], *qs = sQ[w][2];
    int qn = 0;       // queue fill (warp-uniform)
    int nHits = 0;    // hits so far in this cell (warp-uniform)
    unsigned long long nCand = 0;
    const int outBase = FILL ? cellHitStart[cell] : 0;
    const double slack = 1.0 + 1e-10;

    for (int i0 = ib; i0 < ie; i0 += kITile) {
        const int nI = min(kITile, ie - i0);
        __syncwarp();
        for (int m = lane; m < nI; m += 32) {
            const int s = i0 + m;
            sI[w][0][m] = in.sX[s];
            sI[w][1][m] = in.sY[s];
            sI[w][2][m] = in.sZ[s];
            sI[w][3][m] = 0.5 * in.sLc[s] + in.sRc[s];
        }
        __syncwarp();
        // half stencil: self, then the 13 "positive" neighbours
        for (int nb = 0; nb < 14; nb++) {
            int dx, dy, dz;
            if (nb == 0) { dx = 0; dy = 0; dz = 0; }
            else if (nb == 1) { dx = 1; dy = 0; dz = 0; }
            else if (nb < 5) { dx = nb - 3; dy = 1; dz = 0; }
            else { dx = (nb - 5) % 3 - 1; dy = (nb - 5) / 3 - 1; dz = 1; }
            int o[3] = {cx + dx, cy + dy, cz + dz};
            int kimg[3] = {0, 0, 0};
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (o[k] < 0) {
                    if (!box.pbc[k]) ok = false;
                    o[k] += g.n[k];
                    kimg[k] = -1;
                } else if (o[k] >= g.n[k]) {
                    if (!box.pbc[k]) ok = false;
                    o[k] -= g.n[k];
                    kimg[k] = 1;
                }
            }
            if (!ok) continue;
            const int cj = (o[2] * g.n[1] + o[1]) * g.n[0] + o[0];
            const int code = (kimg[0] + 1) + 3 * (kimg[1] + 1) + 9 * (kimg[2] + 1);
            const double shx = kimg[0] * box.len[0], shy = kimg[1] * box.len[1], shz = kimg[2] * box.len[2];
            const int jb = in.cellStart[cj], je = in.cellStart[cj + 1];
            for (int j0 = jb; j0 < je; j0 += 32) {
                const int sj = j0 + lane;
                const bool jv = sj < je;
                double xj = 0, yj = 0, zj = 0, Rj = 0;
                if (jv) {
                    xj = in.sX[sj] + shx;
                    yj = in.sY[sj] + shy;
                    zj = in.sZ[sj] + shz;
                    Rj = 0.5 * in.sLc[sj] + in.sRc[sj];
                }
                for (int m = 0; m < nI; m++) {
                    const int si = i0 + m;
                    bool pass = jv;
                    if (nb == 0) pass = pass && (sj > si);         // own cell: each unordered pair once
                    else pass = pass && (sj != si);                // a rod never pairs with its own image
                    if (pass) {
                        const double ddx = xj - sI[w][0][m], ddy = yj - sI[w][1][m], ddz = zj - sI[w][2][m];
                        const double cut = sI[w][3][m] + Rj + colBuf;
                        pass = (ddx * ddx + ddy * ddy + ddz * ddz) <= cut * cut * slack;
                    }
                    const unsigned msk = __ballot_sync(0xffffffffu, pass);
                    if (msk == 0) continue;
                    if (pass) {
                        const int p = qn + __popc(msk & ((1u << lane) - 1));
                        qi[p] = si;
                        qj[p] = sj;
                        qs[p] = code;
                    }
                    qn += __popc(msk);
                    nCand += __popc(msk);
                    __syncwarp();
                    if (qn >= 32) {
                        nHits += narrowBatch<FILL>(in, out, box, colBuf, 32, qi, qj, qs, lane, outBase + nHits);
                        __syncwarp();
                        // move the tail to the front
                        const int rem = qn - 32;
                        int ti = 0, tj = 0, ts = 0;
                        if (lane < rem) { ti = qi[32 + lane]; tj = qj[32 + lane]; ts = qs[32 + lane]; }
                        __syncwarp();
                        if (lane < rem) { qi[lane] = ti; qj[lane] = tj; qs[lane] = ts; }
                        qn = rem;
                        __syncwarp();
                    }
                }
            }
        }
    }
    if (qn > 0) nHits += narrowBatch<FILL>(in, out, box, colBuf, qn, qi, qj, qs, lane, outBase + nHits);
    if (!FILL && lane == 0) {
        cellHits[cell] = nHits;
        atomicAdd(&counters[0], nCand);
        atomicAdd(&counters[1], (unsigned long long)nHits);
    }
}

void launchScanInt(const int *in, int *out, int n, cudaStream_t st) { k_scan_int<<<1, 1024, 0, st>>>(in, out, n); }

// ------------------------------------------------------------------------------------------------
void ctxInit(Context &c) {
    ALENS_CUDA(cudaSetDevice(c.device));
    ALENS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.ownStream = true;
    for (auto &e : c.ev) ALENS_CUDA(cudaEventCreate(&e));
    c.dScal.reserve(1);
    ALENS_CUDA(cudaMemset(c.dScal.p, 0, sizeof(SolverScalars)));
    ALENS_CUDA(cudaMallocHost((void **)&c.hScal, sizeof(SolverScalars)));
    memset(c.hScal, 0, sizeof(SolverScalars));
    c.dCounters.reserve(4);
}

void ctxFree(Context &c) {
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto &e : c.ev)
        if (e) cudaEventDestroy(e);
    if (c.hScal) cudaFreeHost(c.hScal);
    if (c.ownStream && c.stream) cudaStreamDestroy(c.stream);
}

static void chooseGrid(Context &c, double maxR) {
    CellGrid &g = c.grid;
    g.cutoff = (2 * maxR + c.colBuf) * (1.0 + 1e-9);
    if (!(g.cutoff > 0)) g.cutoff = 1.0;
    long long total = 1;
    for (int k = 0; k < 3; k++) {
        double m = std::floor(c.box.len[k] / g.cutoff);
        if (!(m >= 1)) m = 1;
        if (m > 1024) m = 1024;
        g.n[k] = (int)m;
        total *= g.n[k];
    }
    // bound the cell count (sparse huge boxes): coarser cells are always valid
    const long long maxCells = std::max<long long>(4096, std::min<long long>(8LL * std::max(c.nRods, 1), 1LL << 23));
    while (total > maxCells) {
        int k = 0;
        for (int d = 1; d < 3; d++)
            if (g.n[d] > g.n[k]) k = d;
        total /= g.n[k];
        g.n[k] = (g.n[k] + 1) / 2;
        total *= g.n[k];
    }
    g.ncell = (int)total;
    for (int k = 0; k < 3; k++) g.inv[k] = c.box.len[k] > 0 ? g.n[k] / c.box.len[k] : 0.0;
}

// host: max bounding radius; called with the host arrays at upload time
double hostMaxRadius(int n, const double *len, const double *rad, double lRatio, double dRatio) {
    double m = 0;
    for (int i = 0; i < n; i++) {
        const double R = 0.5 * len[i] * lRatio + rad[i] * dRatio;
        if (R > m) m = R;
    }
    return m;
}

double g_lastMaxR = 0; // set by the C API before rodsUploaded (single-threaded boundary)

void rodsUploaded(Context &c, bool wrap) {
    const int n = c.nRods;
    cudaStream_t st = c.stream;
    chooseGrid(c, g_lastMaxR);
    const CellGrid g = c.grid;
    c.uCell.reserve(n);
    c.userToSorted.reserve(n);
    c.cellCount.reserve(g.ncell + 1);
    c.cellStart.reserve(g.ncell + 1);
    c.cellFill.reserve(g.ncell + 1);
    c.sUser.reserve(n); c.sGid.reserve(n);
    c.sX.reserve(n); c.sY.reserve(n); c.sZ.reserve(n);
    c.sDx.reserve(n); c.sDy.reserve(n); c.sDz.reserve(n);
    c.sLc.reserve(n); c.sRc.reserve(n); c.sLen.reserve(n); c.sRad.reserve(n);
    c.sImm.reserve(n);
    DevBuf<int> &order = c.incFill; // scratch (rebuilt later by setup)
    order.reserve(n + 1);
    ALENS_CUDA(cudaMemsetAsync(c.cellCount.p, 0, sizeof(int) * (g.ncell + 1), st));
    ALENS_CUDA(cudaMemsetAsync(c.cellFill.p, 0, sizeof(int) * (g.ncell + 1), st));
    if (n > 0) {
        k_rod_pack<<<gridFor(n, 256), 256, 0, st>>>(n, c.uPos.p, c.box, g, wrap ? 1 : 0, c.uCell.p, c.cellCount.p);
        c.launches++;
    }
    k_scan_int<<<1, 1024, 0, st>>>(c.cellCount.p, c.cellStart.p, g.ncell);
    c.launches++;
    if (n > 0) {
        k_cell_scatter<<<gridFor(n, 256), 256, 0, st>>>(n, c.uCell.p, c.cellStart.p, c.cellFill.p, order.p);
        RodArrays a{c.uGid.p, c.uPos.p, c.uQuat.p, c.uLen.p, c.uRad.p, c.uImm.p, c.sUser.p, c.sGid.p,
                    c.userToSorted.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p, c.sDz.p, c.sLc.p, c.sRc.p,
                    c.sLen.p, c.sRad.p, c.sImm.p};
        k_cell_order<<<gridFor((long long)g.ncell * 32, 128), 128, 0, st>>>(g.ncell, c.cellStart.p, order.p, a,
                                                                            c.dRatio, c.lRatio);
        c.launches += 2;
    }
    ALENS_CUDA(cudaGetLastError());
    c.sorted = true;
    c.haveMob = false;
    c.haveSetup = false;
    c.haveSolution = false;
    c.nCon = c.nColl = 0;
    c.hostBlocks.clear();
}

void reserveConstraints(Context &c, size_t n, bool keep) {
    if (n <= c.conCap) return;
    // SoA arrays use conCap as the component stride, so growth re-lays them out
    const size_t ncap = n + n / 4 + 1024;
    const size_t old = c.conCap;
    const size_t live = keep ? (size_t)c.nCon : 0;
    cudaStream_t st = c.stream;
    auto grow1 = [&](auto &buf, size_t comps) {
        using T = std::remove_pointer_t<decltype(buf.p)>;
        T *np = nullptr;
        ALENS_CUDA(cudaMalloc(&np, ncap * comps * sizeof(T)));
        if (live && buf.p)
            for (size_t k = 0; k < comps; k++)
                ALENS_CUDA(cudaMemcpyAsync(np + k * ncap, buf.p + k * old, live * sizeof(T), cudaMemcpyDeviceToDevice,
                                           st));
        if (buf.p) {
            ALENS_CUDA(cudaStreamSynchronize(st));
            cudaFree(buf.p);
        }
        buf.p = np;
        buf.cap = ncap * comps;
    };
    grow1(c.cIdxI, 1); grow1(c.cIdxJ, 1); grow1(c.cGidI, 1); grow1(c.cGidJ, 1);
    grow1(c.cN, 3); grow1(c.cPI, 3); grow1(c.cPJ, 3); grow1(c.cLabI, 3); grow1(c.cLabJ, 3);
    grow1(c.cDelta0, 1); grow1(c.cGamma0, 1); grow1(c.cInvKappa, 1); grow1(c.cKappa, 1);
    grow1(c.cBi, 1); grow1(c.cOneSide, 1); grow1(c.cShift, 1);
    c.conCap = ncap;
}

static PairIn pairIn(Context &c) {
    return PairIn{c.cellStart.p, c.sGid.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p, c.sDz.p, c.sLc.p, c.sRc.p};
}

void collectPairs(Context &c) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_collect_pair_collision: call alens_set_rods first"};
    cudaStream_t st = c.stream;
    const CellGrid g = c.grid;
    c.nCon = c.nColl = 0;
    c.hostBlocks.clear();
    c.haveSetup = false;
    c.haveSolution = false;
    c.cellHits.reserve(g.ncell + 1);
    c.cellHitStart.reserve(g.ncell + 1);
    ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 4 * sizeof(unsigned long long), st));
    const int ctas = gridFor(g.ncell, kWarpsPerCta);
    PairOut none{};
    k_pairs<false><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), none, c.box, g, c.colBuf, c.cellHits.p, nullptr,
                                                       c.dCounters.p);
    k_scan_int<<<1, 1024, 0, st>>>(c.cellHits.p, c.cellHitStart.p, g.ncell);
    c.launches += 2;
    int total = 0;
    unsigned long long cnt[2];
    ALENS_CUDA(cudaMemcpyAsync(&total, c.cellHitStart.p + g.ncell, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(cnt, c.dCounters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.statCand = (long long)cnt[0];
    reserveConstraints(c, (size_t)total, false);
    if (total > 0) {
        PairOut out{c.cIdxI.p, c.cIdxJ.p, c.cGidI.p, c.cGidJ.p, c.cShift.p, c.cDelta0.p,
                    c.cGamma0.p, c.cN.p, c.cPI.p, c.cPJ.p, c.cLabI.p, c.cLabJ.p, c.conCap};
        k_pairs<true><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), out, c.box, g, c.colBuf, nullptr,
                                                          c.cellHitStart.p, c.dCounters.p);
        c.launches++;
        ALENS_CUDA(cudaMemsetAsync(c.cBi.p, 0, (size_t)total, st));
        ALENS_CUDA(cudaMemsetAsync(c.cOneSide.p, 0, (size_t)total, st));
        ALENS_CUDA(cudaMemsetAsync(c.cInvKappa.p, 0, (size_t)total * sizeof(double), st));
        ALENS_CUDA(cudaMemsetAsync(c.cKappa.p, 0, (size_t)total * sizeof(double), st));
    }
    ALENS_CUDA(cudaGetLastError());
    c.nCon = c.nColl = total;
}

} // namespace alens
