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
//   k_pairs_find    one warp per cell, half stencil as 5 x-contiguous rows; broad phase in fp32 with conservative slack:
//                   bounding-sphere tests (passers kept as per-lane bit masks), then the point/axis capsule tests and a
//                   separating-direction test on 32 queued pairs at a time; survivors are staged as 16-byte candidates
//                   (i, j, cell, seq|image) through a warp-aggregated atomic
//   k_cand_narrow   one thread per candidate: exact fp64 closest-point query (DCP); a contact sets its bit in the cell's bitmap
//   k_cell_hit_count + scan: contacts per cell and their exclusive prefix (3-kernel tiled scan)
//   k_pairs_emit2   one thread per candidate with its bit set: contact re-evaluated and written to the constraint SoA at
//                   cellHitStart[cell] + rank of the bit (deterministic order)
//   (k_pairs_find<NARROW=true> + k_pairs_emit: the single-kernel variant, kept as fallback and cross-check)
//   k_long_cells / k_long_long: exact passes for rods longer than the cells were sized for (polydisperse lengths)
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
                           int *__restrict__ cellCount, const signed char *__restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    int c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double x = p[k];
        if (wrap) {
            const double len = box.len[k];
            // only periodic axes: FDPS leaves the root domain at +-LARGE_FLOAT on an open axis
            // (FDPS/domain_info.hpp:1206), so adjustPositionIntoRootDomain never moves a rod along it
            if (box.pbc[k] && len > 0 && isfinite(x)) {
                if (fabs(x - box.lo[k]) > 64.0 * len) x = box.lo[k] + fmod(x - box.lo[k], len); // far-away guard
                while (x < box.lo[k]) x += len;
                while (x >= box.hi[k]) x -= len;
                if (x == box.hi[k]) x = box.lo[k];
            }
            p[k] = x;
        }
        if (k == g.axis && img) x += img[i] * g.axisLen; // a ghost is binned where it appears in this slab's frame
        int ci = (int)floor((x - g.lo[k]) * g.inv[k]);
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

// applyBoxBC alone (multi-rank: the owned rods are wrapped before the ghost exchange)
__global__ void k_rod_wrap(int n, double *__restrict__ pos, Box box) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double x = pos[3 * i + k];
        const double len = box.len[k];
        if (box.pbc[k] && len > 0 && isfinite(x)) {
            if (fabs(x - box.lo[k]) > 64.0 * len) x = box.lo[k] + fmod(x - box.lo[k], len);
            while (x < box.lo[k]) x += len;
            while (x >= box.hi[k]) x -= len;
            if (x == box.hi[k]) x = box.lo[k];
        }
        pos[3 * i + k] = x;
    }
}
__global__ void k_global_index(int n, int base, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = base + i;
}
// Image of every OWNED rod in its slab's frame: a rod that strayed (by less than the skin) across a periodic box
// face has a wrapped coordinate at the far end of the box; img = -1 / +1 brings it back next to its slab.
// Rods further than the skin from their slab are counted: the host has to migrate them.
__global__ void k_local_image(int n, const double *__restrict__ pos, int axis, double lo, double hi, double skin,
                              double boxLen, int periodic, signed char *__restrict__ img, int *__restrict__ strays) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[3 * (size_t)i + axis];
    int im = 0;
    bool ok = x >= lo - skin && x < hi + skin;
    if (!ok && periodic) {
        if (x - boxLen >= lo - skin && x - boxLen < hi + skin) { im = -1; ok = true; }
        else if (x + boxLen >= lo - skin && x + boxLen < hi + skin) { im = 1; ok = true; }
    }
    img[i] = (signed char)im;
    if (!ok) atomicAdd(strays, 1);
}

// single-CTA exclusive scan; out has n+1 entries (out[n] = total).  Used for short arrays and for the
// tile sums of the tiled scan below.
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

// tiled scan: tile sums -> single-CTA scan of the sums -> per-tile rescan with the tile offset
static constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int blockExclusive256(int v, int &total) { // exclusive scan over 256 threads
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int i = 0; i < 8; i++) {
        const int x = wsum[i];
        if (i < w) base += x;
        tot += x;
    }
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const int *__restrict__ in, int n,
                                                                 int *__restrict__ tileSum) {
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) s += (base + i < n) ? in[base + i] : 0;
    int tot;
    blockExclusive256(s, tot);
    if (threadIdx.x == 0) tileSum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_apply(const int *__restrict__ in, int *__restrict__ out,
                                                                  int n, const int *__restrict__ tileOff) {
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int tot;
    int run = tileOff[blockIdx.x] + blockExclusive256(s, tot);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) out[n] = run;
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
    const signed char *uImg;
    int nLocal;
    // sorted outputs
    int *sUser, *sGid, *userToSorted;
    double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLc, *sRc, *sLen, *sRad;
    unsigned char *sImm, *sGhost;
    signed char *sImg;
    float *bUx, *bUy, *bUz, *bH, *bRho;
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
        const double lc = len * lRatio, rc = rad * dRatio;
        a.sLc[s] = lc;
        a.sRc[s] = rc;
        a.sImm[s] = a.uImm ? a.uImm[u] : 0;
        a.sGhost[s] = u >= a.nLocal ? 1 : 0;
        a.sImg[s] = a.uImg[u];
        // broad-phase shape (axis segment of half length h around the centre, thickened by rho): a rod with
        // lc < 2 rc collides as a sphere of radius lc/2 + rc (SylinderNear.hpp:241,259).  The axis is
        // normalised here so that a non-unit quaternion cannot make the capsule tests optimistic.
        const double ddx = a.sDx[s], ddy = a.sDy[s], ddz = a.sDz[s];
        const double dn = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        const bool sphere = lc < 2 * rc;
        const bool okAxis = dn > 0 && isfinite(dn);
        a.bUx[s] = okAxis ? (float)(ddx / dn) : 0.f;
        a.bUy[s] = okAxis ? (float)(ddy / dn) : 0.f;
        a.bUz[s] = okAxis ? (float)(ddz / dn) : 1.f;
        a.bH[s] = (sphere || !okAxis) ? 0.f : __double2float_ru(0.5 * lc * dn);
        a.bRho[s] = __double2float_ru(sphere ? (0.5 * lc + rc) : rc);
    }
}

// ------------------------------------------------------------------------------------------------
struct PairIn {
    const int *cellStart;
    const int *sGid;
    const double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLc, *sRc;
    const float *bUx, *bUy, *bUz, *bH, *bRho;
    const signed char *sImg;      // image along the slab axis (0 on a single rank)
    const unsigned char *sGhost;
};
struct PairOut {
    int *idxI, *idxJ, *gidI, *gidJ;
    signed char *shift;
    unsigned char *bi, *oneSide, *own;
    double *delta0, *gamma0, *invKappa, *kappa;
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

__device__ __forceinline__ void imageOf(int code, int &kx, int &ky, int &kz) {
    kx = code % 3 - 1;
    ky = (code / 3) % 3 - 1;
    kz = code / 9 - 1;
}

// Narrow phase for up to 32 queued candidates (one per lane); returns the number of hits and stages them.
// Canonical roles (reference: gid filter SylinderNear.hpp:210,225 + FDPS image rule
// FDPS/tree_for_force_utils.hpp:256-262): I = lower gid at its own position, J = higher gid at
// pos + k*boxLen where k is J's image relative to I.
__device__ __forceinline__ int narrowBatch(const PairIn &in, const Box &box, double colBuf, int cnt, const int *qi,
                                           const int *qj, const int *qs, int lane, int cell, int seqBase,
                                           int4 *__restrict__ hitList, unsigned long long hitCap,
                                           unsigned long long *__restrict__ counters) {
    bool hit = false;
    int si = 0, sj = 0, code = 13;
    if (lane < cnt) {
        si = qi[lane];
        sj = qj[lane];
        code = qs[lane]; // image of sj relative to si
        int kx, ky, kz;
        imageOf(code, kx, ky, kz);
        if (in.sGid[si] > in.sGid[sj]) { // swap roles; relative image flips sign
            const int t = si; si = sj; sj = t;
            kx = -kx; ky = -ky; kz = -kz;
            code = (kx + 1) + 3 * (ky + 1) + 9 * (kz + 1);
        }
        RodGeom a = loadRod(in, si), b = loadRod(in, sj);
        b.c = v3(b.c.x + kx * box.len[0], b.c.y + ky * box.len[1], b.c.z + kz * box.len[2]);
        Contact ct;
        // the reference's filter is gidI >= gidJ -> skip (SylinderNear.hpp:210,225): equal gids never collide
        hit = in.sGid[si] != in.sGid[sj] && pairContact(a, b, colBuf, ct);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    const int nh = __popc(m);
    if (nh == 0) return 0;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&counters[1], (unsigned long long)nh);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit) {
        const int r = __popc(m & ((1u << lane) - 1));
        const unsigned long long p = base + r;
        if (p < hitCap) hitList[p] = make_int4(si, sj, cell, ((seqBase + r) << 5) | code);
    }
    return nh;
}

// ------------------------------------------------------------------------------------------------
// Search kernel.  One warp per cell.  The half stencil (own cell + 13 "positive" neighbours) is walked as 5 rows of
// x-adjacent cells: cells adjacent in x are adjacent in the sorted arrays, so a row is one contiguous range of source
// rods (plus at most two wrapped single cells at a periodic boundary).  Three stages with two warp-private
// compaction queues, so that every stage runs on (nearly) full warps.
//   stage 1  bounding-sphere test, fp32, lanes = source rods (two per lane: independent dependency chains), loop over a
//            block of 32 targets; a passer is one bit in the lane's two mask registers, and after the block the bits are
//            queued as (target slot, source slot) -- no vote or shared-memory traffic inside the arithmetic loop
//   stage 2  on 32 queued pairs at a time: the two point/axis capsule tests, fp32, operands of both rods from
//            the shared-memory tiles; passers are queued as (i, j, image code)
//   stage 3  narrowBatch: exact fp64 closest-point query on 32 queued candidates at a time
// (The first version of this kernel ran the capsule tests for the whole warp whenever ANY lane passed an fp64 sphere
// test -- 56 % of the iterations for 2.5 % of the lanes; profiles/README.md.)
// The broad phase works on coordinates RELATIVE TO THE CENTRE OF THE TARGET CELL in fp32.  Conservativeness:
// converting a relative coordinate to fp32 moves a centre by at most 2^-24 |coord| per axis; every rod carries
// eps = 2^-22 (|x-Ox| + |y-Oy| + |z-Oz|) in its radius-like terms (rounded up), the fp32 arithmetic (relative error
// ~1e-6 of |dd| <= cut once the sphere test has passed) is covered by the factor 1 + 2e-5 and the absolute slack
// 1e-5 cut.  A pair the fp64 narrow phase would accept is never rejected; what is accepted in excess is decided
// exactly by the narrow phase.  The candidate order (hence the constraint order inside a cell) is deterministic:
// (target tile, stencil row, source tile, block of 32 targets, k-th passing target of a source slot, source slot).
static constexpr int kJTile = 64; // source rods staged per warp

// NARROW = false: the kernel stops after stage 2 and stages its queued candidates (i, j, cell, seq | image) instead of
// running the exact query: without the fp64 closest-point code it needs a third of the registers, and stages 1-2 --
// dependent fp32 arithmetic, bound by how many warps the scheduler can choose from -- run at 2.5x the occupancy.  The
// exact query then runs one candidate per thread (k_cand_narrow), see collectPairs.
__device__ __forceinline__ int stageCandidates(int cnt, const int *qi, const int *qj, const int *qs, int lane, int cell,
                                               int seqBase, int4 *__restrict__ candList, unsigned long long cap,
                                               unsigned long long *__restrict__ counters) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&counters[2], (unsigned long long)cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < cnt) {
        const unsigned long long p = base + lane;
        if (p < cap) candList[p] = make_int4(qi[lane], qj[lane], cell, ((seqBase + lane) << 5) | qs[lane]);
    }
    return cnt;
}

// MULTI = false: the instantiation for a context without ghost rods (no slab decomposition): the ghost bookkeeping of the
// inner loop is compiled out.
template <int MINB, bool NARROW, bool MULTI = true>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB)
k_pairs_find(PairIn in, Box box, CellGrid g, double colBuf, int *__restrict__ cellHits, int4 *__restrict__ hitList,
              unsigned long long hitCap, unsigned long long *__restrict__ counters) {
    __shared__ float4 tA[kWarpsPerCta][kITile]; // target: x, y, z (relative), A = h + rho + colBuf + slack
    __shared__ float4 tB[kWarpsPerCta][kITile]; //         ux, uy, uz, h
    __shared__ float2 tC[kWarpsPerCta][kITile]; //         B = rho + colBuf + slack, B + h
    __shared__ float4 jA[kWarpsPerCta][kJTile]; // source: x, y, z (relative, image applied), S = h + rho + eps
    __shared__ float4 jB[kWarpsPerCta][kJTile]; //         ux, uy, uz, h
    __shared__ float jC[kWarpsPerCta][kJTile];  //         rho + eps
    __shared__ unsigned short sQ1[kWarpsPerCta][32 + 2 * 32]; // stage-1 queue: target slot | source slot << 6
    __shared__ int sQ[kWarpsPerCta][3][kQueue];               // stage-2 queue: i, j, image code
    __shared__ signed char sG[kWarpsPerCta][kITile];  // target: image along the slab axis, +64 if ghost
    __shared__ signed char sGj[kWarpsPerCta][kJTile]; // source: same
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    const int axMul = g.axis == 0 ? 1 : (g.axis == 1 ? 3 : (g.axis == 2 ? 9 : 0));
    const bool multi = MULTI && g.axis >= 0; // ghost rods exist
    const int cell = blockIdx.x * kWarpsPerCta + w;
    if (cell >= g.ncell) return;
    const int ib = in.cellStart[cell], ie = in.cellStart[cell + 1];
    if (ib == ie) {
        if (lane == 0) cellHits[cell] = 0;
        return;
    }
    const int cx = cell % g.n[0], cy = (cell / g.n[0]) % g.n[1], cz = cell / (g.n[0] * g.n[1]);
    const double Ox = g.inv[0] > 0 ? g.lo[0] + (cx + 0.5) / g.inv[0] : g.lo[0];
    const double Oy = g.inv[1] > 0 ? g.lo[1] + (cy + 0.5) / g.inv[1] : g.lo[1];
    const double Oz = g.inv[2] > 0 ? g.lo[2] + (cz + 0.5) / g.inv[2] : g.lo[2];
    unsigned short *q1 = sQ1[w];
    int *qi = sQ[w][0], *qj = sQ[w][1], *qs = sQ[w][2];
    int qn1 = 0;   // stage-1 queue fill (warp-uniform)
    int qn = 0;    // stage-2 queue fill (warp-uniform)
    int nHits = 0; // hits so far in this cell (warp-uniform)
    unsigned long long nCand = 0;
    const float slackF = 1.0f + 2e-5f;
    const double kEps = 2.384185791015625e-07; // 2^-22

    for (int i0 = ib; i0 < ie; i0 += kITile) {
        const int nI = min(kITile, ie - i0);
        __syncwarp();
        for (int m = lane; m < nI; m += 32) {
            const int s = i0 + m;
            double x = in.sX[s], y = in.sY[s], z = in.sZ[s];
            const int im = in.sImg[s];
            sG[w][m] = (signed char)(im + (in.sGhost[s] ? 64 : 0));
            if (im) { // apparent position in this slab's frame
                const double sh = im * g.axisLen;
                if (g.axis == 0) x += sh; else if (g.axis == 1) y += sh; else z += sh;
            }
            const float h = in.bH[s], rho = in.bRho[s];
            // absolute slack: rounding of coordinate differences in the narrow phase + the fp32 conversion here
            const double rx = x - Ox, ry = y - Oy, rz = z - Oz;
            const double sl = 256.0 * DBL_EPSILON * (fabs(x) + fabs(y) + fabs(z) + box.len[0] + box.len[1] + box.len[2]) +
                              kEps * (fabs(rx) + fabs(ry) + fabs(rz));
            const float B = __double2float_ru((double)rho + colBuf + sl);
            tA[w][m] = make_float4((float)rx, (float)ry, (float)rz, __double2float_ru((double)h + (double)rho + colBuf + sl));
            tB[w][m] = make_float4(in.bUx[s], in.bUy[s], in.bUz[s], h);
            tC[w][m] = make_float2(B, __fadd_ru(B, h));
        }
        __syncwarp();
        for (int row = 0; row < 5; row++) {
            const int dy = row == 0 ? 0 : (row == 1 ? 1 : row - 3);
            const int dz = row < 2 ? 0 : 1;
            int oy = cy + dy, oz = cz + dz, ky = 0, kz = 0;
            bool ok = true;
            if (oy < 0) { ok = ok && g.per[1]; oy += g.n[1]; ky = -1; }
            else if (oy >= g.n[1]) { ok = ok && g.per[1]; oy -= g.n[1]; ky = 1; }
            if (oz >= g.n[2]) { ok = ok && g.per[2]; oz -= g.n[2]; kz = 1; }
            if (!ok) continue;
            const int rowBase = (oz * g.n[1] + oy) * g.n[0];
            const int xlo = row == 0 ? cx : cx - 1, xhi = cx + 1;
            for (int seg = 0; seg < 3; seg++) {
                int ca, cb, kx; // cell range [ca, cb] of this row, image in x
                if (seg == 0) { ca = max(xlo, 0); cb = min(xhi, g.n[0] - 1); kx = 0; }
                else if (seg == 1) { if (xlo >= 0 || !g.per[0]) continue; ca = cb = g.n[0] - 1; kx = -1; }
                else { if (xhi < g.n[0] || !g.per[0]) continue; ca = cb = 0; kx = 1; }
                const int jb = in.cellStart[rowBase + ca], je = in.cellStart[rowBase + cb + 1];
                if (jb == je) continue;
                const int code = (kx + 1) + 3 * (ky + 1) + 9 * (kz + 1);
                const bool own = (code == 13) && row == 0; // contains the target cell itself: each pair once
                const double shx = kx * box.len[0] - Ox, shy = ky * box.len[1] - Oy, shz = kz * box.len[2] - Oz;
                for (int j0 = jb; j0 < je; j0 += kJTile) {
                    // ---- stage the source tile: lane owns slots lane and lane + 32.  A slot without a rod sits
                    // far away (the sphere test fails) and has index -1.
                    float xj[2], yj[2], zj[2], Sj[2];
                    int sjv[2];
                    bool gh[2];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int js = 32 * q + lane, sj = j0 + js;
                        const bool jv = sj < je;
                        sjv[q] = jv ? sj : -1;
                        xj[q] = 1e18f; yj[q] = zj[q] = 0.f; Sj[q] = 0.f;
                        gh[q] = false;
                        if (jv) {
                            double x = in.sX[sj] + shx, y = in.sY[sj] + shy, z = in.sZ[sj] + shz;
                            const int im = in.sImg[sj];
                            gh[q] = in.sGhost[sj] != 0;
                            if (im) {
                                const double sh = im * g.axisLen;
                                if (g.axis == 0) x += sh; else if (g.axis == 1) y += sh; else z += sh;
                            }
                            const float e = __double2float_ru(kEps * (fabs(x) + fabs(y) + fabs(z)));
                            const float h = in.bH[sj], r = in.bRho[sj];
                            const float rE = __fadd_ru(r, e);
                            xj[q] = (float)x; yj[q] = (float)y; zj[q] = (float)z;
                            Sj[q] = __fadd_ru(h, rE);
                            jA[w][js] = make_float4(xj[q], yj[q], zj[q], Sj[q]);
                            jB[w][js] = make_float4(in.bUx[sj], in.bUy[sj], in.bUz[sj], h);
                            jC[w][js] = rE;
                            sGj[w][js] = (signed char)(im + (gh[q] ? 64 : 0));
                        }
                    }
                    __syncwarp();
                    // Targets in blocks of 32.  Stage 1 only records, per lane and staged source, WHICH targets of the block pass
                    // (one bit each): no warp vote, no queue traffic inside the arithmetic loop.  The bits are then turned
                    // into stage-1 queue entries, per step one for each of a lane's two source slots (slots `lane` of all lanes
                    // first, then `32 + lane`; a slot's targets in ascending order), stage 2 draining the queue whenever
                    // it holds 32.  The last pass (mb >= nI) only flushes: the staged source tile is about to be replaced.
                    for (int mb = 0;; mb += 32) {
                        const bool flush = mb >= nI;
                        unsigned h0 = 0, h1 = 0;
                        if (!flush) {
                            const int mEnd = min(nI, mb + 32);
                            for (int m = mb; m < mEnd; m++) {
                                // ---- stage 1: target m against the 64 staged sources (branch-free)
                                const float4 a = tA[w][m];
                                // each pair once inside the target cell (the row that holds it lists the sources in sorted
                                // order: those behind the target).  Elsewhere every source counts; a rod can meet itself
                                // there only through a periodic image (fewer than 3 cells on that axis), and equal gids are
                                // dropped by the exact query (canonicalPair / narrowBatch), as the reference's gid filter does.
                                const int after = own ? i0 + m : -2;
                                bool gi = false;
                                if (MULTI) gi = multi && sG[w][m] >= 32;
                                bool pass[2];
#pragma unroll
                                for (int q = 0; q < 2; q++) {
                                    const float fx = xj[q] - a.x, fy = yj[q] - a.y, fz = zj[q] - a.z;
                                    const float cutS = (a.w + Sj[q]) * slackF;
                                    pass[q] = (__fmaf_rn(fx, fx, __fmaf_rn(fy, fy, fz * fz)) <= cutS * cutS) && sjv[q] > after;
                                    if (MULTI) pass[q] = pass[q] && !(gi && gh[q]); // two ghosts: not this rank's constraint
                                }
                                h0 |= (pass[0] ? 1u : 0u) << (m - mb);
                                h1 |= (pass[1] ? 1u : 0u) << (m - mb);
                            }
                        }
                        for (;;) {
                            bool more = true; // (warp-uniform) some lane still holds a passer
                            if (qn1 < 32) {   // room for 64 more entries
                                const unsigned a0 = __ballot_sync(0xffffffffu, h0 != 0), a1 = __ballot_sync(0xffffffffu, h1 != 0);
                                more = (a0 | a1) != 0;
                                if (more) {
                                    const int n0 = __popc(a0);
                                    if (h0) {
                                        q1[qn1 + __popc(a0 & lt)] = (unsigned short)((mb + __ffs(h0) - 1) | (lane << 6));
                                        h0 &= h0 - 1;
                                    }
                                    if (h1) {
                                        q1[qn1 + n0 + __popc(a1 & lt)] = (unsigned short)((mb + __ffs(h1) - 1) | ((32 + lane) << 6));
                                        h1 &= h1 - 1;
                                    }
                                    qn1 += n0 + __popc(a1);
                                    __syncwarp();
                                }
                            }
                            if (!(qn1 >= 32 || (flush && qn1 > 0))) {
                                if (more) continue;
                                break;
                            }
                            {
                            // ---- stage 2 on the first min(32, qn1) queued pairs: the two capsule tests
                            const int cnt = min(32, qn1);
                            bool pass = false;
                            int tm = 0, js = 0;
                            if (lane < cnt) {
                                const unsigned e = q1[lane];
                                tm = e & 63; js = e >> 6;
                                const float4 a = tA[w][tm], b = tB[w][tm], ja = jA[w][js], jb4 = jB[w][js];
                                const float2 c = tC[w][tm];
                                const float rE = jC[w][js];
                                const float fx = ja.x - a.x, fy = ja.y - a.y, fz = ja.z - a.z;
                                const float cutF = (a.w + ja.w) * 1e-5f; // absolute slack of the fp32 evaluation (|dd| <= cut)
                                // dist(c_j, axis_i) <= h_j + rho_i + rho_j + buf, dist(c_i, axis_j) <= h_i + rho_i + rho_j + buf
                                float t = __fmaf_rn(fx, b.x, __fmaf_rn(fy, b.y, fz * b.z));
                                t = fminf(fmaxf(t, -b.w), b.w);
                                float ex = __fmaf_rn(-t, b.x, fx), ey = __fmaf_rn(-t, b.y, fy), ez = __fmaf_rn(-t, b.z, fz);
                                float c2 = __fmaf_rn(c.x + ja.w, slackF, cutF);
                                const bool p1 = __fmaf_rn(ex, ex, __fmaf_rn(ey, ey, ez * ez)) <= c2 * c2;
                                t = -__fmaf_rn(fx, jb4.x, __fmaf_rn(fy, jb4.y, fz * jb4.z));
                                t = fminf(fmaxf(t, -jb4.w), jb4.w);
                                ex = __fmaf_rn(t, jb4.x, fx); ey = __fmaf_rn(t, jb4.y, fy); ez = __fmaf_rn(t, jb4.z, fz);
                                c2 = __fmaf_rn(c.y + rE, slackF, cutF);
                                const bool p2 = __fmaf_rn(ex, ex, __fmaf_rn(ey, ey, ez * ez)) <= c2 * c2;
                                // separating direction w = u_i x u_j (the common perpendicular): both segments project
                                // onto single points of w, so dist(seg_i, seg_j) >= |w . (c_j - c_i)| / |w|.  Only used
                                // when the axes are more than ~6 degrees apart (|w| >= 0.1): then the fp32 errors of w
                                // (1e-7 absolute), of the projections of the half axes (h * 4e-6) and of the dot product
                                // stay below 2 cutF = 2e-5 (A_i + S_j).  Removes ~60 % of what the two capsule tests let
                                // through (3.5 candidates per contact -> 1.4).
                                const float wx = __fmaf_rn(b.y, jb4.z, -b.z * jb4.y), wy = __fmaf_rn(b.z, jb4.x, -b.x * jb4.z),
                                            wz = __fmaf_rn(b.x, jb4.y, -b.y * jb4.x);
                                const float w2 = __fmaf_rn(wx, wx, __fmaf_rn(wy, wy, wz * wz));
                                const float wf = __fmaf_rn(wx, fx, __fmaf_rn(wy, fy, wz * fz));
                                const float c3 = __fmaf_rn(c.x + rE, slackF, 2.0f * cutF);
                                const bool p3 = !(w2 >= 1e-2f && wf * wf > c3 * c3 * w2 * slackF);
                                pass = p1 & p2 & p3;
                            }
                            { // move the tail of the stage-1 queue (< 64 entries) to the front
                                const int rem = qn1 - cnt;
                                unsigned short t0 = 0, t1 = 0;
                                if (lane < rem) t0 = q1[32 + lane];
                                if (32 + lane < rem) t1 = q1[64 + lane];
                                __syncwarp();
                                if (lane < rem) q1[lane] = t0;
                                if (32 + lane < rem) q1[32 + lane] = t1;
                                qn1 = rem;
                            }
                            const unsigned msk = __ballot_sync(0xffffffffu, pass);
                            if (msk) {
                                if (pass) {
                                    const int p = qn + __popc(msk & lt);
                                    qi[p] = i0 + tm;
                                    qj[p] = j0 + js;
                                    // image of j relative to i: cell wrap + difference of the ghost images
                                    int rel = code;
                                    if (MULTI) {
                                        const int gi = sG[w][tm], gj = sGj[w][js];
                                        rel += (((gj + 32) & 63) - ((gi + 32) & 63)) * axMul;
                                    }
                                    qs[p] = rel;
                                }
                                qn += __popc(msk);
                                nCand += __popc(msk);
                                __syncwarp();
                                if (qn >= 32) { // ---- stage 3: exact closest-point query on 32 candidates
                                    if (NARROW)
                                        nHits += narrowBatch(in, box, colBuf, 32, qi, qj, qs, lane, cell, nHits, hitList,
                                                             hitCap, counters);
                                    else nHits += stageCandidates(32, qi, qj, qs, lane, cell, nHits, hitList, hitCap, counters);
                                    __syncwarp();
                                    const int rem = qn - 32; // move the tail to the front
                                    int ti = 0, tj = 0, ts = 0;
                                    if (lane < rem) { ti = qi[32 + lane]; tj = qj[32 + lane]; ts = qs[32 + lane]; }
                                    __syncwarp();
                                    if (lane < rem) { qi[lane] = ti; qj[lane] = tj; qs[lane] = ts; }
                                    qn = rem;
                                }
                            }
                            __syncwarp();
                            }
                        }
                        if (flush) break;
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (qn > 0) {
        if (NARROW) nHits += narrowBatch(in, box, colBuf, qn, qi, qj, qs, lane, cell, nHits, hitList, hitCap, counters);
        else nHits += stageCandidates(qn, qi, qj, qs, lane, cell, nHits, hitList, hitCap, counters);
    }
    if (lane == 0) {
        cellHits[cell] = nHits; // NARROW: contacts of this cell; else: its candidates
        atomicAdd(&counters[0], nCand);
        if (!NARROW) atomicMax(&counters[3], (unsigned long long)nHits);
    }
}

// one thread per staged hit: evaluate the contact (hits only) and write the constraint at its
// deterministic position cellHitStart[cell] + seq
__global__ void __launch_bounds__(128)
k_pairs_emit(long long nHits, const int4 *__restrict__ hitList, const int *__restrict__ cellHitStart, PairIn in,
             PairOut out, Box box, double colBuf) {
    const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nHits) return;
    const int4 rec = hitList[h];
    const int si = rec.x, sj = rec.y, code = rec.w & 31;
    const size_t k = (size_t)cellHitStart[rec.z] + (size_t)(rec.w >> 5);
    int kx, ky, kz;
    imageOf(code, kx, ky, kz);
    RodGeom a = loadRod(in, si), b = loadRod(in, sj);
    b.c = v3(b.c.x + kx * box.len[0], b.c.y + ky * box.len[1], b.c.z + kz * box.len[2]);
    Contact ct;
    pairContact(a, b, colBuf, ct); // same code, same inputs as in k_pairs_find: a hit
    const size_t S = out.stride;
    out.idxI[k] = si;
    out.idxJ[k] = sj;
    out.gidI[k] = in.sGid[si];
    out.gidJ[k] = in.sGid[sj];
    out.shift[k] = (signed char)code;
    out.bi[k] = 0;
    out.oneSide[k] = 0;
    out.own[k] = in.sGhost[si] ? 0 : 1; // the owner of rod I counts the row in global reductions
    out.delta0[k] = ct.sep;
    out.gamma0[k] = ct.sep < 0 ? -ct.sep : 0;
    out.invKappa[k] = 0;
    out.kappa[k] = 0;
    out.n[k] = ct.normI.x; out.n[k + S] = ct.normI.y; out.n[k + 2 * S] = ct.normI.z;
    out.pI[k] = ct.posI.x; out.pI[k + S] = ct.posI.y; out.pI[k + 2 * S] = ct.posI.z;
    out.pJ[k] = ct.posJ.x; out.pJ[k + S] = ct.posJ.y; out.pJ[k + 2 * S] = ct.posJ.z;
    out.labI[k] = ct.labI.x; out.labI[k + S] = ct.labI.y; out.labI[k + 2 * S] = ct.labI.z;
    out.labJ[k] = ct.labJ.x; out.labJ[k + S] = ct.labJ.y; out.labJ[k + 2 * S] = ct.labJ.z;
}

// canonical roles of a staged pair (reference: gid filter SylinderNear.hpp:210,225 + FDPS image rule
// FDPS/tree_for_force_utils.hpp:256-262): I = lower gid at its own position, J = higher gid at pos + k*boxLen
__device__ __forceinline__ void canonicalPair(const PairIn &in, int &si, int &sj, int &code) {
    if (in.sGid[si] > in.sGid[sj]) { // swap roles; relative image flips sign
        int kx, ky, kz;
        imageOf(code, kx, ky, kz);
        const int t = si; si = sj; sj = t;
        code = (1 - kx) + 3 * (1 - ky) + 9 * (1 - kz);
    }
}

// Exact narrow phase, one staged candidate per thread (grid-stride; the count is read on the device).  A contact sets
// bit `seq` of its cell's bitmap: the order of a cell's contacts is the order of its candidates, as in k_pairs_find.
__global__ void __launch_bounds__(128)
k_cand_narrow(const unsigned long long *__restrict__ counters, const int4 *__restrict__ cand, unsigned long long cap,
              PairIn in, Box box, double colBuf, unsigned *__restrict__ bits, int W) {
    const unsigned long long n = min(counters[2], cap);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const int4 rec = cand[t];
        int si = rec.x, sj = rec.y, code = rec.w & 31;
        const int seq = rec.w >> 5;
        canonicalPair(in, si, sj, code);
        int kx, ky, kz;
        imageOf(code, kx, ky, kz);
        RodGeom a = loadRod(in, si), b = loadRod(in, sj);
        b.c = v3(b.c.x + kx * box.len[0], b.c.y + ky * box.len[1], b.c.z + kz * box.len[2]);
        Contact ct;
        if (in.sGid[si] != in.sGid[sj] && pairContact(a, b, colBuf, ct) && seq < 32 * W) atomicOr(&bits[(size_t)rec.z * W + (seq >> 5)], 1u << (seq & 31));
    }
}

// per cell: contacts = set bits; the bitmap words get their exclusive prefix counts (rank of a contact = prefix + popc)
__global__ void k_cell_hit_count(int ncell, const int *__restrict__ cellCand, const unsigned *__restrict__ bits, int W,
                                 int *__restrict__ prefix, int *__restrict__ cellHits) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    const int nw = min(W, (cellCand[cell] + 31) >> 5);
    int run = 0;
    for (int w = 0; w < nw; w++) {
        prefix[(size_t)cell * W + w] = run;
        run += __popc(bits[(size_t)cell * W + w]);
    }
    cellHits[cell] = run;
}

// emission for the split search: every staged candidate looks up its bit; a contact recomputes its geometry (hits only)
// and writes the constraint at cellHitStart[cell] + rank
__global__ void __launch_bounds__(128, 5)
k_pairs_emit2(unsigned long long n, const int4 *__restrict__ cand, const unsigned *__restrict__ bits,
              const int *__restrict__ prefix, int W, const int *__restrict__ cellHitStart, PairIn in, PairOut out, Box box,
              double colBuf) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    {
        const int4 rec = cand[t];
        const int seq = rec.w >> 5;
        if (seq >= 32 * W) return;
        const size_t wi = (size_t)rec.z * W + (seq >> 5);
        const unsigned word = bits[wi];
        if (!((word >> (seq & 31)) & 1u)) return;
        const size_t k = (size_t)cellHitStart[rec.z] + (size_t)(prefix[wi] + __popc(word & ((1u << (seq & 31)) - 1u)));
        int si = rec.x, sj = rec.y, code = rec.w & 31;
        canonicalPair(in, si, sj, code);
        int kx, ky, kz;
        imageOf(code, kx, ky, kz);
        RodGeom a = loadRod(in, si), b = loadRod(in, sj);
        b.c = v3(b.c.x + kx * box.len[0], b.c.y + ky * box.len[1], b.c.z + kz * box.len[2]);
        Contact ct;
        pairContact(a, b, colBuf, ct); // same code, same inputs as in k_cand_narrow: a hit
        const size_t S = out.stride;
        out.idxI[k] = si;
        out.idxJ[k] = sj;
        out.gidI[k] = in.sGid[si];
        out.gidJ[k] = in.sGid[sj];
        out.shift[k] = (signed char)code;
        out.bi[k] = 0;
        out.oneSide[k] = 0;
        out.own[k] = in.sGhost[si] ? 0 : 1; // the owner of rod I counts the row in global reductions
        out.delta0[k] = ct.sep;
        out.gamma0[k] = ct.sep < 0 ? -ct.sep : 0;
        out.invKappa[k] = 0;
        out.kappa[k] = 0;
        out.n[k] = ct.normI.x; out.n[k + S] = ct.normI.y; out.n[k + 2 * S] = ct.normI.z;
        out.pI[k] = ct.posI.x; out.pI[k + S] = ct.posI.y; out.pI[k + 2 * S] = ct.posI.z;
        out.pJ[k] = ct.posJ.x; out.pJ[k + S] = ct.posJ.y; out.pJ[k + 2 * S] = ct.posJ.z;
        out.labI[k] = ct.labI.x; out.labI[k + S] = ct.labI.y; out.labI[k + 2 * S] = ct.labI.z;
        out.labJ[k] = ct.labJ.x; out.labJ[k + S] = ct.labJ.y; out.labJ[k + 2 * S] = ct.labJ.z;
    }
}

// exclusive scan of n ints into out[0..n] (out[n] = total)
void launchScanInt(Context &c, const int *in, int *out, int n) {
    cudaStream_t st = c.stream;
    if (n <= 4 * kScanTile) {
        k_scan_int<<<1, 1024, 0, st>>>(in, out, n);
        c.launches++;
        return;
    }
    const int tiles = (n + kScanTile - 1) / kScanTile;
    c.scanTmp.reserve(2 * (size_t)tiles + 4);
    int *tileSum = c.scanTmp.p, *tileOff = c.scanTmp.p + tiles + 1;
    k_scan_tile_sums<<<tiles, kScanThreads, 0, st>>>(in, n, tileSum);
    k_scan_int<<<1, 1024, 0, st>>>(tileSum, tileOff, tiles);
    k_scan_tile_apply<<<tiles, kScanThreads, 0, st>>>(in, out, n, tileOff);
    c.launches += 3;
}

// ------------------------------------------------------------------------------------------------
void ctxInit(Context &c) {
    ALENS_CUDA(cudaSetDevice(c.device));
    ALENS_CUDA(cudaDeviceGetAttribute(&c.numSMs, cudaDevAttrMultiProcessorCount, c.device));
    ALENS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.ownStream = true;
    {
        cudaMemPool_t pool;
        ALENS_CUDA(cudaDeviceGetDefaultMemPool(&pool, c.device));
        unsigned long long keep = ~0ULL; // grow-only buffers: never hand memory back to the driver
        ALENS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    g_allocStream = c.stream;
    g_allocAsync = true;
    for (auto &e : c.ev) ALENS_CUDA(cudaEventCreate(&e));
    c.dScal.reserve(1);
    ALENS_CUDA(cudaMemset(c.dScal.p, 0, sizeof(SolverScalars)));
    ALENS_CUDA(cudaMallocHost((void **)&c.hScal, sizeof(SolverScalars)));
    memset(c.hScal, 0, sizeof(SolverScalars));
    if (cudaHostAlloc((void **)&c.hProg, 64, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer((void **)&c.hProgDev, c.hProg, 0) == cudaSuccess) {
        c.hProg[0] = c.hProg[1] = 0;
    } else { // no mapped host memory: the solver falls back to batched launches + stream synchronisation
        cudaGetLastError();
        c.hProg = c.hProgDev = nullptr;
    }
    c.dCounters.reserve(4);
}

void ctxFree(Context &c) {
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto &e : c.ev)
        if (e) cudaEventDestroy(e);
    if (c.hScal) cudaFreeHost(c.hScal);
    if (c.hProg) cudaFreeHost(c.hProg);
    if (c.copyStream) cudaStreamDestroy(c.copyStream);
    if (c.evVelNC) cudaEventDestroy(c.evVelNC);
    if (c.evMain) cudaEventDestroy(c.evMain);
    if (c.ownStream && c.stream) cudaStreamDestroy(c.stream);
}

static void chooseGrid(Context &c, double maxR) {
    CellGrid &g = c.grid;
    g.cutoff = (2 * maxR + c.colBuf) * (1.0 + 1e-9);
    if (!(g.cutoff > 0)) g.cutoff = 1.0;
    const bool multi = c.comm.active;
    g.axis = multi ? c.slabAxis : -1;
    g.axisLen = multi ? c.box.len[c.slabAxis] : 0.0;
    double ext[3];
    long long total = 1;
    for (int k = 0; k < 3; k++) {
        ext[k] = c.box.len[k];
        g.lo[k] = c.box.lo[k];
        g.per[k] = c.box.pbc[k];
        if (multi && k == c.slabAxis) { // the slab plus a ghost layer on both sides; images come as ghost rods
            ext[k] = (c.slabHi - c.slabLo) + 2 * c.ghostWidth;
            g.lo[k] = c.slabLo - c.ghostWidth;
            g.per[k] = 0;
        }
        double m = std::floor(ext[k] / g.cutoff);
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
    for (int k = 0; k < 3; k++) g.inv[k] = ext[k] > 0 ? g.n[k] / ext[k] : 0.0;
}

// host: max bounding radius; called with the host arrays at upload time
double hostMaxRadius(int n, const double *len, const double *rad, double lRatio, double dRatio, double *meanOut) {
    double m = 0, sum = 0;
#pragma omp parallel for schedule(static) reduction(max : m) reduction(+ : sum)
    for (int i = 0; i < n; i++) {
        const double R = 0.5 * len[i] * lRatio + rad[i] * dRatio;
        if (R > m) m = R;
        sum += R;
    }
    if (meanOut) *meanOut = n > 0 ? sum / n : 0.0;
    return m;
}

// device: the same maximum over the resident rods (positive doubles order like their bit patterns)
__global__ void k_max_radius(int n, const double *__restrict__ len, const double *__restrict__ rad, double lRatio,
                             double dRatio, unsigned long long *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double R = 0;
    if (i < n) R = 0.5 * len[i] * lRatio + rad[i] * dRatio;
    for (int o = 16; o; o >>= 1) R = fmax(R, __shfl_xor_sync(0xffffffffu, R, o));
    if ((threadIdx.x & 31) == 0 && R > 0) atomicMax(out, (unsigned long long)__double_as_longlong(R));
}
thread_local cudaStream_t g_allocStream = nullptr;
thread_local bool g_allocAsync = false;

// applyBoxBC on the resident owned rods (periodic axes only)
void wrapRodPositions(Context &c) {
    if (c.nLocal > 0) k_rod_wrap<<<gridFor(c.nLocal, 256), 256, 0, c.stream>>>(c.nLocal, c.uPos.p, c.box);
}

void rodsUploaded(Context &c, bool wrap) {
    cudaStream_t st = c.stream;
    const bool multi = c.comm.active;
    c.nRods = c.nLocal;
    c.nGhost = 0;
    {
        const size_t NL = (size_t)c.nLocal;
        c.uImg.reserve(NL + 1);
        c.uGlobalIdx.reserve(NL + 1);
        ALENS_CUDA(cudaMemsetAsync(c.uImg.p, 0, NL + 1, st));
        if (c.nLocal > 0) k_global_index<<<gridFor(c.nLocal, 256), 256, 0, st>>>(c.nLocal, c.globalBase, c.uGlobalIdx.p);
    }
    if (c.maxRLRatio != c.lRatio || c.maxRDRatio != c.dRatio) {
        // the collision ratios changed after the upload (alens_set_collision_params + alens_prepare_step):
        // the cell size has to follow, or the 27-cell stencil would miss contacts
        ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, sizeof(unsigned long long), st));
        if (c.nLocal > 0)
            k_max_radius<<<gridFor(c.nLocal, 256), 256, 0, st>>>(c.nLocal, c.uLen.p, c.uRad.p, c.lRatio, c.dRatio,
                                                                 c.dCounters.p);
        unsigned long long bits = 0;
        ALENS_CUDA(cudaMemcpyAsync(&bits, c.dCounters.p, sizeof(bits), cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        memcpy(&c.maxRLocal, &bits, sizeof(double));
        c.meanRLocal = 0; // (unknown for the new ratios: no long-rod pass until the next alens_set_rods)
        c.maxRLRatio = c.lRatio;
        c.maxRDRatio = c.dRatio;
    }
    double maxR = c.maxRLocal;
    // Polydisperse rods: the cell edge follows shortR = min(maxR, long_rods x mean bounding radius); the few rods above it
    // ("long" rods) are paired with partners beyond the 27-cell stencil by a separate pass (collectLongRods).  One rank only.
    c.shortR = maxR;
    if (!multi && c.optLongRods > 0 && c.meanRLocal > 0 && maxR > c.optLongRods * c.meanRLocal) c.shortR = c.optLongRods * c.meanRLocal;
    if (multi) {
        // (the ghost layer is as wide as the longest rod of ALL ranks needs; the cells follow this rank's own mean)
        if (c.optLongRods > 0 && c.meanRLocal > 0 && c.maxRadiusGlobal > c.optLongRods * c.meanRLocal)
            c.shortR = c.optLongRods * c.meanRLocal;
        else c.shortR = c.maxRadiusGlobal;
        c.ghostWidth = (2 * c.maxRadiusGlobal + c.colBuf) * (1.0 + 1e-9) + c.skin;
        if (c.slabHi - c.slabLo < 2 * c.ghostWidth)
            throw ArgError{ALENS_ERR_ARG, "slab decomposition: a slab must be at least 2 x (cutoff + skin) wide"};
        // owned rods are wrapped first, then the neighbours' rods near my faces are appended as ghosts
        if (wrap && c.nLocal > 0) k_rod_wrap<<<gridFor(c.nLocal, 256), 256, 0, st>>>(c.nLocal, c.uPos.p, c.box);
        ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 4 * sizeof(unsigned long long), st));
        // an open slab axis: the first and the last slab extend to infinity (a rod may leave the box there, as on one rank)
        const bool openAxis = !c.box.pbc[c.slabAxis];
        const double ownLo = (openAxis && c.rank == 0) ? -INFINITY : c.slabLo;
        const double ownHi = (openAxis && c.rank == c.nranks - 1) ? INFINITY : c.slabHi;
        if (c.nLocal > 0)
            k_local_image<<<gridFor(c.nLocal, 256), 256, 0, st>>>(
                c.nLocal, c.uPos.p, c.slabAxis, ownLo, ownHi, c.skin, c.box.len[c.slabAxis],
                c.box.pbc[c.slabAxis], c.uImg.p, reinterpret_cast<int *>(c.dCounters.p));
        int strays = 0;
        ALENS_CUDA(cudaMemcpyAsync(&strays, c.dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        c.strays = strays;
        commExchangeGhosts(c); // sets nRods = nLocal + nGhost (collective; strays are reported afterwards)
        wrap = false;
        maxR = c.maxRadiusGlobal;
    }
    const int n = c.nRods;
    c.gridMaxR = maxR;
    chooseGrid(c, c.shortR);
    if (c.shortR < maxR) { // the long-rod pass needs at least 3 cells along a periodic axis (one image per neighbour cell)
        bool ok = true;
        for (int k = 0; k < 3; k++) ok = ok && (!c.grid.per[k] || c.grid.n[k] >= 3);
        if (!ok) {
            c.shortR = maxR;
            chooseGrid(c, maxR);
        }
    }
    const CellGrid g = c.grid;
    c.uCell.reserve(n);
    c.userToSorted.reserve(n);
    c.cellCount.reserve(g.ncell + 1);
    c.cellStart.reserve(g.ncell + 1);
    c.cellFill.reserve(g.ncell + 1);
    c.sUser.reserve(n); c.sGid.reserve(n);
    c.sX.reserve(n); c.sY.reserve(n); c.sZ.reserve(n);
    c.sDx.reserve(n + 2); c.sDy.reserve(n + 2); c.sDz.reserve(n + 2); // +2: 16-byte bulk copies over-read
    c.sLc.reserve(n); c.sRc.reserve(n); c.sLen.reserve(n); c.sRad.reserve(n);
    c.sImm.reserve(n); c.sGhost.reserve(n); c.sImg.reserve(n);
    c.bUx.reserve(n); c.bUy.reserve(n); c.bUz.reserve(n); c.bH.reserve(n); c.bRho.reserve(n);
    DevBuf<int> &order = c.incFill; // scratch (rebuilt later by setup)
    order.reserve(n + 1);
    ALENS_CUDA(cudaMemsetAsync(c.cellCount.p, 0, sizeof(int) * (g.ncell + 1), st));
    ALENS_CUDA(cudaMemsetAsync(c.cellFill.p, 0, sizeof(int) * (g.ncell + 1), st));
    if (n > 0) {
        k_rod_pack<<<gridFor(n, 256), 256, 0, st>>>(n, c.uPos.p, c.box, g, wrap ? 1 : 0, c.uCell.p, c.cellCount.p,
                                                    c.uImg.p);
        c.launches++;
    }
    launchScanInt(c, c.cellCount.p, c.cellStart.p, g.ncell);
    if (n > 0) {
        k_cell_scatter<<<gridFor(n, 256), 256, 0, st>>>(n, c.uCell.p, c.cellStart.p, c.cellFill.p, order.p);
        RodArrays a{c.uGid.p, c.uPos.p, c.uQuat.p, c.uLen.p, c.uRad.p, c.uImm.p, c.uImg.p, c.nLocal, c.sUser.p,
                    c.sGid.p, c.userToSorted.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p, c.sDz.p, c.sLc.p, c.sRc.p,
                    c.sLen.p, c.sRad.p, c.sImm.p, c.sGhost.p, c.sImg.p, c.bUx.p, c.bUy.p, c.bUz.p, c.bH.p, c.bRho.p};
        k_cell_order<<<gridFor((long long)g.ncell * 32, 128), 128, 0, st>>>(g.ncell, c.cellStart.p, order.p, a,
                                                                            c.dRatio, c.lRatio);
        c.launches += 2;
    }
    ALENS_CUDA(cudaGetLastError());
    if (multi) commExchangeGhostIndices(c);
    c.sorted = true;
    c.haveMob = false;
    c.haveSetup = false;
    c.haveSolution = false;
    c.nCon = c.nColl = 0;
    c.nOneSide = c.nBilateral = 0;
    c.hostBlocks.clear();
    if (multi && c.strays > 0)
        throw ArgError{ALENS_ERR_STATE, "a rod left its slab by more than the skin: redistribute the rods (alens_get_rod_state / alens_set_rods)"};
}

void reserveConstraints(Context &c, size_t n, bool keep) {
    if (n <= c.conCap) return;
    // SoA arrays use conCap as the component stride, so growth re-lays them out
    const size_t ncap = (n + n / 4 + 1024 + 31) & ~(size_t)31; // component stride: 16-byte aligned components, padded tail
    const size_t old = c.conCap;
    const size_t live = keep ? (size_t)c.nCon : 0;
    cudaStream_t st = c.stream;
    auto grow1 = [&](auto &buf, size_t comps) {
        using T = std::remove_pointer_t<decltype(buf.p)>;
        T *np = static_cast<T *>(devAlloc(ncap * comps * sizeof(T)));
        if (live && buf.p)
            for (size_t k = 0; k < comps; k++)
                ALENS_CUDA(cudaMemcpyAsync(np + k * ncap, buf.p + k * old, live * sizeof(T), cudaMemcpyDeviceToDevice,
                                           st));
        devFree(buf.p);
        buf.p = np;
        buf.cap = ncap * comps;
    };
    grow1(c.cIdxI, 1); grow1(c.cIdxJ, 1); grow1(c.cGidI, 1); grow1(c.cGidJ, 1);
    grow1(c.cN, 3); grow1(c.cPI, 3); grow1(c.cPJ, 3); grow1(c.cLabI, 3); grow1(c.cLabJ, 3);
    grow1(c.cDelta0, 1); grow1(c.cGamma0, 1); grow1(c.cInvKappa, 1); grow1(c.cKappa, 1);
    grow1(c.cBi, 1); grow1(c.cOneSide, 1); grow1(c.cShift, 1); grow1(c.cOwn, 1);
    c.conCap = ncap;
}


// ------------------------------------------------------------------------------------------------
// Long rods (polydisperse suspensions).  The cell edge is sized for rods up to a bounding radius shortR; the stencil search
// above finds every contact between rods whose cells are neighbours, whatever their length.  What it cannot see is a contact
// of a LONG rod (bounding radius > shortR) with a partner more than one cell away.  Those come from two extra passes, both
// exact from the start (closest-point query on every pair that survives a point / segment distance test):
//   k_long_cells  warp per long rod A: the cells overlapping A's box grown by shortR + r_A + colBuf hold the centres of all
//                 SHORT partners; cells within the stencil of A's own cell are skipped (already searched)
//   k_long_long   warp per long rod A: all long rods behind it in the list (few), nearest periodic image, pairs whose cells
//                 are neighbours skipped
// Count pass, scan, fill pass; rows go behind the stencil rows in (long rod, cell walk, partner) order: deterministic.
// One image per pair: rods must be shorter than half a periodic box edge (as the image code of a row assumes anyway).
__device__ __forceinline__ void emitRow(const PairOut &out, size_t k, int si, int sj, int code, const Contact &ct,
                                        const PairIn &in) {
    const size_t S = out.stride;
    out.idxI[k] = si;
    out.idxJ[k] = sj;
    out.gidI[k] = in.sGid[si];
    out.gidJ[k] = in.sGid[sj];
    out.shift[k] = (signed char)code;
    out.bi[k] = 0;
    out.oneSide[k] = 0;
    out.own[k] = in.sGhost[si] ? 0 : 1;
    out.delta0[k] = ct.sep;
    out.gamma0[k] = ct.sep < 0 ? -ct.sep : 0;
    out.invKappa[k] = 0;
    out.kappa[k] = 0;
    out.n[k] = ct.normI.x; out.n[k + S] = ct.normI.y; out.n[k + 2 * S] = ct.normI.z;
    out.pI[k] = ct.posI.x; out.pI[k + S] = ct.posI.y; out.pI[k + 2 * S] = ct.posI.z;
    out.pJ[k] = ct.posJ.x; out.pJ[k + S] = ct.posJ.y; out.pJ[k + 2 * S] = ct.posJ.z;
    out.labI[k] = ct.labI.x; out.labI[k + S] = ct.labI.y; out.labI[k + 2 * S] = ct.labI.z;
    out.labJ[k] = ct.labJ.x; out.labJ[k + S] = ct.labJ.y; out.labJ[k + 2 * S] = ct.labJ.z;
}
__global__ void k_long_flags(int n, const double *__restrict__ sLc, const double *__restrict__ sRc, double shortR,
                             int *__restrict__ flag) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) flag[s] = (0.5 * sLc[s] + sRc[s] > shortR) ? 1 : 0;
}
__global__ void k_long_list(int n, const int *__restrict__ flag, const int *__restrict__ scan, int *__restrict__ list) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n && flag[s]) list[scan[s]] = s;
}
// squared distance of point p to the segment c -+ h d (d unit)
__device__ __forceinline__ double pointSegDist2(Vec3 p, Vec3 c, Vec3 d, double h) {
    const Vec3 w = p - c;
    double t = w.x * d.x + w.y * d.y + w.z * d.z;
    t = fmin(fmax(t, -h), h);
    const double ex = w.x - t * d.x, ey = w.y - t * d.y, ez = w.z - t * d.z;
    return ex * ex + ey * ey + ez * ez;
}
// the exact test of long rod a against partner b seen through image (kx, ky, kz) of b; canonical roles as everywhere
__device__ __forceinline__ bool longPairHit(const PairIn &in, const Box &box, double colBuf, int a, int b, int kx, int ky, int kz,
                                            int &si, int &sj, int &code, Contact &ct) {
    const int ga = in.sGid[a], gb = in.sGid[b];
    if (ga == gb) return false; // equal gids never collide (SylinderNear.hpp:210,225)
    if (ga < gb) {
        si = a; sj = b;
    } else {
        si = b; sj = a;
        kx = -kx; ky = -ky; kz = -kz;
    }
    code = (kx + 1) + 3 * (ky + 1) + 9 * (kz + 1);
    RodGeom I = loadRod(in, si), J = loadRod(in, sj);
    J.c = v3(J.c.x + kx * box.len[0], J.c.y + ky * box.len[1], J.c.z + kz * box.len[2]);
    return pairContact(I, J, colBuf, ct);
}
struct LongIn {
    int nLong;
    const int *list;     // sorted indices of the long rods, ascending
    const int *flag;     // per sorted rod: 1 = long
    const int *cellOfUser, *sUser; // assigned cell of a sorted rod s = cellOfUser[sUser[s]]
    double shortR;
};
template <bool FILL>
__global__ void __launch_bounds__(128) k_long_cells(LongIn L, PairIn in, Box box, CellGrid g, double colBuf, int *__restrict__ counts,
                                                    const int *__restrict__ starts, long long base, PairOut out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= L.nLong) return;
    const int a = L.list[w];
    const RodGeom A = loadRod(in, a);
    const double hA = 0.5 * A.lc;
    const double reach = (L.shortR + A.rc + colBuf) * (1.0 + 1e-9);
    // slab decomposition: a ghost (or an owned rod that strayed across the periodic face) is binned, and seen, at its
    // APPARENT position along the slab axis: original coordinate + image x box length; the pair's image code is the
    // difference of the two images there (the slab axis of the grid is not periodic)
    const int imgA = g.axis >= 0 ? in.sImg[a] : 0;
    const bool ghostA = g.axis >= 0 && in.sGhost[a] != 0;
    double ctr[3] = {A.c.x, A.c.y, A.c.z};
    if (g.axis >= 0) ctr[g.axis] += imgA * g.axisLen;
    const Vec3 cA = v3(ctr[0], ctr[1], ctr[2]);
    const double dir[3] = {A.d.x, A.d.y, A.d.z};
    const int cellA = L.cellOfUser[L.sUser[a]];
    const int ca[3] = {cellA % g.n[0], (cellA / g.n[0]) % g.n[1], cellA / (g.n[0] * g.n[1])};
    int u0[3], u1[3];
    for (int k = 0; k < 3; k++) {
        const double e = fabs(hA * dir[k]) + reach;
        long long lo = (long long)floor((ctr[k] - e - g.lo[k]) * g.inv[k]), hi = (long long)floor((ctr[k] + e - g.lo[k]) * g.inv[k]);
        if (!g.per[k] || g.inv[k] <= 0) { // open axis: rods outside the grid sit in its boundary cells
            lo = lo < 0 ? 0 : (lo >= g.n[k] ? g.n[k] - 1 : lo);
            hi = hi < 0 ? 0 : (hi >= g.n[k] ? g.n[k] - 1 : hi);
        } else if (hi - lo + 1 > g.n[k]) { // never the same cell through two images
            lo = ca[k] - (g.n[k] - 1) / 2;
            hi = lo + g.n[k] - 1;
        }
        u0[k] = (int)lo;
        u1[k] = (int)hi;
    }
    long long pos = FILL ? base + starts[w] : 0;
    int cnt = 0;
    for (int uz = u0[2]; uz <= u1[2]; uz++)
        for (int uy = u0[1]; uy <= u1[1]; uy++)
            for (int ux = u0[0]; ux <= u1[0]; ux++) {
                if (abs(ux - ca[0]) <= 1 && abs(uy - ca[1]) <= 1 && abs(uz - ca[2]) <= 1) continue; // the stencil search has it
                const int u[3] = {ux, uy, uz};
                int cc[3], kk[3];
                for (int k = 0; k < 3; k++) {
                    int img = 0, ck = u[k];
                    if (ck < 0 || ck >= g.n[k]) { // (periodic axis: open axes were clamped above)
                        img = (int)floor((double)ck / g.n[k]);
                        ck -= img * g.n[k];
                    }
                    cc[k] = ck;
                    kk[k] = img;
                }
                const int cell = (cc[2] * g.n[1] + cc[1]) * g.n[0] + cc[0];
                const int jb = in.cellStart[cell], je = in.cellStart[cell + 1];
                const Vec3 shift = v3(kk[0] * box.len[0], kk[1] * box.len[1], kk[2] * box.len[2]);
                for (int j0 = jb; j0 < je; j0 += 32) {
                    const int b = j0 + lane;
                    bool hit = false;
                    int si = 0, sj = 0, code = 13;
                    Contact ct;
                    if (b < je && !L.flag[b] && !(ghostA && in.sGhost[b])) { // short partners only: long ones belong to k_long_long
                        double cbv[3] = {in.sX[b] + shift.x, in.sY[b] + shift.y, in.sZ[b] + shift.z};
                        int kb[3] = {kk[0], kk[1], kk[2]};
                        if (g.axis >= 0) {
                            const int imgB = in.sImg[b];
                            cbv[g.axis] += imgB * g.axisLen;
                            kb[g.axis] = imgB - imgA;
                        }
                        const double rb = 0.5 * in.sLc[b] + in.sRc[b] + A.rc + colBuf;
                        if (pointSegDist2(v3(cbv[0], cbv[1], cbv[2]), cA, A.d, hA) <= rb * rb * (1.0 + 1e-9))
                            hit = longPairHit(in, box, colBuf, a, b, kb[0], kb[1], kb[2], si, sj, code, ct);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (FILL && hit) emitRow(out, (size_t)(pos + cnt + __popc(m & ((1u << lane) - 1u))), si, sj, code, ct, in);
                    cnt += __popc(m);
                }
            }
    if (!FILL && lane == 0) counts[w] = cnt;
}
template <bool FILL>
__global__ void __launch_bounds__(128) k_long_long(LongIn L, PairIn in, Box box, CellGrid g, double colBuf, int *__restrict__ counts,
                                                   const int *__restrict__ starts, long long base, PairOut out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= L.nLong) return;
    const int a = L.list[w];
    const RodGeom A = loadRod(in, a);
    const double hA = 0.5 * A.lc;
    const int cellA = L.cellOfUser[L.sUser[a]];
    const int ca[3] = {cellA % g.n[0], (cellA / g.n[0]) % g.n[1], cellA / (g.n[0] * g.n[1])};
    const int imgA = g.axis >= 0 ? in.sImg[a] : 0;
    const bool ghostA = g.axis >= 0 && in.sGhost[a] != 0;
    double ctrA[3] = {A.c.x, A.c.y, A.c.z};
    if (g.axis >= 0) ctrA[g.axis] += imgA * g.axisLen; // apparent position (see k_long_cells)
    const Vec3 cA = v3(ctrA[0], ctrA[1], ctrA[2]);
    long long pos = FILL ? base + starts[w] : 0;
    int cnt = 0;
    for (int l0 = w + 1; l0 < L.nLong; l0 += 32) {
        const int lb = l0 + lane;
        bool hit = false;
        int si = 0, sj = 0, code = 13;
        Contact ct;
        if (lb < L.nLong) {
            const int b = L.list[lb];
            double cb[3] = {in.sX[b], in.sY[b], in.sZ[b]};
            const double *ctr = ctrA;
            int kk[3] = {0, 0, 0};
            int kcell[3] = {0, 0, 0}; // image in units of the grid (the slab axis of the grid is not periodic)
            for (int k = 0; k < 3; k++)
                if (g.per[k] && box.len[k] > 0) { // nearest image of b
                    kk[k] = -(int)rint((cb[k] - ctr[k]) / box.len[k]);
                    cb[k] += kk[k] * box.len[k];
                    kcell[k] = kk[k];
                }
            if (g.axis >= 0) {
                const int imgB = in.sImg[b];
                cb[g.axis] += imgB * g.axisLen;
                kk[g.axis] = imgB - imgA;
            }
            const bool bothGhost = ghostA && in.sGhost[b];
            const int cellB = L.cellOfUser[L.sUser[b]];
            const int cbx = cellB % g.n[0] + kcell[0] * g.n[0], cby = (cellB / g.n[0]) % g.n[1] + kcell[1] * g.n[1],
                      cbz = cellB / (g.n[0] * g.n[1]) + kcell[2] * g.n[2];
            const bool stencil = abs(cbx - ca[0]) <= 1 && abs(cby - ca[1]) <= 1 && abs(cbz - ca[2]) <= 1;
            const double rb = 0.5 * in.sLc[b] + in.sRc[b] + A.rc + colBuf;
            if (!stencil && !bothGhost && pointSegDist2(v3(cb[0], cb[1], cb[2]), cA, A.d, hA) <= rb * rb * (1.0 + 1e-9))
                hit = longPairHit(in, box, colBuf, a, b, kk[0], kk[1], kk[2], si, sj, code, ct);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) emitRow(out, (size_t)(pos + cnt + __popc(m & ((1u << lane) - 1u))), si, sj, code, ct, in);
        cnt += __popc(m);
    }
    if (!FILL && lane == 0) counts[w] = cnt;
}

static PairIn pairIn(Context &c) {
    return PairIn{c.cellStart.p, c.sGid.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p, c.sDz.p, c.sLc.p, c.sRc.p,
                  c.bUx.p, c.bUy.p, c.bUz.p, c.bH.p, c.bRho.p, c.sImg.p, c.sGhost.p};
}


// the rows of the long rods (see k_long_cells); `total` stencil rows are in place.  Returns the number of rows added.
static long long collectLongRods(Context &c, long long total) {
    cudaStream_t st = c.stream;
    const int n = c.nRods;
    const CellGrid g = c.grid;
    DevBuf<int> flag, scan, list, cnt1, cnt2, st1, st2;
    flag.reserve((size_t)n + 8); scan.reserve((size_t)n + 8);
    k_long_flags<<<gridFor(n, 256), 256, 0, st>>>(n, c.sLc.p, c.sRc.p, c.shortR, flag.p);
    launchScanInt(c, flag.p, scan.p, n);
    int nLong = 0;
    ALENS_CUDA(cudaMemcpyAsync(&nLong, scan.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.launches += 1;
    c.nLongRods = nLong;
    c.nLongRows = 0;
    if (nLong == 0) return 0;
    list.reserve((size_t)nLong + 8); cnt1.reserve((size_t)nLong + 8); cnt2.reserve((size_t)nLong + 8);
    st1.reserve((size_t)nLong + 8); st2.reserve((size_t)nLong + 8);
    k_long_list<<<gridFor(n, 256), 256, 0, st>>>(n, flag.p, scan.p, list.p);
    const LongIn L{nLong, list.p, flag.p, c.uCell.p, c.sUser.p, c.shortR};
    auto pairOut = [&]() {
        return PairOut{c.cIdxI.p, c.cIdxJ.p, c.cGidI.p, c.cGidJ.p, c.cShift.p, c.cBi.p, c.cOneSide.p, c.cOwn.p, c.cDelta0.p,
                       c.cGamma0.p, c.cInvKappa.p, c.cKappa.p, c.cN.p, c.cPI.p, c.cPJ.p, c.cLabI.p, c.cLabJ.p, c.conCap};
    };
    const int grid = gridFor((long long)nLong * 32, 128);
    k_long_cells<false><<<grid, 128, 0, st>>>(L, pairIn(c), c.box, g, c.colBuf, cnt1.p, nullptr, 0, pairOut());
    k_long_long<false><<<grid, 128, 0, st>>>(L, pairIn(c), c.box, g, c.colBuf, cnt2.p, nullptr, 0, pairOut());
    launchScanInt(c, cnt1.p, st1.p, nLong);
    launchScanInt(c, cnt2.p, st2.p, nLong);
    int t1 = 0, t2 = 0;
    ALENS_CUDA(cudaMemcpyAsync(&t1, st1.p + nLong, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(&t2, st2.p + nLong, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.launches += 3;
    const long long extra = (long long)t1 + t2;
    if (total + extra > 0x7fffffffLL) throw ArgError{ALENS_ERR_UNSUPPORTED, "collect: more than 2^31 constraints on one GPU"};
    if (extra > 0) {
        c.nCon = total; // (rows to keep when the arrays are re-laid out)
        reserveConstraints(c, (size_t)(total + extra), true);
        if (t1 > 0) k_long_cells<true><<<grid, 128, 0, st>>>(L, pairIn(c), c.box, g, c.colBuf, nullptr, st1.p, total, pairOut());
        if (t2 > 0) k_long_long<true><<<grid, 128, 0, st>>>(L, pairIn(c), c.box, g, c.colBuf, nullptr, st2.p, total + t1, pairOut());
        c.launches += 2;
    }
    ALENS_CUDA(cudaGetLastError());
    c.nLongRows = extra;
    return extra;
}

void collectPairs(Context &c) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_collect_pair_collision: call alens_set_rods first"};
    cudaStream_t st = c.stream;
    const CellGrid g = c.grid;
    c.nCon = c.nColl = 0;
    c.nOneSide = c.nBilateral = 0;
    c.hostBlocks.clear();
    c.haveSetup = false;
    c.haveSolution = false;
    c.cellHits.reserve(g.ncell + 1);
    c.cellHitStart.reserve(g.ncell + 1);
    c.hitList.reserve(4 * (size_t)std::max(c.nRods, 256));
    const int ctas = gridFor(g.ncell, kWarpsPerCta);
    long long total = 0;
    auto pairOut = [&]() {
        return PairOut{c.cIdxI.p, c.cIdxJ.p, c.cGidI.p, c.cGidJ.p, c.cShift.p, c.cBi.p, c.cOneSide.p, c.cOwn.p, c.cDelta0.p,
                       c.cGamma0.p, c.cInvKappa.p, c.cKappa.p, c.cN.p, c.cPI.p, c.cPJ.p, c.cLabI.p, c.cLabJ.p, c.conCap};
    };
    // ---- split search (default): stages 1-2 at high occupancy -> candidates; exact query and ordered emission in dense
    // one-candidate-per-thread kernels.  Contacts keep the order of their candidates inside a cell (a bitmap of W words
    // per cell carries the hit flags), so the constraint list is the one the single-kernel search produces, row for row.
    bool done = false;
    if (c.optFindSplit) {
        int W = std::max(c.candWords, 4);
        const size_t bitmapBytes = (size_t)g.ncell * W * 8;
        if (bitmapBytes <= ((size_t)1 << 30)) {
            c.candBits.reserve((size_t)g.ncell * W + 4);
            c.candPrefix.reserve((size_t)g.ncell * W + 4);
            c.hitList.reserve(std::max<size_t>(8 * (size_t)std::max(c.nRods, 256), (size_t)(1.3 * c.lastCand) + 1024));
            ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 4 * sizeof(unsigned long long), st));
            ALENS_CUDA(cudaMemsetAsync(c.candBits.p, 0, sizeof(unsigned) * (size_t)g.ncell * W, st));
            if (c.optFindSplitMinB == 6)
                k_pairs_find<6, false><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), c.box, g, c.colBuf, c.cellHits.p,
                                                                           c.hitList.p, (unsigned long long)c.hitList.cap,
                                                                           c.dCounters.p);
            else if (g.axis < 0)
                k_pairs_find<8, false, false><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), c.box, g, c.colBuf, c.cellHits.p,
                                                                                  c.hitList.p, (unsigned long long)c.hitList.cap,
                                                                                  c.dCounters.p);
            else
                k_pairs_find<8, false><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), c.box, g, c.colBuf, c.cellHits.p,
                                                                           c.hitList.p, (unsigned long long)c.hitList.cap,
                                                                           c.dCounters.p);
            // grids of the dense kernels: sized from the last step's candidate count (grid-stride: any count works)
            const long long guess = std::max<long long>(c.lastCand + c.lastCand / 4, 1 << 16);
            const int gridDense = (int)std::min<long long>(gridFor(guess, 128), (long long)c.numSMs * 64);
            k_cand_narrow<<<gridDense, 128, 0, st>>>(c.dCounters.p, c.hitList.p, (unsigned long long)c.hitList.cap, pairIn(c),
                                                     c.box, c.colBuf, c.candBits.p, W);
            c.cellCand.reserve(g.ncell + 1);
            ALENS_CUDA(cudaMemcpyAsync(c.cellCand.p, c.cellHits.p, sizeof(int) * g.ncell, cudaMemcpyDeviceToDevice, st));
            k_cell_hit_count<<<gridFor(g.ncell, 128), 128, 0, st>>>(g.ncell, c.cellCand.p, c.candBits.p, W, c.candPrefix.p,
                                                                    c.cellHits.p);
            c.launches += 3;
            unsigned long long cnt[4];
            ALENS_CUDA(cudaMemcpyAsync(cnt, c.dCounters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
            launchScanInt(c, c.cellHits.p, c.cellHitStart.p, g.ncell);
            int tot = 0;
            ALENS_CUDA(cudaMemcpyAsync(&tot, c.cellHitStart.p + g.ncell, sizeof(int), cudaMemcpyDeviceToHost, st));
            ALENS_CUDA(cudaStreamSynchronize(st));
            c.statCand = (long long)cnt[0];
            c.lastCand = (long long)cnt[2];
            // next step's bitmap width follows the fullest cell; this step is valid only if nothing overflowed
            const long long maxCell = (long long)cnt[3];
            c.candWords = (int)std::min<long long>(4096, std::max<long long>(4, (maxCell + maxCell / 2 + 63) / 32));
            if (cnt[2] <= (unsigned long long)c.hitList.cap && maxCell <= 32LL * W) {
                total = tot;
                if (total > 0x7fffffffLL)
                    throw ArgError{ALENS_ERR_UNSUPPORTED, "collect: more than 2^31 constraints on one GPU"};
                reserveConstraints(c, (size_t)total, false);
                if (total > 0) {
                    k_pairs_emit2<<<gridFor((long long)cnt[2], 128), 128, 0, st>>>(cnt[2], c.hitList.p, c.candBits.p,
                                                                                   c.candPrefix.p, W, c.cellHitStart.p,
                                                                                   pairIn(c), pairOut(), c.box, c.colBuf);
                    c.launches++;
                }
                done = true;
            } // else: a cell or the list overflowed -- the single-kernel search below redoes this step
        }
    }
    for (int attempt = 0; attempt < 2 && !done; attempt++) {
        ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 4 * sizeof(unsigned long long), st));
        if (c.optFindMinB == 5)
            k_pairs_find<5, true><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), c.box, g, c.colBuf, c.cellHits.p, c.hitList.p,
                                                                      (unsigned long long)c.hitList.cap, c.dCounters.p);
        else if (c.optFindMinB == 3)
            k_pairs_find<3, true><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), c.box, g, c.colBuf, c.cellHits.p, c.hitList.p,
                                                                      (unsigned long long)c.hitList.cap, c.dCounters.p);
        else
            k_pairs_find<4, true><<<ctas, kWarpsPerCta * 32, 0, st>>>(pairIn(c), c.box, g, c.colBuf, c.cellHits.p, c.hitList.p,
                                                                      (unsigned long long)c.hitList.cap, c.dCounters.p);
        c.launches++;
        unsigned long long cnt[2];
        ALENS_CUDA(cudaMemcpyAsync(cnt, c.dCounters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        launchScanInt(c, c.cellHits.p, c.cellHitStart.p, g.ncell); // overlaps with the readback
        ALENS_CUDA(cudaStreamSynchronize(st));
        c.statCand = (long long)cnt[0];
        total = (long long)cnt[1];
        if ((size_t)total <= c.hitList.cap) break;
        if (attempt == 1) throw ArgError{ALENS_ERR_STATE, "collect: hit list overflow after regrowth"};
        c.hitList.reserve((size_t)total + (size_t)total / 8); // staged records were dropped: run the search again
    }
    if (total > 0x7fffffffLL) throw ArgError{ALENS_ERR_UNSUPPORTED, "collect: more than 2^31 constraints on one GPU"};
    if (!done) {
        reserveConstraints(c, (size_t)total, false);
        if (total > 0) {
            k_pairs_emit<<<gridFor(total, 128), 128, 0, st>>>(total, c.hitList.p, c.cellHitStart.p, pairIn(c), pairOut(),
                                                              c.box, c.colBuf);
            c.launches++;
        }
    }

    ALENS_CUDA(cudaGetLastError());
    c.nLongRods = c.nLongRows = 0;
    if (c.shortR < c.gridMaxR) total += collectLongRods(c, total);
    c.nCon = c.nColl = total;
}

// Force the (lazily loaded) kernels of this file into the context now: loading a kernel at its first launch can
// ------------------------------------------------------------------------------------------------
// Two-species short-range search on the rods' cell list: MixPairInteraction<...>::computeForce
// (SimToolbox/MPI/MixPairInteraction.hpp:148-311: an FDPS Symmetry tree over targets + sources) for TARGET points the
// caller hands in (protein ends, ...) against the resident rods as SOURCES.  For target t every rod j (and every periodic
// image of it) with |x_t - x_j| <= max(rs_t, rs_j) -- the distance the reference's search guarantees -- is reported, in
// (target, cell walk) order; the caller's functor does the fine test, as in the reference.  Thread per target, count pass +
// fill pass; the walk covers ceil(rsMax / cell edge) cells per direction.
struct MixIn {
    long long nTrg;
    const double *tPos, *tRs; // [3 nTrg], [nTrg]
    const double *sX, *sY, *sZ, *sRs; // sorted rods
    const int *cellStart, *sUser;
    CellGrid g;
    Box box;
    int reach[3];
};
template <bool FILL>
__global__ void k_mix_search(MixIn in, long long *__restrict__ rowPtr, int *__restrict__ out, long long cap) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= in.nTrg) return;
    double x[3] = {in.tPos[3 * t], in.tPos[3 * t + 1], in.tPos[3 * t + 2]};
    int c[3];
    for (int k = 0; k < 3; k++) {
        if (in.box.pbc[k] && in.box.len[k] > 0) { // applyBoxBC on the mixed system (MixPairInteraction.hpp:268)
            while (x[k] < in.box.lo[k]) x[k] += in.box.len[k];
            while (x[k] >= in.box.hi[k]) x[k] -= in.box.len[k];
        }
        int ci = (int)floor((x[k] - in.g.lo[k]) * in.g.inv[k]);
        c[k] = ci < 0 ? 0 : (ci >= in.g.n[k] ? in.g.n[k] - 1 : ci);
    }
    const double rt = in.tRs[t];
    long long pos = FILL ? rowPtr[t] : 0, cnt = 0;
    for (int dz = -in.reach[2]; dz <= in.reach[2]; dz++)
        for (int dy = -in.reach[1]; dy <= in.reach[1]; dy++)
            for (int dx = -in.reach[0]; dx <= in.reach[0]; dx++) {
                const int d[3] = {dx, dy, dz};
                int cc[3];
                double shift[3];
                bool ok = true;
                for (int k = 0; k < 3; k++) {
                    int ck = c[k] + d[k];
                    int img = 0;
                    if (ck < 0 || ck >= in.g.n[k]) {
                        if (!in.g.per[k]) { ok = false; break; }
                        img = (int)floor((double)ck / in.g.n[k]);
                        ck -= img * in.g.n[k];
                    }
                    cc[k] = ck;
                    shift[k] = img * in.box.len[k];
                }
                if (!ok) continue;
                const int cell = (cc[2] * in.g.n[1] + cc[1]) * in.g.n[0] + cc[0];
                for (int s = in.cellStart[cell]; s < in.cellStart[cell + 1]; s++) {
                    const double ex = in.sX[s] + shift[0] - x[0], ey = in.sY[s] + shift[1] - x[1], ez = in.sZ[s] + shift[2] - x[2];
                    const double rr = fmax(rt, in.sRs[s]);
                    if (ex * ex + ey * ey + ez * ez <= rr * rr) {
                        if (FILL && pos + cnt < cap) out[pos + cnt] = in.sUser[s];
                        cnt++;
                    }
                }
            }
    if (!FILL) rowPtr[t] = cnt;
}
__global__ void k_mix_src_radius(int n, const int *__restrict__ sUser, const double *__restrict__ userRs, const double *__restrict__ sLen,
                                 const double *__restrict__ sRad, const double *__restrict__ sLc, const double *__restrict__ sRc,
                                 double colBuf, double *__restrict__ sRs) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    // SylinderNearEP::getRSearch (SylinderNear.hpp:108-113) unless the caller brings its own radii
    sRs[s] = userRs ? userRs[sUser[s]] : 0.5 * fmax(sLen[s] + 2 * sRad[s], sLc[s] + 2 * sRc[s]) + colBuf;
}
long long mixPairSearch(Context &c, long long nTrg, const double *trgPos, const double *trgRs, const double *srcRs,
                        long long *rowPtr, int *srcIdx, long long cap) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_mix_pair_search: call alens_set_rods first"};
    if (c.comm.active) throw ArgError{ALENS_ERR_UNSUPPORTED, "alens_mix_pair_search: single-rank entry"};
    if (nTrg < 0 || (nTrg > 0 && (!trgPos || !trgRs || !rowPtr))) throw ArgError{ALENS_ERR_ARG, "alens_mix_pair_search: null input"};
    cudaStream_t st = c.stream;
    const int n = c.nRods;
    DevBuf<double> dPos, dRs, dSrcUser, dSrcRs;
    DevBuf<long long> dRow;
    DevBuf<int> dOut;
    dPos.reserve(3 * (size_t)nTrg + 3); dRs.reserve((size_t)nTrg + 1); dRow.reserve((size_t)nTrg + 2); dSrcRs.reserve((size_t)n + 1);
    double rsMax = 0;
    for (long long t = 0; t < nTrg; t++) rsMax = std::max(rsMax, trgRs[t]);
    if (srcRs) {
        dSrcUser.reserve((size_t)n + 1);
        ALENS_CUDA(cudaMemcpyAsync(dSrcUser.p, srcRs, 8 * (size_t)c.nLocal, cudaMemcpyHostToDevice, st));
        for (int i = 0; i < c.nLocal; i++) rsMax = std::max(rsMax, srcRs[i]);
    } else {
        rsMax = std::max(rsMax, c.maxRLocal + c.colBuf); // getRSearch <= lengthCollision/2 + radiusCollision + colBuf for ratios >= 1
        rsMax = std::max(rsMax, c.maxRLocal / std::min(std::min(c.lRatio, c.dRatio), 1.0) + c.colBuf);
    }
    if (nTrg > 0) {
        ALENS_CUDA(cudaMemcpyAsync(dPos.p, trgPos, 24 * (size_t)nTrg, cudaMemcpyHostToDevice, st));
        ALENS_CUDA(cudaMemcpyAsync(dRs.p, trgRs, 8 * (size_t)nTrg, cudaMemcpyHostToDevice, st));
    }
    if (n > 0)
        k_mix_src_radius<<<gridFor(n, 256), 256, 0, st>>>(n, c.sUser.p, srcRs ? dSrcUser.p : nullptr, c.sLen.p, c.sRad.p, c.sLc.p,
                                                          c.sRc.p, c.colBuf, dSrcRs.p);
    MixIn in{nTrg, dPos.p, dRs.p, c.sX.p, c.sY.p, c.sZ.p, dSrcRs.p, c.cellStart.p, c.sUser.p, c.grid, c.box, {0, 0, 0}};
    for (int k = 0; k < 3; k++) {
        const double edge = c.grid.inv[k] > 0 ? 1.0 / c.grid.inv[k] : 1e300;
        in.reach[k] = (int)std::ceil(rsMax / edge);
        if (in.reach[k] > 64) throw ArgError{ALENS_ERR_UNSUPPORTED, "alens_mix_pair_search: search radius spans more than 64 cells"};
    }
    long long total = 0;
    if (nTrg > 0) {
        k_mix_search<false><<<gridFor(nTrg, 128), 128, 0, st>>>(in, dRow.p, nullptr, 0);
        // the row pointer is a host output: counts come back, the exclusive sum runs here, the offsets go down again
        ALENS_CUDA(cudaMemcpyAsync(rowPtr, dRow.p, 8 * (size_t)nTrg, cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        for (long long t = 0; t < nTrg; t++) {
            const long long cnt = rowPtr[t];
            rowPtr[t] = total;
            total += cnt;
        }
        rowPtr[nTrg] = total;
        c.launches += 2;
        if (total > 0 && srcIdx && cap >= total) {
            dOut.reserve((size_t)total);
            ALENS_CUDA(cudaMemcpyAsync(dRow.p, rowPtr, 8 * (size_t)nTrg, cudaMemcpyHostToDevice, st));
            k_mix_search<true><<<gridFor(nTrg, 128), 128, 0, st>>>(in, dRow.p, dOut.p, total);
            c.launches++;
            ALENS_CUDA(cudaMemcpyAsync(srcIdx, dOut.p, 4 * (size_t)total, cudaMemcpyDeviceToHost, st));
            ALENS_CUDA(cudaStreamSynchronize(st));
        }
    } else if (rowPtr) {
        rowPtr[0] = 0;
    }
    ALENS_CUDA(cudaGetLastError());
    return total;
}

// synchronise the context, which deadlocks against a peer rank's waiting kernel when two ranks share one GPU.
void preloadCollideKernels() {
    cudaFuncAttributes a;
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_rod_pack));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_rod_wrap));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_mix_search<false>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_long_flags));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_long_list));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_long_cells<false>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_long_cells<true>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_long_long<false>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_long_long<true>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_mix_search<true>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_global_index));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_local_image));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_scan_int));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_scan_tile_sums));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_scan_tile_apply));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_cell_scatter));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_cell_order));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_pairs_find<4, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_pairs_find<5, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_pairs_find<3, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_pairs_find<8, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_pairs_find<8, false, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_pairs_find<6, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_cand_narrow));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_cell_hit_count));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_pairs_emit2));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_pairs_emit));
}

} // namespace alens
