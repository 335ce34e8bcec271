// blocks.cu -- ConstraintBlock traffic across the boundary: host-generated blocks pushed into the pool
// (boundary / link / protein constraints) and the refill of a host pool for output and stress.
//
// Reference: SimToolbox/Constraint/ConstraintBlock.hpp:30-127 (record), ConstraintCollector::writeBackGamma
// (Constraint/ConstraintCollector.cpp:439-461), CalcSylinderNearForce::collideStress
// (Sylinder/SylinderNear.hpp:432-519).  Compiled with -fmad=false (geometry.cuh).
#include "context.hpp"
#include "geometry.cuh"

#include <climits>
#include <cstring>
#include <string>
#include <vector>

namespace alens {

static_assert(sizeof(alens_constraint_block) == 272, "ConstraintBlock layout");

struct AppendIn {
    const int *uI, *uJ, *gidI, *gidJ;
    const unsigned char *oneSide, *bi, *own;
    const double *delta0, *gamma, *kappa;
    const double *vec; // [15][n]: n(3) pI(3) pJ(3) labI(3) labJ(3)
    long long n;
};
struct ConOut {
    int *idxI, *idxJ, *gidI, *gidJ;
    signed char *shift;
    unsigned char *bi, *oneSide, *own;
    double *delta0, *gamma0, *invKappa, *kappa;
    double *n, *pI, *pJ, *labI, *labJ;
    size_t stride;
};

__global__ void k_append(AppendIn in, ConOut o, const int *__restrict__ userToSorted, long long base) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    const size_t k = (size_t)(base + i), S = o.stride, N = (size_t)in.n;
    o.idxI[k] = userToSorted[in.uI[i]];
    o.idxJ[k] = in.oneSide[i] ? -1 : userToSorted[in.uJ[i]];
    o.gidI[k] = in.gidI[i];
    o.gidJ[k] = in.gidJ[i];
    o.shift[k] = 13;
    o.bi[k] = in.bi[i];
    o.oneSide[k] = in.oneSide[i];
    o.own[k] = in.own[i]; // 0: rod I is a ghost here, the row is counted by the rank that owns it
    o.delta0[k] = in.delta0[i];
    o.gamma0[k] = in.gamma[i];
    const double kap = in.kappa[i];
    o.kappa[k] = kap;
    o.invKappa[k] = (in.bi[i] && kap > 0) ? 1 / kap : 0.0; // ConstraintCollector.cpp:415-418
    for (int c = 0; c < 3; c++) {
        o.n[k + c * S] = in.vec[(0 + c) * N + i];
        o.pI[k + c * S] = in.vec[(3 + c) * N + i];
        o.pJ[k + c * S] = in.vec[(6 + c) * N + i];
        o.labI[k + c * S] = in.vec[(9 + c) * N + i];
        o.labJ[k + c * S] = in.vec[(12 + c) * N + i];
    }
}

// userIdx (optional): 2 n user indices (I, J) into this rank's rod arrays, ghosts included -- the device-side block
// generators know them; host-generated blocks are addressed by globalIndex and must refer to owned rods
void appendBlocks(Context &c, const alens_constraint_block *b, long long n, const int *userIdx) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_append_constraints: call alens_set_rods first"};
    if (n <= 0) return;
    const int off = c.globalBase; // rank offset of globalIndex (SylinderSystem.cpp:868-880)
    std::vector<int> uI(n), uJ(n), gI(n), gJ(n);
    std::vector<unsigned char> one(n), bi(n), own(n);
    std::vector<double> d0(n), gm(n), kp(n), vec(15 * (size_t)n);
    for (long long i = 0; i < n; i++) {
        const alens_constraint_block &q = b[i];
        const int li = userIdx ? userIdx[2 * i] : q.globalIndexI - off, lj = userIdx ? userIdx[2 * i + 1] : q.globalIndexJ - off;
        const int lim = userIdx ? c.nRods : c.nLocal;
        if (li < 0 || li >= lim || (!q.oneSide && (lj < 0 || lj >= lim)))
            throw ArgError{ALENS_ERR_ARG, "alens_append_constraints: globalIndex out of range"};
        own[i] = li < c.nLocal ? 1 : 0;
        if (!q.oneSide)
            for (int k = 0; k < 3; k++)
                if (q.normJ[k] != -q.normI[k])
                    throw ArgError{ALENS_ERR_UNSUPPORTED, "alens_append_constraints: two-sided block with normJ != -normI"};
        uI[i] = li; uJ[i] = q.oneSide ? li : lj;
        gI[i] = q.gidI; gJ[i] = q.gidJ;
        one[i] = q.oneSide ? 1 : 0; bi[i] = q.bilateral ? 1 : 0;
        d0[i] = q.delta0; gm[i] = q.gamma; kp[i] = q.kappa;
        for (int k = 0; k < 3; k++) {
            vec[(0 + k) * n + i] = q.normI[k];
            vec[(3 + k) * n + i] = q.posI[k];
            vec[(6 + k) * n + i] = q.posJ[k];
            vec[(9 + k) * n + i] = q.labI[k];
            vec[(12 + k) * n + i] = q.labJ[k];
        }
    }
    cudaStream_t st = c.stream;
    reserveConstraints(c, (size_t)(c.nCon + n), true);
    DevBuf<int> dI, dJ, dgI, dgJ;
    DevBuf<unsigned char> dOne, dBi, dOwn;
    DevBuf<double> dD0, dGm, dKp, dVec;
    dI.reserve(n); dJ.reserve(n); dgI.reserve(n); dgJ.reserve(n); dOne.reserve(n); dBi.reserve(n); dOwn.reserve(n);
    dD0.reserve(n); dGm.reserve(n); dKp.reserve(n); dVec.reserve(15 * (size_t)n);
    auto up = [&](void *d, const void *h, size_t bytes) {
        ALENS_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
    };
    up(dI.p, uI.data(), 4 * n); up(dJ.p, uJ.data(), 4 * n); up(dgI.p, gI.data(), 4 * n); up(dgJ.p, gJ.data(), 4 * n);
    up(dOne.p, one.data(), n); up(dBi.p, bi.data(), n); up(dOwn.p, own.data(), n);
    up(dD0.p, d0.data(), 8 * n); up(dGm.p, gm.data(), 8 * n); up(dKp.p, kp.data(), 8 * n);
    up(dVec.p, vec.data(), 8 * 15 * (size_t)n);
    AppendIn in{dI.p, dJ.p, dgI.p, dgJ.p, dOne.p, dBi.p, dOwn.p, dD0.p, dGm.p, dKp.p, dVec.p, n};
    ConOut o{c.cIdxI.p, c.cIdxJ.p, c.cGidI.p, c.cGidJ.p, c.cShift.p, c.cBi.p, c.cOneSide.p, c.cOwn.p, c.cDelta0.p,
             c.cGamma0.p, c.cInvKappa.p, c.cKappa.p, c.cN.p, c.cPI.p, c.cPJ.p, c.cLabI.p, c.cLabJ.p, c.conCap};
    k_append<<<gridFor(n, 256), 256, 0, st>>>(in, o, c.userToSorted.p, c.nCon);
    c.launches++;
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.hostBlocks.insert(c.hostBlocks.end(), b, b + n);
    for (long long i = 0; i < n; i++) {
        c.nOneSide += b[i].oneSide ? 1 : 0;
        c.nBilateral += b[i].bilateral ? 1 : 0;
    }
    c.nCon += n;
    c.haveSetup = false;
    c.haveSolution = false;
}

// ------------------------------------------------------------------------------------------------
// Boundary collisions (SylinderSystem::collectBoundaryCollision, SylinderSystem.cpp:1093-1150; Boundary::project,
// Boundary/Boundary.cpp:25-41, :108-124, :185-209).  Same expression order as the reference (and the oracle): with
// -fmad=false the blocks are bit-identical.
__device__ __forceinline__ void boundaryProject(const alens_boundary &b, Vec3 q, Vec3 &proj, Vec3 &delta) {
    const Vec3 ctr = v3(b.center[0], b.center[1], b.center[2]), ax = v3(b.axis[0], b.axis[1], b.axis[2]);
    if (b.type == 0) { // spherical shell
        const Vec3 Q = q - ctr;
        const double QueryR = norm(Q);
        const double f = b.radius * (1 / QueryR);
        const Vec3 P = v3(f * Q.x, f * Q.y, f * Q.z);
        Vec3 PQ = Q - P;
        const bool out = QueryR > b.radius;
        if ((b.inside && out) || (!b.inside && !out)) PQ = v3(PQ.x * -1, PQ.y * -1, PQ.z * -1);
        proj = P + ctr;
        delta = PQ;
    } else if (b.type == 1) { // flat wall
        const Vec3 CQ = q - ctr;
        const double t = dot(CQ, ax);
        const Vec3 P = v3(q.x - t * ax.x, q.y - t * ax.y, q.z - t * ax.z);
        Vec3 PQ = q - P;
        if (t < 0) PQ = v3(PQ.x * -1, PQ.y * -1, PQ.z * -1);
        proj = P;
        delta = PQ;
    } else { // infinite tube
        const Vec3 CQ = q - ctr;
        const double t = dot(CQ, ax);
        const Vec3 PA = v3(ctr.x + t * ax.x, ctr.y + t * ax.y, ctr.z + t * ax.z);
        const Vec3 PAQ = q - PA;
        const double r = norm(PAQ);
        const Vec3 P = v3(PA.x + b.radius * (PAQ.x / r), PA.y + b.radius * (PAQ.y / r), PA.z + b.radius * (PAQ.z / r));
        Vec3 d = q - P;
        if (r > b.radius) {
            if (b.inside) d = v3(d.x * -1, d.y * -1, d.z * -1);
        } else {
            if (!b.inside) d = v3(d.x * -1, d.y * -1, d.z * -1);
        }
        proj = P;
        delta = d;
    }
}

struct BoundaryRods {
    const int *userToSorted, *sGid;
    const double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLc, *sRc;
    int nLocal, globalBase;
    double colBuf;
};

// checkEnd (SylinderSystem.cpp:1111-1133); returns whether a block is due, fills it when blk != nullptr
__device__ __forceinline__ bool boundaryCheckEnd(const alens_boundary &b, Vec3 center, Vec3 Query, double radius,
                                                 double radiusCollision, double colBuf, int gid, int globalIndex,
                                                 alens_constraint_block *blk) {
    Vec3 Proj, delta;
    boundaryProject(b, Query, Proj, delta);
    const double deltanorm = norm(delta);
    const double inv = 1 / deltanorm;
    const Vec3 nrm = v3(delta.x * inv, delta.y * inv, delta.z * inv);
    const Vec3 posI = Query - center;
    double d0;
    if (dot(Query - Proj, delta) < 0) d0 = -deltanorm - radius;
    else if (deltanorm < (1 + colBuf * 2) * radiusCollision) d0 = deltanorm - radius;
    else return false;
    if (blk) {
        alens_constraint_block q;
        memset(&q, 0, sizeof(q));
        q.delta0 = d0;
        q.gamma = 0;
        q.gidI = q.gidJ = gid;
        q.globalIndexI = q.globalIndexJ = globalIndex;
        q.oneSide = 1;
        q.bilateral = 0;
        q.kappa = 0;
        q.normI[0] = q.normJ[0] = nrm.x; q.normI[1] = q.normJ[1] = nrm.y; q.normI[2] = q.normJ[2] = nrm.z;
        q.posI[0] = q.posJ[0] = posI.x; q.posI[1] = q.posJ[1] = posI.y; q.posI[2] = q.posJ[2] = posI.z;
        q.labI[0] = Query.x; q.labI[1] = Query.y; q.labI[2] = Query.z;
        q.labJ[0] = Proj.x; q.labJ[1] = Proj.y; q.labJ[2] = Proj.z;
        *blk = q;
    }
    return true;
}

// one thread per (boundary, local rod in the caller's order).  EMIT = false: number of blocks (0..2) -> cnt;
// EMIT = true: the blocks at out[start[...]], minus end first
template <bool EMIT>
__global__ void k_boundary(BoundaryRods R, const alens_boundary *__restrict__ bnd, int nb, int *__restrict__ cnt,
                           const int *__restrict__ start, alens_constraint_block *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nb * R.nLocal) return;
    const int ib = (int)(t / R.nLocal), u = (int)(t - (long long)ib * R.nLocal);
    const alens_boundary b = bnd[ib];
    const int s = R.userToSorted[u];
    const Vec3 c = v3(R.sX[s], R.sY[s], R.sZ[s]), d = v3(R.sDx[s], R.sDy[s], R.sDz[s]);
    const double lc = R.sLc[s], rc = R.sRc[s];
    const int gid = R.sGid[s], gi = R.globalBase + u;
    alens_constraint_block *o = EMIT ? out + start[t] : nullptr;
    int n = 0;
    if (lc < 2 * rc) { // sphere for collisions (SylinderNear.hpp:241; SylinderSystem.cpp:1135-1137)
        const double radius = lc * 0.5 + rc;
        if (boundaryCheckEnd(b, c, c, radius, rc, R.colBuf, gid, gi, EMIT ? o : nullptr)) n++;
    } else {
        const double h = lc * 0.5;
        const Vec3 Qm = v3(c.x - d.x * h, c.y - d.y * h, c.z - d.z * h), Qp = v3(c.x + d.x * h, c.y + d.y * h, c.z + d.z * h);
        if (boundaryCheckEnd(b, c, Qm, rc, rc, R.colBuf, gid, gi, EMIT ? o : nullptr)) n++;
        if (boundaryCheckEnd(b, c, Qp, rc, rc, R.colBuf, gid, gi, EMIT ? o + n : nullptr)) n++;
    }
    if (!EMIT) cnt[t] = n;
}

void launchScanInt(Context &c, const int *in, int *out, int n);

long long collectBoundary(Context &c, const alens_boundary *bnd, int nb) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_collect_boundary_collision: call alens_set_rods first"};
    if (nb <= 0 || c.nLocal == 0) return 0;
    cudaStream_t st = c.stream;
    std::vector<alens_boundary> hb(bnd, bnd + nb);
    for (auto &b : hb) { // the reference's constructors normalise the wall normal / tube axis (Boundary.cpp:96-100, :166-172)
        if (b.type < 0 || b.type > 2) throw ArgError{ALENS_ERR_ARG, "alens_collect_boundary_collision: type must be 0, 1 or 2"};
        const double a = std::sqrt(b.axis[0] * b.axis[0] + b.axis[1] * b.axis[1] + b.axis[2] * b.axis[2]);
        if (b.type != 0) {
            if (!(a > 0)) throw ArgError{ALENS_ERR_ARG, "alens_collect_boundary_collision: zero axis"};
            for (int k = 0; k < 3; k++) b.axis[k] = b.axis[k] / a;
        }
    }
    const long long nt = (long long)nb * c.nLocal;
    if (nt > 0x7fffffffLL - 8) throw ArgError{ALENS_ERR_UNSUPPORTED, "alens_collect_boundary_collision: too many (boundary, rod) pairs"};
    DevBuf<alens_boundary> dB;
    DevBuf<int> dCnt, dStart;
    dB.reserve(nb);
    dCnt.reserve((size_t)nt + 1);
    dStart.reserve((size_t)nt + 8);
    ALENS_CUDA(cudaMemcpyAsync(dB.p, hb.data(), sizeof(alens_boundary) * nb, cudaMemcpyHostToDevice, st));
    const BoundaryRods R{c.userToSorted.p, c.sGid.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p, c.sDz.p, c.sLc.p, c.sRc.p,
                         c.nLocal, c.globalBase, c.colBuf};
    k_boundary<false><<<gridFor(nt, 128), 128, 0, st>>>(R, dB.p, nb, dCnt.p, nullptr, nullptr);
    launchScanInt(c, dCnt.p, dStart.p, (int)nt);
    int total = 0;
    ALENS_CUDA(cudaMemcpyAsync(&total, dStart.p + nt, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.launches += 1;
    if (total == 0) return 0;
    DevBuf<alens_constraint_block> dOut;
    dOut.reserve((size_t)total);
    k_boundary<true><<<gridFor(nt, 128), 128, 0, st>>>(R, dB.p, nb, nullptr, dStart.p, dOut.p);
    c.launches += 1;
    std::vector<alens_constraint_block> host((size_t)total);
    ALENS_CUDA(cudaMemcpyAsync(host.data(), dOut.p, sizeof(alens_constraint_block) * (size_t)total, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    appendBlocks(c, host.data(), total); // into the constraint arrays + the host copies kept for the pool refill
    return total;
}

// ------------------------------------------------------------------------------------------------
// Bilateral links (SylinderSystem::collectLinkBilateral, SylinderSystem.cpp:1386-1482).  gid -> rod through an open
// addressing hash table built on the device (the reference asks its ZDD data directory).
static constexpr int kEmptyKey = INT_MIN;
__device__ __forceinline__ unsigned hashGid(int g) {
    unsigned x = (unsigned)g;
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__global__ void k_gid_table_build(int nLocal, const int *__restrict__ gid, int *__restrict__ keys, int *__restrict__ vals,
                                  unsigned mask) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nLocal) return;
    const int g = gid[u];
    unsigned h = hashGid(g) & mask;
    while (true) {
        const int prev = atomicCAS(&keys[h], kEmptyKey, g);
        if (prev == kEmptyKey || prev == g) {
            atomicMin(&vals[h], u); // a duplicated gid resolves to the first rod
            return;
        }
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ int gidLookup(int g, const int *__restrict__ keys, const int *__restrict__ vals, unsigned mask) {
    unsigned h = hashGid(g) & mask;
    while (true) {
        const int k = keys[h];
        if (k == g) return vals[h];
        if (k == kEmptyKey) return -1;
        h = (h + 1) & mask;
    }
}
// findPBCImage(lb, ub, x, trg) of Util/GeoUtil.hpp:27-60
__device__ __forceinline__ void pbcImage1(double lb, double ub, double &x) {
    const double L = ub - lb;
    while (x >= ub) x -= L;
    while (x < lb) x += L;
}
__device__ __forceinline__ void pbcImage2(double lb, double ub, double &x, double &trg) {
    pbcImage1(lb, ub, trg);
    double dist = x - trg;
    pbcImage1(0.0, ub - lb, dist);
    if (dist > (ub - lb) * 0.5) x = trg + dist - (ub - lb);
    else x = trg + dist;
}

struct LinkRods {
    const int *userToSorted, *sGid;
    const double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLen, *sRad;
    const int *uGlobalIdx; // global index of every rod of this rank, ghosts included (the owner's numbering)
    int nLocal;            // user indices >= nLocal are ghost rods (slab decomposition)
    int multi;
};
__global__ void k_links(long long nLinks, const int *__restrict__ prevGid, const int *__restrict__ nextGid,
                        const int *__restrict__ keys, const int *__restrict__ vals, unsigned mask, LinkRods R, Box box,
                        double linkKappa, double linkGap, alens_constraint_block *__restrict__ out, int *__restrict__ missing,
                        int2 *__restrict__ userIdx) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nLinks) return;
    const int uI = gidLookup(prevGid[l], keys, vals, mask), uJ = gidLookup(nextGid[l], keys, vals, mask);
    userIdx[l] = make_int2(-1, -1); // "not this rank's link"
    // Slab decomposition: every rank is handed the whole link map (as every rank of the reference reads it, :377-405).  A
    // link is this rank's business when it owns at least one of the two rods; the other one is then an owned rod or a
    // ghost (both ranks build the block from identical rod data, as for a contact across a slab face).
    const bool ownI = uI >= 0 && uI < R.nLocal, ownJ = uJ >= 0 && uJ < R.nLocal;
    if (R.multi && !ownI && !ownJ) return;
    if (uI < 0 || uJ < 0) { // an end is neither owned nor within the ghost layer (single rank: not there at all)
        atomicAdd(missing, 1);
        return;
    }
    userIdx[l] = make_int2(uI, uJ);
    const int sI = R.userToSorted[uI], sJ = R.userToSorted[uJ];
    const Vec3 cI = v3(R.sX[sI], R.sY[sI], R.sZ[sI]), dI = v3(R.sDx[sI], R.sDy[sI], R.sDz[sI]);
    const Vec3 dJ = v3(R.sDx[sJ], R.sDy[sJ], R.sDz[sJ]);
    double cj[3] = {R.sX[sJ], R.sY[sJ], R.sZ[sJ]};
    const double ci[3] = {cI.x, cI.y, cI.z};
    for (int k = 0; k < 3; k++) { // nearest periodic image of J (:1436-1448)
        if (!box.pbc[k]) continue;
        double trg = ci[k], xk = cj[k];
        pbcImage2(box.lo[k], box.hi[k], xk, trg);
        cj[k] = xk;
    }
    const Vec3 cJ = v3(cj[0], cj[1], cj[2]);
    const double lenI = R.sLen[sI], lenJ = R.sLen[sJ], radI = R.sRad[sI], radJ = R.sRad[sJ];
    const double hI = 0.5 * lenI, hJ = 0.5 * lenJ;
    const Vec3 Pp = v3(cI.x + dI.x * hI, cI.y + dI.y * hI, cI.z + dI.z * hI);
    const Vec3 Qm = v3(cJ.x - dJ.x * hJ, cJ.y - dJ.y * hJ, cJ.z - dJ.z * hJ);
    const Vec3 rvec = Qm - Pp;
    const double rnorm = norm(rvec);
    const double delta0 = rnorm - radI - radJ - linkGap;
    const Vec3 PQ = Pp - Qm;
    const double pqn = norm(PQ);
    const Vec3 nI = pqn > 0 ? v3(PQ.x / pqn, PQ.y / pqn, PQ.z / pqn) : PQ;
    alens_constraint_block q;
    memset(&q, 0, sizeof(q));
    q.delta0 = delta0;
    q.gamma = delta0 < 0 ? -delta0 : 0;
    q.gidI = R.sGid[sI];
    q.gidJ = R.sGid[sJ];
    q.globalIndexI = R.uGlobalIdx[uI];
    q.globalIndexJ = R.uGlobalIdx[uJ];
    q.oneSide = 0;
    q.bilateral = 1;
    q.kappa = linkKappa;
    q.normI[0] = nI.x; q.normI[1] = nI.y; q.normI[2] = nI.z;
    q.normJ[0] = -nI.x; q.normJ[1] = -nI.y; q.normJ[2] = -nI.z;
    q.posI[0] = Pp.x - cI.x; q.posI[1] = Pp.y - cI.y; q.posI[2] = Pp.z - cI.z;
    q.posJ[0] = Qm.x - cJ.x; q.posJ[1] = Qm.y - cJ.y; q.posJ[2] = Qm.z - cJ.z;
    q.labI[0] = Pp.x; q.labI[1] = Pp.y; q.labI[2] = Pp.z;
    q.labJ[0] = Qm.x; q.labJ[1] = Qm.y; q.labJ[2] = Qm.z;
    collideStress(dI, dJ, cI, cJ, lenI, lenJ, radI, radJ, 1.0, Pp, Qm, q.stress);
    out[l] = q;
}

long long collectLinks(Context &c, const int *prevGid, const int *nextGid, long long nLinks, double linkKappa,
                       double linkGap) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_collect_link_bilateral: call alens_set_rods first"};
    if (nLinks <= 0) return 0;
    if (!prevGid || !nextGid) throw ArgError{ALENS_ERR_ARG, "alens_collect_link_bilateral: NULL gid list"};
    cudaStream_t st = c.stream;
    unsigned size = 64;
    while (size < 2u * (unsigned)std::max(c.nRods, 1)) size <<= 1;
    DevBuf<int> keys, vals, dPrev, dNext, dMissing;
    keys.reserve(size); vals.reserve(size); dPrev.reserve((size_t)nLinks); dNext.reserve((size_t)nLinks); dMissing.reserve(1);
    {
        std::vector<int> fill(size, kEmptyKey);
        ALENS_CUDA(cudaMemcpyAsync(keys.p, fill.data(), sizeof(int) * size, cudaMemcpyHostToDevice, st));
        ALENS_CUDA(cudaMemsetAsync(vals.p, 0x7f, sizeof(int) * size, st)); // 0x7f7f7f7f: larger than any rod index
        ALENS_CUDA(cudaMemsetAsync(dMissing.p, 0, sizeof(int), st));
        ALENS_CUDA(cudaStreamSynchronize(st)); // `fill` is pageable host memory
    }
    ALENS_CUDA(cudaMemcpyAsync(dPrev.p, prevGid, sizeof(int) * (size_t)nLinks, cudaMemcpyHostToDevice, st));
    ALENS_CUDA(cudaMemcpyAsync(dNext.p, nextGid, sizeof(int) * (size_t)nLinks, cudaMemcpyHostToDevice, st));
    const int nTab = c.nRods; // owned rods and, with the slab decomposition, the ghost layer
    if (nTab > 0) k_gid_table_build<<<gridFor(nTab, 256), 256, 0, st>>>(nTab, c.uGid.p, keys.p, vals.p, size - 1);
    DevBuf<alens_constraint_block> dOut;
    DevBuf<int2> dIdx;
    dOut.reserve((size_t)nLinks);
    dIdx.reserve((size_t)nLinks);
    const LinkRods R{c.userToSorted.p, c.sGid.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p, c.sDz.p, c.sLen.p, c.sRad.p,
                     c.uGlobalIdx.p, c.nLocal, c.comm.active ? 1 : 0};
    k_links<<<gridFor(nLinks, 128), 128, 0, st>>>(nLinks, dPrev.p, dNext.p, keys.p, vals.p, size - 1, R, c.box, linkKappa,
                                                  linkGap, dOut.p, dMissing.p, dIdx.p);
    c.launches += 2;
    int missing = 0;
    std::vector<alens_constraint_block> host((size_t)nLinks);
    std::vector<int2> idx((size_t)nLinks);
    ALENS_CUDA(cudaMemcpyAsync(&missing, dMissing.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(host.data(), dOut.p, sizeof(alens_constraint_block) * (size_t)nLinks, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(idx.data(), dIdx.p, sizeof(int2) * (size_t)nLinks, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    if (missing)
        throw ArgError{ALENS_ERR_ARG, "alens_collect_link_bilateral: " + std::to_string(missing) +
                                          (c.comm.active ? " link(s) join an owned rod to a rod outside this rank's ghost layer"
                                                         : " link end(s) refer to a gid this rank does not own")};
    std::vector<int> user; // links this rank takes part in, in the caller's order
    user.reserve(2 * (size_t)nLinks);
    size_t m = 0;
    for (long long l = 0; l < nLinks; l++) {
        if (idx[l].x < 0) continue;
        host[m++] = host[l];
        user.push_back(idx[l].x);
        user.push_back(idx[l].y);
    }
    if (m) appendBlocks(c, host.data(), (long long)m, user.data());
    return (long long)m;
}

// ------------------------------------------------------------------------------------------------
// Protein (crosslinker / motor) bilateral constraints: TubuleSystem::setProteinConstraints (SRC/TubuleSystem.cpp:694-745).
// One thread per doubly bound protein: delta0 = forceLength - freeLength, gamma0 = -delta0 kappa, normI = (P - Q)/|P - Q|,
// posI = P - centerI, posJ = Q - centerJ, lab = (P, Q), bilateral with stiffness kappa, stress by collideStress with
// radius tubuleDiameter / 2 on both sides.  Singly bound / unbound proteins (an id < 0) produce nothing.
__global__ void k_protein_blocks(long long n, const alens_protein_bind *__restrict__ pr, double tubuleDiameter,
                                 alens_constraint_block *__restrict__ out, unsigned char *__restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const alens_protein_bind p = pr[i];
    alens_constraint_block q;
    memset(&q, 0, sizeof(q));
    const bool both = p.idBind[0] >= 0 && p.idBind[1] >= 0; // ID_UB = -1 (Protein/ProteinBindStatus.hpp)
    keep[i] = both ? 1 : 0;
    if (both) {
        const Vec3 cI = v3(p.centerBind[0][0], p.centerBind[0][1], p.centerBind[0][2]);
        const Vec3 cJ = v3(p.centerBind[1][0], p.centerBind[1][1], p.centerBind[1][2]);
        const Vec3 dI = v3(p.directionBind[0][0], p.directionBind[0][1], p.directionBind[0][2]);
        const Vec3 dJ = v3(p.directionBind[1][0], p.directionBind[1][1], p.directionBind[1][2]);
        const Vec3 P = v3(p.posEndBind[0][0], p.posEndBind[0][1], p.posEndBind[0][2]);
        const Vec3 Q = v3(p.posEndBind[1][0], p.posEndBind[1][1], p.posEndBind[1][2]);
        const double delta0 = p.forceLength - p.freeLength;
        const Vec3 PQ = P - Q;
        const double pqn = norm(PQ);
        const Vec3 nI = pqn > 0 ? v3(PQ.x / pqn, PQ.y / pqn, PQ.z / pqn) : PQ;
        q.delta0 = delta0;
        q.gamma = -delta0 * p.kappa;
        q.gidI = p.idBind[0]; q.gidJ = p.idBind[1];
        q.globalIndexI = p.indexBind[0]; q.globalIndexJ = p.indexBind[1];
        q.oneSide = 0;
        q.bilateral = 1;
        q.kappa = p.kappa;
        q.normI[0] = nI.x; q.normI[1] = nI.y; q.normI[2] = nI.z;
        q.normJ[0] = -nI.x; q.normJ[1] = -nI.y; q.normJ[2] = -nI.z;
        q.posI[0] = P.x - cI.x; q.posI[1] = P.y - cI.y; q.posI[2] = P.z - cI.z;
        q.posJ[0] = Q.x - cJ.x; q.posJ[1] = Q.y - cJ.y; q.posJ[2] = Q.z - cJ.z;
        q.labI[0] = P.x; q.labI[1] = P.y; q.labI[2] = P.z;
        q.labJ[0] = Q.x; q.labJ[1] = Q.y; q.labJ[2] = Q.z;
        collideStress(dI, dJ, cI, cJ, p.lenBind[0], p.lenBind[1], tubuleDiameter / 2, tubuleDiameter / 2, 1.0, P, Q, q.stress);
    }
    out[i] = q;
}

long long collectProteins(Context &c, const alens_protein_bind *proteins, long long n, double tubuleDiameter) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_collect_protein_bilateral: call alens_set_rods first"};
    if (n <= 0) return 0;
    if (!proteins) throw ArgError{ALENS_ERR_ARG, "alens_collect_protein_bilateral: NULL protein list"};
    cudaStream_t st = c.stream;
    DevBuf<alens_protein_bind> dP;
    DevBuf<alens_constraint_block> dOut;
    DevBuf<unsigned char> dKeep;
    dP.reserve((size_t)n); dOut.reserve((size_t)n); dKeep.reserve((size_t)n);
    ALENS_CUDA(cudaMemcpyAsync(dP.p, proteins, sizeof(alens_protein_bind) * (size_t)n, cudaMemcpyHostToDevice, st));
    k_protein_blocks<<<gridFor(n, 128), 128, 0, st>>>(n, dP.p, tubuleDiameter, dOut.p, dKeep.p);
    c.launches++;
    ALENS_CUDA(cudaGetLastError());
    std::vector<alens_constraint_block> host((size_t)n);
    std::vector<unsigned char> keep((size_t)n);
    ALENS_CUDA(cudaMemcpyAsync(host.data(), dOut.p, sizeof(alens_constraint_block) * (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(keep.data(), dKeep.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    size_t m = 0;
    for (long long i = 0; i < n; i++)
        if (keep[i]) host[m++] = host[i]; // protein order is kept
    if (m) appendBlocks(c, host.data(), (long long)m);
    return (long long)m;
}

// ------------------------------------------------------------------------------------------------
struct BlocksIn {
    const int *idxI, *idxJ, *gidI, *gidJ, *sUser, *uGlobalIdx;
    const signed char *shift;
    const double *delta0, *gamma0, *n, *pI, *pJ, *labI, *labJ;
    const double *sX, *sY, *sZ, *sDx, *sDy, *sDz, *sLc, *sRc;
    const double *gamma; // solved values or nullptr
    size_t stride;
    Box box;
    int globalIndexBase;
};

// collideStress of collision block k for unit gamma, by the shape of the two rods (SylinderNear.hpp:289-290, :350-351,
// :409-410); J at the periodic image the block was found with
__device__ __forceinline__ void blockUnitStress(const BlocksIn &in, long long k, int si, int sj, Vec3 labI, Vec3 labJ,
                                                double stress[9]) {
    const int code = in.shift[k];
    const int kx = code % 3 - 1, ky = (code / 3) % 3 - 1, kz = code / 9 - 1;
    const Vec3 cI = v3(in.sX[si], in.sY[si], in.sZ[si]);
    const Vec3 cJ = v3(in.sX[sj] + kx * in.box.len[0], in.sY[sj] + ky * in.box.len[1], in.sZ[sj] + kz * in.box.len[2]);
    const Vec3 dI = v3(in.sDx[si], in.sDy[si], in.sDz[si]), dJ = v3(in.sDx[sj], in.sDy[sj], in.sDz[sj]);
    const double lcI = in.sLc[si], rcI = in.sRc[si], lcJ = in.sLc[sj], rcJ = in.sRc[sj];
    const bool sa = lcI < 2 * rcI, sb = lcJ < 2 * rcJ;
    const Vec3 ez = v3(0, 0, 1);
    if (sa && sb) { // two spheres
        collideStress(ez, ez, cI, cJ, 0, 0, lcI * 0.5 + rcI, lcJ * 0.5 + rcJ, 1.0, labI, labJ, stress);
    } else if (sa) { // sphere I, sylinder J
        collideStress(ez, dJ, cI, cJ, 0, lcJ, lcI * 0.5 + rcI, rcJ, 1.0, labI, labJ, stress);
    } else if (sb) { // sylinder I, sphere J: same call with (sphere, sylinder) argument order
        collideStress(ez, dI, cJ, cI, 0, lcI, lcJ * 0.5 + rcJ, rcI, 1.0, labJ, labI, stress);
    } else {
        collideStress(dI, dJ, cI, cJ, lcI, lcJ, rcI, rcJ, 1.0, labI, labJ, stress);
    }
}

// one thread per collision block: assemble the 272-byte record (and the unit-gamma stress)
__global__ void k_blocks_out(long long n, BlocksIn in, int withStress, int writeBack, alens_constraint_block *out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t S = in.stride;
    alens_constraint_block b;
    memset(&b, 0, sizeof(b));
    const int si = in.idxI[k], sj = in.idxJ[k];
    b.delta0 = in.delta0[k];
    b.gamma = in.gamma0[k];
    b.gammaLB = 0;
    b.gidI = in.gidI[k];
    b.gidJ = in.gidJ[k];
    b.globalIndexI = in.uGlobalIdx[in.sUser[si]];
    b.globalIndexJ = in.uGlobalIdx[in.sUser[sj]];
    b.oneSide = 0;
    b.bilateral = 0;
    b.kappa = 0;
    for (int c = 0; c < 3; c++) {
        b.normI[c] = in.n[k + c * S];
        b.normJ[c] = -b.normI[c];
        b.posI[c] = in.pI[k + c * S];
        b.posJ[c] = in.pJ[k + c * S];
        b.labI[c] = in.labI[k + c * S];
        b.labJ[c] = in.labJ[k + c * S];
    }
    if (withStress)
        blockUnitStress(in, k, si, sj, v3(b.labI[0], b.labI[1], b.labI[2]), v3(b.labJ[0], b.labJ[1], b.labJ[2]), b.stress);
    if (writeBack && in.gamma) { // ConstraintCollector.cpp:449-458
        b.gamma = in.gamma[k];
        for (int c = 0; c < 9; c++) b.stress[c] *= b.gamma;
    }
    out[k] = b;
}

void downloadBlocks(Context &c, alens_constraint_block *out, long long cap, bool withStress, bool writeBack) {
    if (cap < c.nCon) throw ArgError{ALENS_ERR_ARG, "alens_get_constraints: output capacity too small"};
    if (writeBack && !c.haveSolution) throw ArgError{ALENS_ERR_STATE, "alens_get_constraints: writeBack needs a solve"};
    cudaStream_t st = c.stream;
    const long long nColl = c.nColl, nHost = c.nCon - c.nColl;
    if (nColl > 0) {
        DevBuf<alens_constraint_block> dOut;
        dOut.reserve((size_t)nColl);
        BlocksIn in{c.cIdxI.p, c.cIdxJ.p, c.cGidI.p, c.cGidJ.p, c.sUser.p, c.uGlobalIdx.p, c.cShift.p, c.cDelta0.p, c.cGamma0.p,
                    c.cN.p, c.cPI.p, c.cPJ.p, c.cLabI.p, c.cLabJ.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p,
                    c.sDz.p, c.sLc.p, c.sRc.p, writeBack ? c.xSolution : nullptr, c.conCap, c.box, 0};
        k_blocks_out<<<gridFor(nColl, 128), 128, 0, st>>>(nColl, in, withStress ? 1 : 0, writeBack ? 1 : 0, dOut.p);
        c.launches++;
        ALENS_CUDA(cudaGetLastError());
        ALENS_CUDA(cudaMemcpyAsync(out, dOut.p, sizeof(alens_constraint_block) * (size_t)nColl,
                                   cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
    }
    if (nHost > 0) {
        memcpy(out + nColl, c.hostBlocks.data(), sizeof(alens_constraint_block) * (size_t)nHost);
        if (writeBack) {
            std::vector<double> g((size_t)nHost);
            ALENS_CUDA(cudaMemcpyAsync(g.data(), c.xSolution + nColl, 8 * (size_t)nHost, cudaMemcpyDeviceToHost, st));
            ALENS_CUDA(cudaStreamSynchronize(st));
            for (long long i = 0; i < nHost; i++) {
                alens_constraint_block &b = out[nColl + i];
                b.gamma = g[i];
                for (int k = 0; k < 9; k++) b.stress[k] *= b.gamma;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sum of the block stresses after the write-back of gamma (ConstraintCollector::sumLocalConstraintStress,
// Constraint/ConstraintCollector.cpp:38-74, as SylinderSystem::calcConStress uses it every step,
// SylinderSystem.cpp:1226-1263) without bringing the blocks to the host: collision blocks are evaluated and reduced on the
// device (fixed grid, per-CTA partial sums added in CTA order: the result does not depend on the run), the appended
// blocks (boundary / link / protein: their unit stress is the host's) on the host with the solved gamma.
static constexpr int kStressCtas = 148 * 4;
__global__ void __launch_bounds__(128) k_stress_sum(long long n, BlocksIn in, const unsigned char *__restrict__ own,
                                                    double *__restrict__ part) {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const size_t S = in.stride;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        if (own && !own[k]) continue; // a row held by two ranks is counted by the owner of rod I
        const Vec3 labI = v3(in.labI[k], in.labI[k + S], in.labI[k + 2 * S]);
        const Vec3 labJ = v3(in.labJ[k], in.labJ[k + S], in.labJ[k + 2 * S]);
        double s[9];
        blockUnitStress(in, k, in.idxI[k], in.idxJ[k], labI, labJ, s);
        const double gm = in.gamma[k];
        for (int c = 0; c < 9; c++) acc[c] += s[c] * gm;
    }
    __shared__ double sh[4][9];
    for (int c = 0; c < 9; c++) {
        double v = acc[c];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) part[9 * (size_t)blockIdx.x + threadIdx.x] = ((sh[0][threadIdx.x] + sh[1][threadIdx.x]) + sh[2][threadIdx.x]) + sh[3][threadIdx.x];
}

void sumConstraintStress(Context &c, bool withOneSide, double uni[9], double bi[9]) {
    if (!c.haveSolution) throw ArgError{ALENS_ERR_STATE, "alens_sum_constraint_stress: needs a solve"};
    for (int k = 0; k < 9; k++) uni[k] = bi[k] = 0;
    cudaStream_t st = c.stream;
    const long long nColl = c.nColl, nHost = c.nCon - c.nColl;
    const bool multi = c.comm.active;
    std::vector<double> part, g((size_t)nHost);
    std::vector<unsigned char> own((size_t)nHost, 1);
    DevBuf<double> dPart;
    int grid = 0;
    if (nColl > 0) {
        grid = (int)std::min<long long>(kStressCtas, (nColl + 127) / 128);
        dPart.reserve(9 * (size_t)grid);
        BlocksIn in{c.cIdxI.p, c.cIdxJ.p, c.cGidI.p, c.cGidJ.p, c.sUser.p, c.uGlobalIdx.p, c.cShift.p, c.cDelta0.p, c.cGamma0.p,
                    c.cN.p, c.cPI.p, c.cPJ.p, c.cLabI.p, c.cLabJ.p, c.sX.p, c.sY.p, c.sZ.p, c.sDx.p, c.sDy.p,
                    c.sDz.p, c.sLc.p, c.sRc.p, c.xSolution, c.conCap, c.box, 0};
        k_stress_sum<<<grid, 128, 0, st>>>(nColl, in, multi ? c.cOwn.p : nullptr, dPart.p);
        c.launches++;
        ALENS_CUDA(cudaGetLastError());
        part.resize(9 * (size_t)grid);
        ALENS_CUDA(cudaMemcpyAsync(part.data(), dPart.p, 8 * part.size(), cudaMemcpyDeviceToHost, st));
    }
    if (nHost > 0) {
        ALENS_CUDA(cudaMemcpyAsync(g.data(), c.xSolution + nColl, 8 * (size_t)nHost, cudaMemcpyDeviceToHost, st));
        if (multi) ALENS_CUDA(cudaMemcpyAsync(own.data(), c.cOwn.p + nColl, (size_t)nHost, cudaMemcpyDeviceToHost, st));
    }
    ALENS_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < grid; b++) // collision blocks are unilateral and two-sided
        for (int k = 0; k < 9; k++) uni[k] += part[9 * (size_t)b + k];
    for (long long i = 0; i < nHost; i++) {
        const alens_constraint_block &b = c.hostBlocks[(size_t)i];
        if ((b.oneSide && !withOneSide) || !own[(size_t)i]) continue;
        double *dst = b.bilateral ? bi : uni;
        for (int k = 0; k < 9; k++) dst[k] += b.stress[k] * g[(size_t)i];
    }
}


// ------------------------------------------------------------------------------------------------
// Batch entry points of the narrow phase itself (the reference exposes both as public functors:
// DCPQuery<3,double,Evec3>::operator(), Collision/DCPQuery.hpp:199-308, and
// CalcSylinderNearForce::operator() / collideStress, Sylinder/SylinderNear.hpp:197-519, the latter also
// called from SRC/TubuleSystem.cpp:738).  One independent query per thread, no neighbour search.
__global__ void k_dcp_batch(long long n, const double *__restrict__ seg, double *__restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double *s = seg + 12 * k; // P0 P1 Q0 Q1
    Vec3 P, Q;
    const double d = segSegClosest(v3(s[0], s[1], s[2]), v3(s[3], s[4], s[5]), v3(s[6], s[7], s[8]), v3(s[9], s[10], s[11]), P, Q);
    double *o = out + 7 * k;
    o[0] = d;
    o[1] = P.x; o[2] = P.y; o[3] = P.z;
    o[4] = Q.x; o[5] = Q.y; o[6] = Q.z;
}

// geom: 9 doubles per rod {pos[3], direction[3], lengthCollision, radiusCollision, colBuf}
__global__ void k_pair_functor_batch(long long n, const double *__restrict__ gI, const double *__restrict__ gJ, int withStress,
                                     unsigned char *__restrict__ hit, alens_constraint_block *__restrict__ out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double *a = gI + 9 * k, *b = gJ + 9 * k;
    RodGeom A{v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), a[6], a[7]};
    RodGeom B{v3(b[0], b[1], b[2]), v3(b[3], b[4], b[5]), b[6], b[7]};
    const double buffer = a[8] > b[8] ? a[8] : b[8]; // std::max(colBufI, colBufJ), SylinderNear.hpp:268,322,391
    Contact ct;
    alens_constraint_block blk;
    memset(&blk, 0, sizeof(blk));
    const bool h = pairContact(A, B, buffer, ct);
    hit[k] = h ? 1 : 0;
    if (h) {
        blk.delta0 = ct.sep;
        blk.gamma = ct.sep < 0 ? -ct.sep : 0;
        blk.gidI = blk.globalIndexI = (int)(2 * k);
        blk.gidJ = blk.globalIndexJ = (int)(2 * k + 1);
        const double nn[3] = {ct.normI.x, ct.normI.y, ct.normI.z};
        const double pi[3] = {ct.posI.x, ct.posI.y, ct.posI.z}, pj[3] = {ct.posJ.x, ct.posJ.y, ct.posJ.z};
        const double li[3] = {ct.labI.x, ct.labI.y, ct.labI.z}, lj[3] = {ct.labJ.x, ct.labJ.y, ct.labJ.z};
        for (int c = 0; c < 3; c++) {
            blk.normI[c] = nn[c]; blk.normJ[c] = -nn[c];
            blk.posI[c] = pi[c]; blk.posJ[c] = pj[c];
            blk.labI[c] = li[c]; blk.labJ[c] = lj[c];
        }
        if (withStress) {
            const bool sa = A.lc < 2 * A.rc, sb = B.lc < 2 * B.rc;
            const Vec3 ez = v3(0, 0, 1);
            if (sa && sb) collideStress(ez, ez, A.c, B.c, 0, 0, A.lc * 0.5 + A.rc, B.lc * 0.5 + B.rc, 1.0, ct.labI, ct.labJ, blk.stress);
            else if (sa) collideStress(ez, B.d, A.c, B.c, 0, B.lc, A.lc * 0.5 + A.rc, B.rc, 1.0, ct.labI, ct.labJ, blk.stress);
            else if (sb) collideStress(ez, A.d, B.c, A.c, 0, A.lc, B.lc * 0.5 + B.rc, A.rc, 1.0, ct.labJ, ct.labI, blk.stress);
            else collideStress(A.d, B.d, A.c, B.c, A.lc, B.lc, A.rc, B.rc, 1.0, ct.labI, ct.labJ, blk.stress);
        }
    }
    out[k] = blk;
}

void dcpBatch(Context &c, long long n, const double *P0, const double *P1, const double *Q0, const double *Q1, double *dist,
              double *Ploc, double *Qloc) {
    if (n <= 0) return;
    cudaStream_t st = c.stream;
    std::vector<double> seg(12 * (size_t)n), res(7 * (size_t)n);
    for (long long k = 0; k < n; k++)
        for (int d = 0; d < 3; d++) {
            seg[12 * k + d] = P0[3 * k + d]; seg[12 * k + 3 + d] = P1[3 * k + d];
            seg[12 * k + 6 + d] = Q0[3 * k + d]; seg[12 * k + 9 + d] = Q1[3 * k + d];
        }
    DevBuf<double> dSeg, dOut;
    dSeg.reserve(seg.size());
    dOut.reserve(res.size());
    ALENS_CUDA(cudaMemcpyAsync(dSeg.p, seg.data(), 8 * seg.size(), cudaMemcpyHostToDevice, st));
    k_dcp_batch<<<gridFor(n, 128), 128, 0, st>>>(n, dSeg.p, dOut.p);
    c.launches++;
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaMemcpyAsync(res.data(), dOut.p, 8 * res.size(), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    for (long long k = 0; k < n; k++) {
        if (dist) dist[k] = res[7 * k];
        for (int d = 0; d < 3; d++) {
            if (Ploc) Ploc[3 * k + d] = res[7 * k + 1 + d];
            if (Qloc) Qloc[3 * k + d] = res[7 * k + 4 + d];
        }
    }
}

void pairFunctorBatch(Context &c, long long n, const double *geomI, const double *geomJ, int withStress, unsigned char *hit,
                      alens_constraint_block *blocks) {
    if (n <= 0) return;
    cudaStream_t st = c.stream;
    DevBuf<double> dI, dJ;
    DevBuf<unsigned char> dHit;
    DevBuf<alens_constraint_block> dBlk;
    dI.reserve(9 * (size_t)n); dJ.reserve(9 * (size_t)n); dHit.reserve((size_t)n); dBlk.reserve((size_t)n);
    ALENS_CUDA(cudaMemcpyAsync(dI.p, geomI, 72 * (size_t)n, cudaMemcpyHostToDevice, st));
    ALENS_CUDA(cudaMemcpyAsync(dJ.p, geomJ, 72 * (size_t)n, cudaMemcpyHostToDevice, st));
    k_pair_functor_batch<<<gridFor(n, 128), 128, 0, st>>>(n, dI.p, dJ.p, withStress, dHit.p, dBlk.p);
    c.launches++;
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaMemcpyAsync(hit, dHit.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(blocks, dBlk.p, sizeof(alens_constraint_block) * (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------------------------------------
// Order-independent digest of the constraint list and of the solved multipliers, computed on the device (bench.py's
// parity flag: the list of a multi-GPU run against the single-GPU run of the same suspension, fused against unfused
// protocol, without moving 272-byte blocks to the host).  A row mirrored on two ranks is counted by the owner of rod I.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct DigestIn {
    long long n;
    const int *gidI, *gidJ;
    const double *delta0, *labJ; // labJ: [3][stride]
    size_t stride;
    const unsigned char *own;    // nullptr: every row counts
    const double *gamma;         // nullptr: no solve yet
};
__global__ void __launch_bounds__(256) k_constraint_digest(DigestIn in, unsigned long long *__restrict__ u64,
                                                           double *__restrict__ part) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cnt = 0, hl = 0, hg = 0;
    double s0 = 0, s1 = 0, s2 = 0;
    if (k < in.n && (!in.own || in.own[k])) {
        const unsigned long long key = ((unsigned long long)(unsigned)in.gidI[k] << 32) | (unsigned)in.gidJ[k];
        unsigned long long h = mix64(key + 0x9E3779B97F4A7C15ull);
        h = mix64(h + (unsigned long long)__double_as_longlong(in.labJ[k]));
        h = mix64(h + (unsigned long long)__double_as_longlong(in.labJ[k + in.stride]));
        h = mix64(h + (unsigned long long)__double_as_longlong(in.labJ[k + 2 * in.stride]));
        cnt = 1;
        hl = mix64(h + (unsigned long long)__double_as_longlong(in.delta0[k]));
        if (in.gamma) {
            const double g = in.gamma[k];
            hg = mix64(h + (unsigned long long)__double_as_longlong(g));
            const double w = (double)(h >> 44) * (1.0 / 1048576.0); // weight in [0, 1) tied to the row's identity
            s0 = g; s1 = g * g; s2 = w * g;
        }
    }
    __shared__ double sh[3][8];
    __shared__ unsigned long long su[3][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        hl += __shfl_xor_sync(0xffffffffu, hl, o);
        hg += __shfl_xor_sync(0xffffffffu, hg, o);
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { su[0][w] = cnt; su[1][w] = hl; su[2][w] = hg; sh[0][w] = s0; sh[1][w] = s1; sh[2][w] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) {
            su[0][0] += su[0][i]; su[1][0] += su[1][i]; su[2][0] += su[2][i];
            sh[0][0] += sh[0][i]; sh[1][0] += sh[1][i]; sh[2][0] += sh[2][i];
        }
        atomicAdd(u64 + 0, su[0][0]); atomicAdd(u64 + 1, su[1][0]); atomicAdd(u64 + 2, su[2][0]);
        part[3 * (size_t)blockIdx.x] = sh[0][0]; part[3 * (size_t)blockIdx.x + 1] = sh[1][0];
        part[3 * (size_t)blockIdx.x + 2] = sh[2][0];
    }
}

void constraintDigest(Context &c, unsigned long long u64[3], double f64[3]) {
    u64[0] = u64[1] = u64[2] = 0;
    f64[0] = f64[1] = f64[2] = 0;
    const long long n = c.nCon;
    if (n <= 0) return;
    cudaStream_t st = c.stream;
    const int grid = gridFor(n, 256);
    DevBuf<unsigned long long> dU;
    DevBuf<double> dP;
    dU.reserve(4);
    dP.reserve(3 * (size_t)grid);
    ALENS_CUDA(cudaMemsetAsync(dU.p, 0, 3 * sizeof(unsigned long long), st));
    DigestIn in{n, c.cGidI.p, c.cGidJ.p, c.cDelta0.p, c.cLabJ.p, c.conCap, c.comm.active ? c.cOwn.p : nullptr,
                c.haveSolution ? c.xSolution : nullptr};
    k_constraint_digest<<<grid, 256, 0, st>>>(in, dU.p, dP.p);
    c.launches++;
    ALENS_CUDA(cudaGetLastError());
    std::vector<double> part(3 * (size_t)grid);
    ALENS_CUDA(cudaMemcpyAsync(u64, dU.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(part.data(), dP.p, 8 * part.size(), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < grid; b++) // fixed order
        for (int k = 0; k < 3; k++) f64[k] += part[3 * (size_t)b + k];
}

// Force the (lazily loaded) kernels of this file into the context now: loading a kernel at its first launch can
// synchronise the context, which deadlocks against a peer rank's waiting kernel when two ranks share one GPU.
void preloadBlockKernels() {
    cudaFuncAttributes a;
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_append));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_blocks_out));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_stress_sum));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_dcp_batch));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_protein_blocks));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_constraint_digest));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_pair_functor_batch));
}

} // namespace alens
