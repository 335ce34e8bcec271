// solver.cu -- mobility, matrix-free constraint operator and the BCQP (BBPGD / APGD) loops.
//
// Replaces ConstraintCollector::buildConstraintMatrixVector (SimToolbox/Constraint/ConstraintCollector.cpp:237-423),
// SylinderSystem::calcMobMatrix (SimToolbox/Sylinder/SylinderSystem.cpp:622-717), ConstraintOperator
// (Constraint/ConstraintOperator.cpp:4-71), ConstraintSolver::setup/solveConstraints
// (Constraint/ConstraintSolver.cpp:4-107) and BCQPSolver::solveBBPGD/solveAPGD (Constraint/BCQPSolver.cpp:134-497).
//
// D is never materialised as CSR.  A constraint k stores (n, posI, posJ, idxI, idxJ); row k of D^T is
// [n, posI x n] on rod I and [-n, posJ x (-n)] on rod J (ConstraintCollector.cpp:298-341 with normJ = -normI).
//   k_force_vel_rec : f = D x and u = M f over the rod -> constraint incidence, touching only the slots whose multiplier can
//                 be non-zero: thread per rod, rod header {first slot, live bits} -> one 64-byte record per live slot
//                 {row id, M * column} + the row's {x, g} pair; the sums are the rod's velocity (force_kernel = 3,
//                 rec_mode = 2: the default; rec_mode 0 / 1 keep the plain column and apply M from (q, 1/drag)).  One
//                 launch per operator apply.  Inside the BBPGD loop x is not read but recomputed on the fly as
//                 P(x_prev - alpha g_prev) from the interleaved {x, g} pairs.  Predecessors kept as cross-checks:
//                 k_force_vel_act (rod-major slots + row mask, force_kernel = 1), k_slot_x + k_rod_sum (2), the dense
//                 level-major k_force_vel_lm (0).
//   k_bb_tail   : x = P(x_prev - alpha g_prev) again (same arithmetic, same bits), y = D^T u + K^-1 x, g = y + b,
//                 projected-gradient residual, BB dot products, the may-be-non-zero bit of every row (and the flips of
//                 the slot bits / rod headers k_force_vel_rec reads), deterministic
//                 two-level reduction, step-size/termination logic (and, multi-GPU, the mailbox allreduce) in the last
//                 CTA.  One BBPGD iteration = these two launches, chained with programmatic dependent launch; the host
//                 follows the loop through two progress words in pinned memory.
// Compiled with -fmad=false so that elementwise arithmetic rounds like the CPU restatement.
#include "context.hpp"
#include "comm_dev.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace alens {

static constexpr int kProfEvery = 8; // profiling on: every 8th BBPGD iteration (1, 9, 17, ...) carries event pairs
static inline void cpuRelax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}
// rod-major incidence: doubles per column record.  6 are used.  8 (= 64 bytes, aligned: one DRAM burst per record instead
// of 1.5 on average) was measured: force kernel 91.3 -> 89.2 us, but the per-step build of the records 0.41 -> 0.51 ms.
static constexpr int kColRec = 6;
static constexpr int kVecBlock = 256; // threads per CTA of the per-constraint kernels
static constexpr double kHuge = DBL_MAX / 10; // BCQPSolver.cpp:499-510

// ------------------------------------------------------------------------------------------------
// mobility coefficients: Sylinder::calcDragCoeff (Sylinder.cpp:69-82); immovable rods get zero
// mobility (SylinderSystem.cpp:660-662)
__global__ void k_mob_coeff(int n, const double *__restrict__ len, const double *__restrict__ rad,
                            const unsigned char *__restrict__ imm, double mu, double *__restrict__ invDrag,
                            size_t stride, const double *__restrict__ dx, const double *__restrict__ dy,
                            const double *__restrict__ dz, double *__restrict__ mobRec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double Pi = 3.14159265358979323846;
    const double length = len[i], radius = rad[i];
    double dPara, dPerp, dRot;
    if (length < radius * 2) {
        const double r = 0.5 * length + radius;
        dPara = 6 * Pi * r * mu;
        dPerp = dPara;
        dRot = 8 * Pi * r * r * r * mu;
    } else {
        const double b = -(1 + 2 * log(radius / (length)));
        dPara = 8 * Pi * length * mu / (2 * b);
        dPerp = 8 * Pi * length * mu / (b + 2);
        dRot = 2 * Pi * mu * length * length * length / (3 * (b + 2));
    }
    const bool im = imm[i] != 0;
    invDrag[i] = im ? 0.0 : 1 / dPara;
    invDrag[stride + i] = im ? 0.0 : 1 / dPerp;
    invDrag[2 * stride + i] = im ? 0.0 : 1 / dRot;
    if (mobRec) { // the same six numbers as ONE 64-byte line per rod: {q, 1/zeta_para, 1/zeta_perp, 1/zeta_rot} (k_force_vel_rec)
        double *o = mobRec + 8 * (size_t)i;
        o[0] = dx[i]; o[1] = dy[i]; o[2] = dz[i];
        o[3] = invDrag[i]; o[4] = invDrag[stride + i]; o[5] = invDrag[2 * stride + i];
        o[6] = 0.0; o[7] = 0.0;
    }
}

struct MobIn {
    const double *dx, *dy, *dz; // unit direction q
    const double *invDrag;      // [3][stride]
    const double *rec;          // [n][8]: {q, 1/zeta_para, 1/zeta_perp, 1/zeta_rot, 0, 0}, one 64-byte line per rod
    int n;
    size_t stride;              // even (16-byte aligned component arrays)
    const unsigned char *ghost; // 1 = rod owned by a neighbour rank: its U row arrives through the halo
};

// u = M f with Mtt = qq^T/zPara + (I - qq^T)/zPerp, Mrr = I/zRot (SylinderSystem.cpp:664-665)
__device__ __forceinline__ void applyMob(const MobIn &m, int r, const double f[6], double u[6]) {
    const double qx = m.dx[r], qy = m.dy[r], qz = m.dz[r];
    const double iPara = m.invDrag[r], iPerp = m.invDrag[m.stride + r], iRot = m.invDrag[2 * m.stride + r];
    const double qf = qx * f[0] + qy * f[1] + qz * f[2];
    const double px = qf * qx, py = qf * qy, pz = qf * qz;
    u[0] = iPara * px + iPerp * (f[0] - px);
    u[1] = iPara * py + iPerp * (f[1] - py);
    u[2] = iPara * pz + iPerp * (f[2] - pz);
    u[3] = iRot * f[3];
    u[4] = iRot * f[4];
    u[5] = iRot * f[5];
}

// y = M x on 6n vectors given in USER order (alens_mobility_apply)
__global__ void k_mob_apply_user(MobIn m, const int *__restrict__ userToSorted, const double *__restrict__ x,
                                 double *__restrict__ y) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= m.n) return;
    double f[6], o[6];
    for (int c = 0; c < 6; c++) f[c] = x[6 * u + c];
    applyMob(m, userToSorted[u], f, o);
    for (int c = 0; c < 6; c++) y[6 * u + c] = o[c];
}

// ------------------------------------------------------------------------------------------------
// calcVelocityBrown (SylinderSystem.cpp:1020-1091).  Counter-based normals: Philox4x32-10 (Salmon et al. 2011) with
// key = (seed, step) and counter = (gid, draw block), two Box-Muller pairs per block.
__device__ __forceinline__ void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                           unsigned out[4]) {
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1,
                       n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void normals4(unsigned long long seed, unsigned long long step, int gid, int block, double w[4]) {
    unsigned r[4];
    philox4x32((unsigned)gid, (unsigned)block, (unsigned)step, (unsigned)(step >> 32), (unsigned)seed, (unsigned)(seed >> 32), r);
    for (int h = 0; h < 2; h++) { // Box-Muller on (0, 1] x [0, 1)
        const double u1 = ((double)r[2 * h] + 1.0) * (1.0 / 4294967296.0), u2 = (double)r[2 * h + 1] * (1.0 / 4294967296.0);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        w[2 * h] = rad * cs;
        w[2 * h + 1] = rad * sn;
    }
}
// q * (0,0,1) (Eigen quaternion rotation: v + w uv + q.vec x uv, uv = 2 q.vec x v)
__device__ __forceinline__ void quatZ(const double q[4], double d[3]) {
    const double ux = q[1] + q[1], uy = -(q[0] + q[0]);
    d[0] = q[3] * ux + (-(q[2] * uy));
    d[1] = q[3] * uy + q[2] * ux;
    d[2] = 1.0 + (q[0] * uy - q[1] * ux);
}
__global__ void k_velocity_brown(int nLocal, const int *__restrict__ userToSorted, const int *__restrict__ uGid,
                                 const double *__restrict__ uQuat, const double *__restrict__ invDrag, size_t stride,
                                 double kBT, double dt, const double *__restrict__ normals, unsigned long long seed,
                                 unsigned long long step, double *__restrict__ out) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nLocal) return;
    const int s = userToSorted[u];
    const double a = invDrag[s], b = invDrag[stride + s], cR = invDrag[2 * stride + s]; // 1/zPara, 1/zPerp, 1/zRot
    double W[12];
    if (normals) {
        for (int k = 0; k < 12; k++) W[k] = normals[12 * (size_t)u + k];
    } else {
        for (int blk = 0; blk < 3; blk++) normals4(seed, step, uGid[u], blk, W + 4 * blk);
    }
    const double *Wrot = W, *Wpos = W + 3, *Wrfdrot = W + 6, *Wrfdpos = W + 9;
    const double delta = dt * 0.1, kBTfactor = sqrt(2 * kBT / dt);
    double q4[4] = {uQuat[4 * (size_t)u], uQuat[4 * (size_t)u + 1], uQuat[4 * (size_t)u + 2], uQuat[4 * (size_t)u + 3]};
    double d[3];
    quatZ(q4, d);
    double N[3][3], L[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) N[i][j] = (a - b) * (d[i] * d[j]) + b * (i == j ? 1.0 : 0.0);
    // lower Cholesky factor as Eigen's unblocked LLT: stops at a non-positive pivot, leaving the input's entries (a zero
    // matrix -- immovable rod -- stays zero)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) L[i][j] = j <= i ? N[i][j] : 0.0;
    for (int k = 0; k < 3; k++) {
        double x = L[k][k];
        for (int j = 0; j < k; j++) x -= L[k][j] * L[k][j];
        if (x <= 0) break;
        x = sqrt(x);
        L[k][k] = x;
        for (int i = k + 1; i < 3; i++) {
            double v = L[i][k];
            for (int j = 0; j < k; j++) v -= L[i][j] * L[k][j];
            L[i][k] = v / x;
        }
    }
    // orientation rotated by Wrfdrot * delta (EquatnHelper::rotateEquatn, Util/EquatnHelper.hpp:74-90)
    {
        const double ox = Wrfdrot[0], oy = Wrfdrot[1], oz = Wrfdrot[2];
        const double w = sqrt(ox * ox + oy * oy + oz * oz);
        if (!(w < (double)FLT_EPSILON)) {
            const double winv = 1 / w, sw = sin(w * delta / 2), cw = cos(w * delta / 2);
            const double sc = q4[3], px = q4[0], py = q4[1], pz = q4[2];
            const double cx = oy * pz - oz * py, cy = oz * px - ox * pz, cz = ox * py - oy * px;
            const double nx = sc * sw * ox * winv + cw * px + sw * winv * cx;
            const double ny = sc * sw * oy * winv + cw * py + sw * winv * cy;
            const double nz = sc * sw * oz * winv + cw * pz + sw * winv * cz;
            const double nw = sc * cw - (px * ox + py * oy + pz * oz) * sw * winv;
            const double nn = sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
            q4[0] = nx / nn; q4[1] = ny / nn; q4[2] = nz / nn; q4[3] = nw / nn;
        }
    }
    double dr[3];
    quatZ(q4, dr);
    double vel[3];
    for (int i = 0; i < 3; i++) {
        double g = 0, r = 0;
        for (int j = 0; j <= i; j++) g += L[i][j] * Wpos[j];
        for (int j = 0; j < 3; j++) {
            const double nr = (a - b) * (dr[i] * dr[j]) + b * (i == j ? 1.0 : 0.0);
            r += (nr - N[i][j]) * Wrfdpos[j];
        }
        vel[i] = kBTfactor * g + (kBT / delta) * r;
    }
    const double so = sqrt(cR) * kBTfactor;
    double *o = out + 6 * (size_t)u;
    o[0] = vel[0]; o[1] = vel[1]; o[2] = vel[2];
    o[3] = so * Wrot[0]; o[4] = so * Wrot[1]; o[5] = so * Wrot[2];
}

// calcVelocityNonCon (SylinderSystem.cpp:724-800): vNC = M f + vNB + vB per local rod in the caller's order; the monolayer
// mask zeroes v_z, omega_x, omega_y of every term (:737-743, :762-769, :789-797); updates are 1.0 * A + 1.0 * Y
__global__ void k_velocity_noncon(MobIn m, int nLocal, const int *__restrict__ userToSorted, const double *__restrict__ force,
                                  const double *__restrict__ velNB, const double *__restrict__ velB, int monolayer,
                                  double *__restrict__ velNC, double *__restrict__ velNonBOut) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nLocal) return;
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (force) {
        double f[6];
        for (int c = 0; c < 6; c++) f[c] = force[6 * (size_t)u + c];
        applyMob(m, userToSorted[u], f, v);
        if (monolayer) v[2] = v[3] = v[4] = 0;
    }
    if (velNB)
        for (int c = 0; c < 6; c++) {
            const double a = (monolayer && (c == 2 || c == 3 || c == 4)) ? 0.0 : velNB[6 * (size_t)u + c];
            v[c] = 1.0 * a + 1.0 * v[c];
        }
    if (velNonBOut)
        for (int c = 0; c < 6; c++) velNonBOut[6 * (size_t)u + c] = v[c];
    if (velB)
        for (int c = 0; c < 6; c++) {
            const double a = (monolayer && (c == 2 || c == 3 || c == 4)) ? 0.0 : velB[6 * (size_t)u + c];
            v[c] = 1.0 * a + 1.0 * v[c];
        }
    for (int c = 0; c < 6; c++) velNC[6 * (size_t)u + c] = v[c];
}

// ------------------------------------------------------------------------------------------------
// incidence rod -> constraints (replaces the explicit transpose of ConstraintOperator.cpp:14-20)
__global__ void k_inc_count(long long nc, const int *__restrict__ idxI, const int *__restrict__ idxJ,
                            const unsigned char *__restrict__ ghost, int *__restrict__ deg) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    const int i = idxI[k], j = idxJ[k];
    if (!ghost[i]) atomicAdd(&deg[i], 1); // ghost rods get no slots: their force/velocity is the owner's business
    if (j >= 0 && !ghost[j]) atomicAdd(&deg[j], 1);
}

// slot code = 4*constraint + 2*bilateral + side (side 1 = the J rod): the force kernel needs the bilateral flag
// of a gathered constraint (lower bound of the projection, gamma_b mask) and gets it with the id
__global__ void k_inc_fill(long long nc, const int *__restrict__ idxI, const int *__restrict__ idxJ,
                           const unsigned char *__restrict__ bi, const unsigned char *__restrict__ ghost,
                           const int *__restrict__ start, int *__restrict__ fill, int *__restrict__ incCon) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    const int i = idxI[k], j = idxJ[k];
    const int code = (int)(4 * k) + (bi[k] ? 2 : 0);
    if (!ghost[i]) incCon[start[i] + atomicAdd(&fill[i], 1)] = code;
    if (j >= 0 && !ghost[j]) incCon[start[j] + atomicAdd(&fill[j], 1)] = code + 1;
}

struct ConGeom {
    const int *idxI, *idxJ;
    const double *n, *pI, *pJ; // [3][stride]
    size_t stride;
};

// One warp per group of 32 consecutive rods.  Each lane sorts its rod's raw slot list by (constraint, side)
// -> fixed summation order; the group's slots are then emitted LEVEL-MAJOR (the 1st slot of every rod of
// the group, then the 2nd, ...).  In the force kernel lane l of a warp reads the k-th slot of rod l: with
// this layout consecutive lanes touch consecutive addresses (no shared-memory bank conflicts, coalesced
// writes here).  Each slot gets its 6-vector D column block s*[n, p x n] (ConstraintCollector.cpp:313-318).
__global__ void k_inc_emit(int nRods, const int *__restrict__ start, int *__restrict__ raw, ConGeom g,
                           int *__restrict__ incCon, double *__restrict__ incCol, size_t stride) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp * 32 >= nRods) return;
    const int r = warp * 32 + lane;
    int b = 0, e = 0;
    if (r < nRods) {
        b = start[r];
        e = start[r + 1];
    }
    for (int a = b + 1; a < e; a++) { // insertion sort, lists are short
        const int v = raw[a];
        int p = a - 1;
        while (p >= b && raw[p] > v) {
            raw[p + 1] = raw[p];
            p--;
        }
        raw[p + 1] = v;
    }
    const int d = e - b;
    int off = __shfl_sync(0xffffffffu, b, 0); // first slot of the group
    const unsigned lt = (1u << lane) - 1;
    for (int k = 0;; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, k < d);
        if (!m) break;
        if (k < d) {
            const size_t s = (size_t)(off + __popc(m & lt));
            const int k2 = raw[b + k];
            const size_t kk = (size_t)(k2 >> 2);
            const bool sideJ = k2 & 1;
            double gx = g.n[kk], gy = g.n[kk + g.stride], gz = g.n[kk + 2 * g.stride];
            const double *P = sideJ ? g.pJ : g.pI;
            const double px = P[kk], py = P[kk + g.stride], pz = P[kk + 2 * g.stride];
            if (sideJ) { gx = -gx; gy = -gy; gz = -gz; }
            incCon[s] = k2;
            incCol[s] = gx;
            incCol[stride + s] = gy;
            incCol[2 * stride + s] = gz;
            incCol[3 * stride + s] = (gz * py - gy * pz);
            incCol[4 * stride + s] = (gx * pz - gz * px);
            incCol[5 * stride + s] = (gy * px - gx * py);
        }
        off += __popc(m);
    }
}

// Rod-major variant for k_force_vel_act: each lane sorts its rod's slot list in place (incCon IS the sorted raw
// list), then the warp walks the group's contiguous slot range slot-parallel and writes each slot's column block as
// one 48-byte record (array of structures): a reader that needs only SOME slots touches 2 sectors per slot.
__global__ void k_inc_emit_rm(int nRods, const int *__restrict__ start, int *__restrict__ incCon, ConGeom g,
                              double *__restrict__ incCol6) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp * 32 >= nRods) return;
    const int r = warp * 32 + lane;
    const int b = start[min(r, nRods)], e = start[min(r + 1, nRods)];
    for (int a = b + 1; a < e; a++) { // insertion sort, lists are short
        const int v = incCon[a];
        int p = a - 1;
        while (p >= b && incCon[p] > v) {
            incCon[p + 1] = incCon[p];
            p--;
        }
        incCon[p + 1] = v;
    }
    __syncwarp();
    const int gb = __shfl_sync(0xffffffffu, b, 0), ge = __shfl_sync(0xffffffffu, e, 31);
    for (int p = gb + lane; p < ge; p += 32) {
        const int k2 = incCon[p];
        const size_t kk = (size_t)(k2 >> 2);
        const bool sideJ = k2 & 1;
        double gx = g.n[kk], gy = g.n[kk + g.stride], gz = g.n[kk + 2 * g.stride];
        const double *P = sideJ ? g.pJ : g.pI;
        const double px = P[kk], py = P[kk + g.stride], pz = P[kk + 2 * g.stride];
        if (sideJ) { gx = -gx; gy = -gy; gz = -gz; }
        double2 *o = reinterpret_cast<double2 *>(incCol6 + kColRec * (size_t)p);
        o[0] = make_double2(gx, gy);
        o[1] = make_double2(gz, (gz * py - gy * pz));
        o[2] = make_double2((gx * pz - gz * px), (gy * px - gx * py));
    }
}

// Record layout for k_force_vel_rec (force_kernel = 3): every incidence slot owns one 64-byte aligned record
// {x, g, D column block[6]}.  The column block is written here, once per step; {x, g} of a slot whose constraint row can
// be non-zero is refreshed by k_bb_tail every iteration, which also keeps the slot-ordered bitmap slotLive up to date.
// The force kernel then needs no constraint ids at all: bitmap word -> one aligned 64-byte record per live slot.
// cSlot[k] = (slot of row k in rod I's list, slot in rod J's list), -1 where the side has no slot (one-sided, ghost rod).
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d); // 256-bit store, below
// mobRec != nullptr (rec_mode 2): the record holds M_rod * column instead of the column, so that the force kernel adds up
// the rod's VELOCITY directly (u = sum_s (M c_s) x_s) and never reads the rod's mobility data.
__global__ void k_inc_emit_rec(int nRods, const int *__restrict__ start, int *__restrict__ incCon, ConGeom g,
                               double *__restrict__ rec, int2 *__restrict__ cSlot, unsigned *__restrict__ slotBi,
                               const double *__restrict__ mobRec) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp * 32 >= nRods) return;
    const int r = warp * 32 + lane;
    const int b = start[min(r, nRods)], e = start[min(r + 1, nRods)];
    for (int a = b + 1; a < e; a++) { // insertion sort by (constraint, side): fixed summation order, lists are short
        const int v = incCon[a];
        int p = a - 1;
        while (p >= b && incCon[p] > v) {
            incCon[p + 1] = incCon[p];
            p--;
        }
        incCon[p + 1] = v;
    }
    __syncwarp();
    const int gb = __shfl_sync(0xffffffffu, b, 0), ge = __shfl_sync(0xffffffffu, e, 31);
    for (int base = gb; base < ge; base += 32) { // (warp-uniform trip count: the owner search below shuffles)
        const int p = base + lane;
        int owner = 0; // lane whose rod owns slot p: the last lane with start <= p
        if (mobRec) {
            int lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi + 1) >> 1;
                const int bm = __shfl_sync(0xffffffffu, b, mid);
                if (bm <= p) lo = mid;
                else hi = mid - 1;
            }
            owner = warp * 32 + lo;
        }
        if (p >= ge) continue;
        const int k2 = incCon[p];
        const size_t kk = (size_t)(k2 >> 2);
        const bool sideJ = k2 & 1;
        double gx = g.n[kk], gy = g.n[kk + g.stride], gz = g.n[kk + 2 * g.stride];
        const double *P = sideJ ? g.pJ : g.pI;
        const double px = P[kk], py = P[kk + g.stride], pz = P[kk + 2 * g.stride];
        if (sideJ) { gx = -gx; gy = -gy; gz = -gz; }
        double c0 = gx, c1 = gy, c2 = gz, c3 = (gz * py - gy * pz), c4 = (gx * pz - gz * px), c5 = (gy * px - gx * py);
        if (mobRec) { // M c with Mtt = qq^T/zPara + (I - qq^T)/zPerp, Mrr = I/zRot
            const double *m = mobRec + 8 * (size_t)owner;
            const double qx = m[0], qy = m[1], qz = m[2], iPara = m[3], iPerp = m[4], iRot = m[5];
            const double qf = qx * c0 + qy * c1 + qz * c2;
            const double ax = qf * qx, ay = qf * qy, az = qf * qz;
            c0 = iPara * ax + iPerp * (c0 - ax);
            c1 = iPara * ay + iPerp * (c1 - ay);
            c2 = iPara * az + iPerp * (c2 - az);
            c3 = iRot * c3; c4 = iRot * c4; c5 = iRot * c5;
        }
        // two 256-bit stores = two full 32-byte sectors (rec_mode 1/2: the slot code in front; rec_mode 0: {x, g} go there later)
        st256(rec + 8 * (size_t)p, __longlong_as_double((long long)k2), 0.0, c0, c1);
        st256(rec + 8 * (size_t)p + 4, c2, c3, c4, c5);
        if (sideJ) cSlot[kk].y = p;
        else cSlot[kk].x = p;
        if (k2 & 2) atomicOr(slotBi + (p >> 5), 1u << (p & 31));
    }
}

// Stable partition of the tile indices by a 0/1 flag, unflagged tiles first (flagFirst = 0) or flagged tiles first
// (flagFirst = 1); order[nTiles] = size of the first group.  One CTA (at most a few 10^4 tiles).
__global__ void k_tile_order(int nTiles, const int *__restrict__ flag, int flagFirst, int *__restrict__ order) {
    __shared__ int sCnt[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int chunk = (nTiles + T - 1) / T, b = min(nTiles, t * chunk), e = min(nTiles, b + chunk);
    int first = 0;
    for (int i = b; i < e; i++) first += ((flag[i] != 0) == (flagFirst != 0)) ? 1 : 0;
    sCnt[t] = first;
    __syncthreads();
    for (int off = 1; off < T; off <<= 1) {
        const int v = t >= off ? sCnt[t - off] : 0;
        __syncthreads();
        sCnt[t] += v;
        __syncthreads();
    }
    const int nFirst = sCnt[T - 1];
    int pf = t == 0 ? 0 : sCnt[t - 1]; // tiles of the first group in front of my chunk
    int ps = nFirst + (b - pf);        // ... of the second group
    for (int i = b; i < e; i++) {
        if ((flag[i] != 0) == (flagFirst != 0)) order[pf++] = i;
        else order[ps++] = i;
    }
    if (t == 0) order[nTiles] = nFirst;
}
// tiles of 128 rods (k_force_vel_rec's CTA) that contain a rod mirrored on a neighbour rank
__global__ void k_rod_tile_flag(int nRods, const int *__restrict__ mirL, const int *__restrict__ mirR, int *__restrict__ flag) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRods) return;
    if ((mirL && mirL[r] >= 0) || (mirR && mirR[r] >= 0)) flag[r >> 7] = 1;
}

// ------------------------------------------------------------------------------------------------
// setup: q = delta0/dt + D^T v_nc (ConstraintSolver.cpp:18-27), K^-1/dt, bilateral flag, x0 = gamma guess
__global__ void k_setup(long long nc, ConGeom g, const int *__restrict__ sUser, const double *__restrict__ velNC,
                        const double *__restrict__ delta0, const double *__restrict__ gamma0,
                        const double *__restrict__ invKappa, const unsigned char *__restrict__ bi, double invDt,
                        double *__restrict__ b, double *__restrict__ invKdt, double *__restrict__ lbFlag,
                        double *__restrict__ x0, const unsigned char *__restrict__ ghost, int *__restrict__ tileFlag) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    if (tileFlag) { // multi-GPU: tiles of 256 rows (k_bb_tail's unit of work) that contain a row reading a ghost rod's velocity
        const int j = g.idxJ[k];
        if (ghost[g.idxI[k]] || (j >= 0 && ghost[j])) tileFlag[k / kVecBlock] = 1;
    }
    double dnc = 0;
    if (velNC) {
        const double gx = g.n[k], gy = g.n[k + g.stride], gz = g.n[k + 2 * g.stride];
        {
            const double *v = velNC + 6 * (size_t)sUser[g.idxI[k]];
            const double px = g.pI[k], py = g.pI[k + g.stride], pz = g.pI[k + 2 * g.stride];
            dnc = gx * v[0];
            dnc += gy * v[1];
            dnc += gz * v[2];
            dnc += (gz * py - gy * pz) * v[3];
            dnc += (gx * pz - gz * px) * v[4];
            dnc += (gy * px - gx * py) * v[5];
        }
        const int j = g.idxJ[k];
        if (j >= 0) {
            const double *v = velNC + 6 * (size_t)sUser[j];
            const double px = g.pJ[k], py = g.pJ[k + g.stride], pz = g.pJ[k + 2 * g.stride];
            const double hx = -gx, hy = -gy, hz = -gz;
            dnc += hx * v[0];
            dnc += hy * v[1];
            dnc += hz * v[2];
            dnc += (hz * py - hy * pz) * v[3];
            dnc += (hx * pz - hz * px) * v[4];
            dnc += (hy * px - hx * py) * v[5];
        }
    }
    b[k] = 1.0 * (delta0[k] * invDt) + 1.0 * dnc;
    invKdt[k] = invKappa[k] * invDt;
    lbFlag[k] = bi[k] ? 1.0 : 0.0;
    x0[k] = gamma0[k];
}

// ------------------------------------------------------------------------------------------------
// streaming (read-once) loads bypass L1 allocation and are first in line for L2 eviction, so that the
// gather targets (U, x) stay resident
// gather targets (U, x) stay resident.  `asm volatile` pins the issue order: the SM issues in order, so a
// dependent gather placed between independent streaming loads would stall everything behind it.
__device__ __forceinline__ double ldStream(const double *p) {
    double v;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldStream(const int *p) {
    int v;
    asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldStream2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// mbarrier + TMA 1-D bulk copy (global -> shared, completion counted in bytes on the barrier)
__device__ __forceinline__ unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    do { // try_wait suspends the thread in hardware until the phase flips or a time limit passes
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smemAddr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulkLoad(void *dstSmem, const void *srcGlobal, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

// Programmatic dependent launch (PDL): inside the BBPGD loop the force and tail kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs become resident while its predecessor drains.
// pdlWait() returns once the predecessor grid has completed and its writes are visible; everything before it may only
// touch data that is at least TWO kernels old.  That holds because every kernel signals pdlLaunchDependents() only
// AFTER its own pdlWait(): when a dependent starts, the predecessor of its predecessor is complete.
// Both are no-ops in a kernel that was launched without the attribute.
// instrumentation (alens_set_option("stamps", 1)): nanosecond stamps of one BBPGD iteration, 8 words per iteration:
// [0] first force CTA past its dependency wait (min), [1] force kernel released the halo flags, [2] last force CTA done (max),
// [3] first tail CTA past its dependency wait (min), [4] first tail CTA reached the halo wait (min), [5] last tail CTA
// saw the halo (max), [6] last tail CTA finished its rows (ticket), [7] allreduce complete / scalar step taken
__device__ __forceinline__ unsigned long long globalNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdlLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// L2 residency control for the one array both BBPGD kernels share through L2: the tail kernel stores {x, g}
// (16 B/row, 55 MB at 3.4M rows) and the force kernel gathers the rows that can be non-zero right afterwards.
// Stored and gathered with an evict_last policy the pairs survive the 400 MB the tail streams past them; the tail's
// own read of the PREVIOUS pairs is a streaming load (evict-first), which demotes the old buffer again.
__device__ __forceinline__ unsigned long long policyEvictLast() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ldGather2Keep(const double2 *p, unsigned long long pol) {
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stKeep2(double2 *p, double2 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
// 256-bit global store / load (PTX ISA 8.8, sm_100+): one 32-byte sector per instruction; p must be 32-byte aligned
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ double2 ldGather2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// f = D x, u = M f
struct FvIn {
    const int *incStart, *incCon; // slots of a 32-rod group are stored level-major (k_inc_emit)
    const double *incCol; // [6][nInc]; nInc here = component stride (Context::incStride, multiple of 4)
    size_t nInc;
    int nRods;
};

// ------------------------------------------------------------------------------------------------
// f = D x, u = M f straight from global memory: one warp per 32-rod group, lane = rod.  With the level-major
// slot layout (k_inc_emit) the k-th slots of the 32 rods are adjacent, so every column / id load of a level is
// one fully coalesced request; positions come from ballot/popc on the per-lane degree.  Levels are processed
// CHUNK at a time and software-pipelined: the constraint ids of chunk c+1 are requested before the x gathers
// and column loads of chunk c, so a warp never waits on an id before it can issue loads.
// Multi-GPU, fused: rows of U that a neighbour mirrors as ghost rods are stored to the neighbour's window as well
// (NVLink remote stores from the kernel that computes them).  mir[d][r] = row of rod r on neighbour d, -1 if it
// is not mirrored there.  The halo sequence number is released by a one-thread kernel right behind this one (a
// last-CTA election inside the kernel costs every CTA a barrier + an atomic: measured +40 us per launch).
struct HaloPush {
    const int *mir[2];
    double *rem[2];
    int on;
    // k_force_vel_act releases the neighbours' halo sequence numbers itself: the last warp of the grid to finish
    // (ticket) stores them (flag[d] = nullptr: no neighbour on that side; ticket = nullptr: a separate kernel signals)
    unsigned long long *flag[2];
    unsigned long long seq;
    unsigned int *ticket;
    // k_force_vel_rec: only the CTAs whose 128-rod tile holds a mirrored rod (tileFlag[b] != 0, *nBoundary of them) read
    // the mirror indices, push, fence and take a ticket; the last of them releases the neighbours' halo flags.  The tail
    // kernel waits for those flags behind all its rows that need no halo, so nothing is gained by running the boundary
    // tiles first -- and a CTA whose tile index hangs on a load starts late.  (nullptr: every CTA takes a ticket)
    const int *tileFlag, *nBoundary;
    int debug; // timing experiments (alens_set_option halo_debug)
};

// Where the force kernel takes x from.  XMODE 0: a plain vector.  XMODE 1: gamma_b = gamma o biFlag
// (ConstraintSolver.cpp:99), the flag being bit 1 of the slot code.  XMODE 2: the BBPGD iterate is never stored
// before it is used -- x = P(x_prev - alpha g_prev) (BCQPSolver.cpp:191-192, :431-459) is evaluated from the
// interleaved {x_prev, g_prev} pair of the gathered constraint (one 16-byte gather), with the arithmetic of
// k_bb_tail, which evaluates and stores the same x: both see identical bits.
struct XIn {
    const double *x;     // XMODE 0 / 1
    const double2 *xg;   // XMODE 2: {x_prev, g_prev}
    int update;          // XMODE 2: 0 = iteration 0 (x = x_prev as given), 1 = projected gradient step
    const unsigned *mask; // XMODE 2, k_force_vel_act: bit k = 0 -> x_k is certainly 0 (nullptr: unknown)
};

// x = P(xp - alpha*gp); lb = -0.1*DBL_MAX*biFlag (-0.0 for unilateral rows, ConstraintSolver.cpp:69), ub = DBL_MAX/10
__device__ __forceinline__ double bbStep(double xp, double gp, double alpha, bool bi) {
    double v = (-alpha) * gp + 1.0 * xp;
    const double lb = bi ? (-DBL_MAX * .1) : -0.0;
    v = v > lb ? v : lb;
    v = v < kHuge ? v : kHuge;
    return v;
}

template <int CHUNK, int XMODE, bool WRITE_F>
__global__ void __launch_bounds__(256)
k_force_vel_lm(FvIn in, MobIn mob, XIn xin, double *__restrict__ U, double *__restrict__ F,
               const SolverScalars *__restrict__ scal, HaloPush hp) {
    if (scal && scal->done) return;
    double alpha = 0.0;
    if (XMODE == 2) alpha = scal->alpha; // plain load, L1 broadcast

    const int lane = threadIdx.x & 31;
    const int grp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int r = grp * 32 + lane;
    const bool act = r < in.nRods;
    int b = 0, d = 0;
    if (act) {
        b = __ldg(in.incStart + r);
        d = __ldg(in.incStart + r + 1) - b;
    }
    double qx = 0, qy = 0, qz = 0, iPara = 0, iPerp = 0, iRot = 0;
    if (act) {
        qx = mob.dx[r]; qy = mob.dy[r]; qz = mob.dz[r];
        iPara = mob.invDrag[r]; iPerp = mob.invDrag[mob.stride + r]; iRot = mob.invDrag[2 * mob.stride + r];
    }
    const unsigned lt = (1u << lane) - 1;
    const int dmax = __reduce_max_sync(0xffffffffu, d);
    size_t off = (size_t)__shfl_sync(0xffffffffu, b, 0); // first slot of the group
    const double *col = in.incCol;
    const size_t S = in.nInc;
    double f[6] = {0, 0, 0, 0, 0, 0};
    // positions + ids of the first chunk
    size_t pos[CHUNK], posN[CHUNK];
    int con[CHUNK], conN[CHUNK];
#pragma unroll
    for (int q = 0; q < CHUNK; q++) {
        const unsigned m = __ballot_sync(0xffffffffu, q < d);
        pos[q] = off + __popc(m & lt);
        off += __popc(m);
        con[q] = (q < d) ? __ldg(in.incCon + pos[q]) : 0;
    }
    for (int k0 = 0; k0 < dmax; k0 += CHUNK) {
        double xv[CHUNK], cv[CHUNK][6];
        double2 xg[CHUNK];
#pragma unroll
        for (int q = 0; q < CHUNK; q++) { // x gathers of this chunk (ids arrived one chunk ago)
            xv[q] = 0.0;
            xg[q] = make_double2(0.0, 0.0);
            if (k0 + q < d) {
                const int kc = con[q] >> 2;
                if (XMODE == 2) xg[q] = ldGather2(xin.xg + kc);
                else xv[q] = __ldg(xin.x + kc);
            }
        }
#pragma unroll
        for (int q = 0; q < CHUNK; q++) // column blocks of this chunk
#pragma unroll
            for (int c = 0; c < 6; c++) cv[q][c] = (k0 + q < d) ? ldStream(col + c * S + pos[q]) : 0.0;
#pragma unroll
        for (int q = 0; q < CHUNK; q++) { // ids of the next chunk
            const int k = k0 + CHUNK + q;
            const unsigned m = __ballot_sync(0xffffffffu, k < d);
            posN[q] = off + __popc(m & lt);
            off += __popc(m);
            conN[q] = (k < d) ? __ldg(in.incCon + posN[q]) : 0;
        }
#pragma unroll
        for (int q = 0; q < CHUNK; q++) {
            if (k0 + q < d) {
                const bool bi = (con[q] & 2) != 0;
                double x;
                if (XMODE == 2) x = xin.update ? bbStep(xg[q].x, xg[q].y, alpha, bi) : xg[q].x;
                else if (XMODE == 1) x = 1.0 * xv[q] * (bi ? 1.0 : 0.0);
                else x = xv[q];
#pragma unroll
                for (int c = 0; c < 6; c++) f[c] += cv[q][c] * x;
            }
        }
#pragma unroll
        for (int q = 0; q < CHUNK; q++) {
            pos[q] = posN[q];
            con[q] = conN[q];
        }
    }
    if (act && !mob.ghost[r]) {
        const double qf = qx * f[0] + qy * f[1] + qz * f[2];
        const double px = qf * qx, py = qf * qy, pz = qf * qz;
        const double2 u0 = make_double2(iPara * px + iPerp * (f[0] - px), iPara * py + iPerp * (f[1] - py));
        const double2 u1 = make_double2(iPara * pz + iPerp * (f[2] - pz), iRot * f[3]);
        const double2 u2 = make_double2(iRot * f[4], iRot * f[5]);
        double2 *Up = reinterpret_cast<double2 *>(U + 6 * (size_t)r);
        Up[0] = u0; Up[1] = u1; Up[2] = u2;
        if (WRITE_F) {
            double2 *Fp = reinterpret_cast<double2 *>(F + 6 * (size_t)r);
            Fp[0] = make_double2(f[0], f[1]);
            Fp[1] = make_double2(f[2], f[3]);
            Fp[2] = make_double2(f[4], f[5]);
        }
        if (hp.on) {
#pragma unroll
            for (int dd = 0; dd < 2; dd++) {
                if (!hp.mir[dd]) continue;
                const int rr = hp.mir[dd][r];
                if (rr >= 0) {
                    double2 *Rp = reinterpret_cast<double2 *>(hp.rem[dd] + 6 * (size_t)rr);
                    Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// f = D x, u = M f, touching only the incidence slots whose multiplier can be non-zero.
// In a BCQP iterate most unilateral rows sit on their bound (x = 0: the rods are within colBuf of each other but not
// pushing); measured on the bench workload 84 % of the rows, every iteration.  A slot with x = 0 adds +-0 to f, which
// leaves every bit of f unchanged (f starts at +0 and a round-to-nearest sum never produces -0 from a non-zero
// cancellation), so skipping it is exact -- and neither its {x, g} pair nor its 48-byte column block is fetched.
// Which rows can be non-zero is known one kernel earlier: k_bb_tail publishes one bit per row (XIn::mask).
// One warp per 32-rod group, slots rod-major (k_inc_emit_rm), per batch of 256 slots:
//   1  slot-parallel: lane l takes slots base + l + 32 k: ids (coalesced, streaming), mask bits (a 0.4 MB array,
//      L1/L2 hits); the surviving slots are compacted, in slot order, into a warp-private queue (ballot / popc)
//   2  a lane finds the queue range of ITS rod by binary search for its first slot (the queue is sorted); its end
//      is the next lane's start
//   3  64 queued slots at a time, lane l takes entries l and l + 32: multiplier ({x_prev, g_prev} gather + the
//      projected step, or a plain x), then -- only if it is non-zero -- the column block (three 16-byte streaming
//      loads), and leaves the 6 products in shared memory; then lane = rod adds the products of its slots in
//      ascending slot order: the summation order of the level-major kernel and of the CPU restatement.
// The kernel is issue-latency bound, not bandwidth bound (ncu: ~800 warp instructions per group on 5 warps per
// scheduler in its first version), hence: few instructions per slot, no per-slot arithmetic before the mask test,
// small register footprint for many resident warps.
static constexpr int kActWarps = 4;   // warps per CTA
static constexpr int kActBatch = 256; // slots per warp and batch (8 per lane)
struct FvAct {
    const int *incStart, *incCon; // rod-major
    const double *incCol6;        // 6 doubles per slot, contiguous
    int nRods;
    int keepXG; // gather {x, g} with the L2 evict_last policy
    int pdlTrig; // signal the dependent launch right after the wait (else: implicitly at exit)
};

// multiplier of a slot from the gathered {x_prev, g_prev} pair (XMODE 2) or the gathered x (XMODE 0 / 1)
template <int XMODE>
__device__ __forceinline__ double actMultiplier(int update, double2 xg, double xv, int code, double alpha) {
    const bool bi = (code & 2) != 0;
    if (XMODE == 2) return update ? bbStep(xg.x, xg.y, alpha, bi) : xg.x;
    return XMODE == 1 ? 1.0 * xv * (bi ? 1.0 : 0.0) : xv;
}

template <int XMODE, bool WRITE_F, int MINB, bool HALO>
__global__ void __launch_bounds__(kActWarps * 32, MINB)
k_force_vel_act(FvAct in, MobIn mob, XIn xin, double *__restrict__ U, double *__restrict__ F,
                const SolverScalars *__restrict__ scal, HaloPush hp) {
    __shared__ int sCode[kActWarps][kActBatch];            // queued slot codes
    __shared__ unsigned short sSlot[kActWarps][kActBatch]; // queued slots, relative to the batch base
    __shared__ double sProd[kActWarps][2][6][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PER = kActBatch / 32;
    const int nGroups = (in.nRods + 31) >> 5;
    const int gStride = gridDim.x * kActWarps;
    int grp = blockIdx.x * kActWarps + w;
    const bool useMask = xin.mask != nullptr;
    const unsigned long long keep = (XMODE == 2 && in.keepXG) ? policyEvictLast() : 0ull;
    // Persistent warps, software-pipelined over their groups: while group g is being worked on, the slot range of
    // g + stride and the ids of its first batch are already in flight.
    int b = 0, e = 0; // slot range of my rod in the current group
    int code[PER];    // ids of the first batch of the current group
    if (grp < nGroups) { // incidence structure: constant during a solve, may be read before the predecessor is done
        const int r = grp * 32 + lane;
        b = __ldg(in.incStart + min(r, in.nRods));
        e = __ldg(in.incStart + min(r + 1, in.nRods));
        const int gb = __shfl_sync(0xffffffffu, b, 0), ge = __shfl_sync(0xffffffffu, e, 31);
        const int lim = min(ge, gb + kActBatch);
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int p = gb + 32 * k + lane;
            code[k] = p < lim ? ldStream(in.incCon + p) : -1;
        }
    }
    pdlWait(); // the iterate, the mask and the step size come from the previous kernel
    if (in.pdlTrig) pdlLaunchDependents();
    if (scal && scal->done) return;
    double alpha = 0.0;
    if (XMODE == 2) alpha = scal->alpha; // plain load, L1 broadcast
    bool pushed = false; // this lane stored into a neighbour's window
    while (grp < nGroups) {
        const int r = grp * 32 + lane;
        const bool act = r < in.nRods;
        const int grpN = grp + gStride;
        const bool more = grpN < nGroups; // warp-uniform
        int bN = 0, eN = 0;               // next group's slot range: in flight during everything below
        if (more) {
            const int rN = grpN * 32 + lane;
            bN = __ldg(in.incStart + min(rN, in.nRods));
            eN = __ldg(in.incStart + min(rN + 1, in.nRods));
        }
        double qx = 0, qy = 0, qz = 0, iPara = 0, iPerp = 0, iRot = 0;
        unsigned ghost = 1;
        int mirL = -1, mirR = -1;
        const int gb = __shfl_sync(0xffffffffu, b, 0), ge = __shfl_sync(0xffffffffu, e, 31);
        double f[6] = {0, 0, 0, 0, 0, 0};
        for (int base = gb; base == gb || base < ge; base += kActBatch) {
            const bool lastBatch = base + kActBatch >= ge;
            if (base != gb) { // further batches of a crowded group: not prefetched
                const int lim = min(ge, base + kActBatch);
#pragma unroll
                for (int k = 0; k < PER; k++) {
                    const int p = base + 32 * k + lane;
                    code[k] = p < lim ? ldStream(in.incCon + p) : -1;
                }
            }
            // ---- 1: mask test + compaction.  The ids arrive slot = 32 k + lane (coalesced loads); they are transposed
            // through shared memory so that lane l holds the 8 CONSECUTIVE slots 8 l .. 8 l + 7: survivors can then be
            // compacted in slot order with one warp prefix sum (instead of 8 ballot/popc rounds), and the queue range of
            // a rod follows from two shuffles (instead of a binary search).
            int *sc = sCode[w];
#pragma unroll
            for (int k = 0; k < PER; k++) sc[32 * k + lane] = code[k];
            __syncwarp();
            int c8[PER];
            {
                const int4 ca = *reinterpret_cast<const int4 *>(sc + 8 * lane);
                const int4 cb = *reinterpret_cast<const int4 *>(sc + 8 * lane + 4);
                c8[0] = ca.x; c8[1] = ca.y; c8[2] = ca.z; c8[3] = ca.w;
                c8[4] = cb.x; c8[5] = cb.y; c8[6] = cb.z; c8[7] = cb.w;
            }
            __syncwarp(); // everybody has its ids: the queue may overwrite the staging area
            unsigned live = 0; // bit k: slot 8 lane + k survives
            if (useMask) {
                unsigned mw[PER];
#pragma unroll
                for (int k = 0; k < PER; k++) mw[k] = c8[k] >= 0 ? __ldg(xin.mask + (c8[k] >> 7)) : 0u;
#pragma unroll
                for (int k = 0; k < PER; k++) live |= ((mw[k] >> ((c8[k] >> 2) & 31)) & 1u) << k;
            } else {
#pragma unroll
                for (int k = 0; k < PER; k++) live |= (c8[k] >= 0 ? 1u : 0u) << k;
            }
            const int cnt = __popc(live);
            int incl = cnt; // inclusive prefix sum of the lanes' survivor counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int excl = incl - cnt;
            const int qn = __shfl_sync(0xffffffffu, incl, 31);
            {
                int pos = excl;
#pragma unroll
                for (int k = 0; k < PER; k++) {
                    if ((live >> k) & 1u) {
                        sSlot[w][pos] = (unsigned short)(8 * lane + k);
                        sc[pos] = c8[k];
                        pos++;
                    }
                }
            }
            __syncwarp();
            if (lastBatch) { // the id registers are free: request the next group's first ids and this group's rod data
                if (more) {
                    const int gbN = __shfl_sync(0xffffffffu, bN, 0), geN = __shfl_sync(0xffffffffu, eN, 31);
                    const int limN = min(geN, gbN + kActBatch);
#pragma unroll
                    for (int k = 0; k < PER; k++) {
                        const int p = gbN + 32 * k + lane;
                        code[k] = p < limN ? ldStream(in.incCon + p) : -1;
                    }
                }
                if (act) { // read once per launch: streaming loads
                    qx = ldStream(mob.dx + r); qy = ldStream(mob.dy + r); qz = ldStream(mob.dz + r);
                    iPara = ldStream(mob.invDrag + r); iPerp = ldStream(mob.invDrag + mob.stride + r);
                    iRot = ldStream(mob.invDrag + 2 * mob.stride + r);
                    ghost = mob.ghost[r];
                    if (HALO && hp.on) { // where the neighbours keep this rod as a ghost (-1: not mirrored)
                        if (hp.mir[0]) mirL = hp.mir[0][r];
                        if (hp.mir[1]) mirR = hp.mir[1][r];
                    }
                }
            }
            if (qn == 0) continue;
            // ---- 2: entries [lo, hi) of the queue belong to my rod: number of survivors in front of my first / last slot
            int lo, hi;
            {
                const int bo = min(max(b - base, 0), kActBatch), eo = min(max(e - base, 0), kActBatch);
                const int lb = min(bo >> 3, 31), le = min(eo >> 3, 31);
                const int xb = __shfl_sync(0xffffffffu, excl, lb), xe = __shfl_sync(0xffffffffu, excl, le);
                const unsigned mb = __shfl_sync(0xffffffffu, live, lb), me = __shfl_sync(0xffffffffu, live, le);
                lo = bo >= kActBatch ? qn : xb + __popc(mb & ((1u << (bo & 7)) - 1u));
                hi = eo >= kActBatch ? qn : xe + __popc(me & ((1u << (eo & 7)) - 1u));
            }
            // ---- 3: multipliers, column blocks, products; two queue chunks in flight
            for (int c0 = 0; c0 < qn; c0 += 64) {
                const int i0 = c0 + lane, i1 = c0 + 32 + lane;
                // multiplier and column block of an entry are requested TOGETHER (one round trip instead of two): the
                // mask has already removed the rows that are certainly 0, few of the survivors turn out to be 0
                double2 g0 = make_double2(0.0, 0.0), g1 = g0;
                double v0 = 0.0, v1 = 0.0;
                int cd0 = 0, cd1 = 0;
                double2 a01 = g0, a23 = g0, a45 = g0, b01 = g0, b23 = g0, b45 = g0;
                if (i0 < qn) {
                    cd0 = sCode[w][i0];
                    if (XMODE == 2) g0 = keep ? ldGather2Keep(xin.xg + (cd0 >> 2), keep) : ldGather2(xin.xg + (cd0 >> 2));
                    else v0 = __ldg(xin.x + (cd0 >> 2));
                    const double2 *cp = reinterpret_cast<const double2 *>(in.incCol6 + kColRec * ((size_t)base + sSlot[w][i0]));
                    a01 = ldStream2(cp); a23 = ldStream2(cp + 1); a45 = ldStream2(cp + 2);
                }
                if (i1 < qn) {
                    cd1 = sCode[w][i1];
                    if (XMODE == 2) g1 = keep ? ldGather2Keep(xin.xg + (cd1 >> 2), keep) : ldGather2(xin.xg + (cd1 >> 2));
                    else v1 = __ldg(xin.x + (cd1 >> 2));
                    const double2 *cp = reinterpret_cast<const double2 *>(in.incCol6 + kColRec * ((size_t)base + sSlot[w][i1]));
                    b01 = ldStream2(cp); b23 = ldStream2(cp + 1); b45 = ldStream2(cp + 2);
                }
                double x0 = 0.0, x1 = 0.0;
                if (i0 < qn) x0 = actMultiplier<XMODE>(xin.update, g0, v0, cd0, alpha);
                if (i1 < qn) x1 = actMultiplier<XMODE>(xin.update, g1, v1, cd1, alpha);
                sProd[w][0][0][lane] = a01.x * x0; sProd[w][0][1][lane] = a01.y * x0;
                sProd[w][0][2][lane] = a23.x * x0; sProd[w][0][3][lane] = a23.y * x0;
                sProd[w][0][4][lane] = a45.x * x0; sProd[w][0][5][lane] = a45.y * x0;
                if (c0 + 32 < qn) {
                    sProd[w][1][0][lane] = b01.x * x1; sProd[w][1][1][lane] = b01.y * x1;
                    sProd[w][1][2][lane] = b23.x * x1; sProd[w][1][3][lane] = b23.y * x1;
                    sProd[w][1][4][lane] = b45.x * x1; sProd[w][1][5][lane] = b45.y * x1;
                }
                __syncwarp();
                const int s = max(lo, c0) - c0, t = min(hi, c0 + 64) - c0;
                for (int j = s; j < t; j++) {
#pragma unroll
                    for (int c = 0; c < 6; c++) f[c] += sProd[w][j >> 5][c][j & 31];
                }
                __syncwarp();
            }
        }
        if (act && !ghost) {
            const double qf = qx * f[0] + qy * f[1] + qz * f[2];
            const double px = qf * qx, py = qf * qy, pz = qf * qz;
            const double2 u0 = make_double2(iPara * px + iPerp * (f[0] - px), iPara * py + iPerp * (f[1] - py));
            const double2 u1 = make_double2(iPara * pz + iPerp * (f[2] - pz), iRot * f[3]);
            const double2 u2 = make_double2(iRot * f[4], iRot * f[5]);
            double2 *Up = reinterpret_cast<double2 *>(U + 6 * (size_t)r);
            Up[0] = u0; Up[1] = u1; Up[2] = u2;
            if (WRITE_F) {
                double2 *Fp = reinterpret_cast<double2 *>(F + 6 * (size_t)r);
                Fp[0] = make_double2(f[0], f[1]);
                Fp[1] = make_double2(f[2], f[3]);
                Fp[2] = make_double2(f[4], f[5]);
            }
            if (HALO && mirL >= 0) {
                double2 *Rp = reinterpret_cast<double2 *>(hp.rem[0] + 6 * (size_t)mirL);
                Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
                pushed = true;
            }
            if (HALO && mirR >= 0) {
                double2 *Rp = reinterpret_cast<double2 *>(hp.rem[1] + 6 * (size_t)mirR);
                Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
                pushed = true;
            }
        }
        if (!more) break;
        grp = grpN;
        b = bN;
        e = eN;
    }
    if (HALO && hp.on && hp.ticket) { // fused multi-GPU: the mirrored rows are out, the last CTA of the grid tells the neighbours
        if (pushed) __threadfence_system(); // a lane's remote stores are performed before its CTA takes the ticket
        __syncthreads();                    // (one ticket per CTA: 4x fewer atomics on the one counter than per warp)
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(hp.ticket, 1u);
            if (t == gridDim.x - 1) {
                *hp.ticket = 0;
                __threadfence_system();
                if (hp.flag[0]) stReleaseSys(hp.flag[0], hp.seq);
                if (hp.flag[1]) stReleaseSys(hp.flag[1], hp.seq);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// force_kernel = 3: f = D x, u = M f from per-slot records.
//   k_slot_init      thread per slot: {x, g} of the slot's constraint row into the record and the row's mask bit into the
//                    slot-ordered bitmap.  Runs once per BBPGD solve (iteration 0) and in front of every apply to a plain
//                    vector (APGD, alens_operator_apply, the gamma o biFlag apply of the split); inside the BBPGD loop
//                    k_bb_tail keeps records and bitmap current (rows that may be non-zero: 16 % of them on the bench
//                    workload, 2 scattered 16-byte stores each; bitmap bits flipped with atomics only when they change).
//   k_force_vel_rec  thread per rod: slot range -> bitmap words -> for every set bit ONE aligned 64-byte record
//                    ({x_prev, g_prev} and the column block in the same line), x = P(x_prev - alpha g_prev) as in
//                    k_bb_tail, products added in ascending slot order (the order of every other force kernel: results
//                    are bit-identical), u = M f.  No constraint ids, no mask-word gathers, no shared memory; rods without
//                    a live slot (45 %) skip their mobility data.  Algorithmic bytes: 4 (N+1) + 2 S/8 + 64 live slots +
//                    49 N_live-rods + N + 48 N.
struct SlotInit {
    const int *incCon;
    long long nInc;
    double *rec; // nullptr (rec_mode 1): only the bitmap is written, the force kernel gathers its multipliers by row id
    unsigned *slotLive;
};
template <int XMODE>
__global__ void __launch_bounds__(256) k_slot_init(SlotInit in, XIn xin) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int code = p < in.nInc ? ldStream(in.incCon + p) : -1;
    bool live = code >= 0;
    if (live && xin.mask) live = (__ldg(xin.mask + (code >> 7)) >> ((code >> 2) & 31)) & 1u;
    if (live && in.rec) {
        double2 v;
        if (XMODE == 2) v = ldGather2(xin.xg + (code >> 2));
        else {
            const double xv = __ldg(xin.x + (code >> 2));
            v = make_double2(XMODE == 1 ? 1.0 * xv * ((code & 2) ? 1.0 : 0.0) : xv, 0.0);
        }
        *reinterpret_cast<double2 *>(in.rec + 8 * (size_t)p) = v;
    }
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && p < in.nInc + 32) in.slotLive[p >> 5] = m;
}

// Rod headers (force_kernel = 3): head[r] = {first slot of rod r, live bits of its first 32 slots}.  One 8-byte load gives
// a rod thread its slot range AND its live mask (a second dependent load -- the word of the slot-ordered bitmap -- only for
// the few rods with more than 32 slots).  The bits are kept current by k_bb_tail next to the slot-ordered bitmap.
__global__ void k_rod_head_build(int nRods, const int *__restrict__ incStart, int2 *__restrict__ head) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nRods) head[r] = make_int2(incStart[r], 0);
}
__global__ void k_rod_head_init(int nRods, int2 *__restrict__ head, const unsigned *__restrict__ slotLive) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRods) return;
    const int b = head[r].x, e = head[r + 1].x;
    unsigned bits = 0;
    if (e > b) { // bits [b, min(e, b + 32)) of the slot-ordered bitmap
        const int wd = b >> 5, sh = b & 31;
        const unsigned long long two = (unsigned long long)slotLive[wd] | ((unsigned long long)slotLive[wd + 1] << 32);
        bits = (unsigned)(two >> sh);
        const int cnt = e - b;
        if (cnt < 32) bits &= (1u << cnt) - 1u;
    }
    head[r].y = (int)bits;
}

struct FvRec {
    const int *incStart;
    const double *rec;
    const unsigned *slotLive, *slotBi;
    int nRods;
    int update; // 1: multiplier = P(x - alpha g) of the record's pair (BBPGD iterations >= 1), 0: the record's x as it is
    // rec_mode 1: the record's first 8 bytes hold the slot code (4 row + 2 bilateral + side); the multiplier is gathered
    // from the row-ordered {x, g} pairs (xg) or a plain vector (x; xmode 1: times the bilateral flag)
    const double2 *xg;
    const double *x;
    int xmode;
    unsigned long long *stamp; // 8 words of this iteration (nullptr: off)
    // rec_mode 2 (MCOL): the records hold M * column.  The two applies of a solve that also return the FORCE rebuild the
    // column of a live slot from the row's geometry (n, posI / posJ), as k_inc_emit_rec does.
    const double *gn, *gpI, *gpJ;
    size_t gstride;
    const int2 *head; // rod headers {first slot, live bits of the first 32 slots}
};

// SRC: where a live slot's multiplier comes from -- 0: {x, g} inside the record (rec_mode 0), 1: the row-ordered {x, g}
// pairs, gathered by the row id in the record (rec_mode 1, BBPGD), 2: a plain vector, gathered by row id (rec_mode 1)
template <bool WRITE_F, bool HALO, int SRC, bool MCOL = false>
__global__ void __launch_bounds__(128) k_force_vel_rec(FvRec in, MobIn mob, double *__restrict__ U, double *__restrict__ F,
                                                       const SolverScalars *__restrict__ scal, HaloPush hp) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = r < in.nRods;
    int b = 0, e = 0;
    unsigned ghost = 1;
    int boundary = 1, nTicket = gridDim.x; // (needed at the very end only: these two loads are never waited for early)
    if (HALO && hp.on && hp.tileFlag) {
        boundary = __ldg(hp.tileFlag + blockIdx.x);
        nTicket = __ldg(hp.nBoundary);
    }
    if (act) { // incidence structure: constant during a solve, requested before the wait
        e = __ldg(in.incStart + r + 1);
        ghost = mob.ghost[r];
    }
    pdlWait(); // records, bitmaps and step size come from the previous kernel
    if (scal && scal->done) return;
    if (in.stamp && threadIdx.x == 0) atomicMin(in.stamp + 0, globalNs());
    const double alpha = (in.update && scal) ? scal->alpha : 0.0;
    double f[6] = {0, 0, 0, 0, 0, 0};
    bool any = false, pushed = false;
    unsigned first = 0;
    if (act) { // {first slot, live bits of the first 32 slots}: written by the tail kernel (plain load, after the wait)
        const int2 h = in.head[r];
        b = h.x;
        first = (unsigned)h.y;
    }
    if (act && e > b && !ghost) { // (a ghost rod's force is its owner's business: its slots are skipped)
        // piece 0: the header's 32 bits (slots b .. b + 31); further pieces (rods with more than 32 slots): words of the
        // slot-ordered bitmap, restricted to this rod's slots beyond the header's
        const int b2 = b + 32;
        for (int piece = 0, wd = b2 >> 5; piece == 0 || (e > b2 && wd <= (e - 1) >> 5); piece++) {
            unsigned bits;
            int lo;
            if (piece == 0) {
                bits = first;
                lo = b;
            } else {
                bits = __ldg(in.slotLive + wd);
                lo = wd << 5;
                if (b2 > lo) bits &= ~((1u << (b2 - lo)) - 1u);
                if (e < lo + 32) bits &= (1u << (e - lo)) - 1u;
                wd++;
            }
            if (!bits) continue;
            any = true;
            while (bits) { // ascending slot order
                const int q = __ffs(bits) - 1;
                bits &= bits - 1u;
                const unsigned biw = (SRC == 0 && in.update) ? __ldg(in.slotBi + ((lo + q) >> 5)) >> ((lo + q) & 31) : 0u;
                const double *cp = in.rec + 8 * (size_t)(lo + q);
                double xp, gp, c0, c1, c2, c3, c4, c5;
                ld256(cp, xp, gp, c0, c1); // the record's two sectors: {x | row id, g, col[0..1]} and {col[2..5]}
                ld256(cp + 4, c2, c3, c4, c5);
                double x;
                if (SRC == 1) { // rec_mode 1: one more (dependent) 16-byte gather, nothing written by the tail
                    const int code = (int)__double_as_longlong(xp);
                    const double2 v = ldGather2(in.xg + (code >> 2));
                    x = in.update ? bbStep(v.x, v.y, alpha, (code & 2) != 0) : v.x;
                } else if (SRC == 2) {
                    const int code = (int)__double_as_longlong(xp);
                    const double xv = __ldg(in.x + (code >> 2));
                    x = in.xmode == 1 ? 1.0 * xv * ((code & 2) ? 1.0 : 0.0) : xv;
                } else {
                    x = in.update ? bbStep(xp, gp, alpha, biw & 1u) : xp;
                }
                if (MCOL && WRITE_F) { // the column itself, from the row's geometry (2 launches per solve)
                    const int code = (int)__double_as_longlong(xp);
                    const size_t kk = (size_t)(code >> 2);
                    const bool sideJ = code & 1;
                    double gx = in.gn[kk], gy = in.gn[kk + in.gstride], gz = in.gn[kk + 2 * in.gstride];
                    const double *P = sideJ ? in.gpJ : in.gpI;
                    const double px = P[kk], py = P[kk + in.gstride], pz = P[kk + 2 * in.gstride];
                    if (sideJ) { gx = -gx; gy = -gy; gz = -gz; }
                    c0 = gx; c1 = gy; c2 = gz;
                    c3 = (gz * py - gy * pz); c4 = (gx * pz - gz * px); c5 = (gy * px - gx * py);
                }
                f[0] += c0 * x; f[1] += c1 * x; f[2] += c2 * x;
                f[3] += c3 * x; f[4] += c4 * x; f[5] += c5 * x;
            }
        }
    }
    // a rod without a live slot (45 % of them) has f = 0 and therefore u = +0 exactly and never reads its mobility data
    double qx = 0, qy = 0, qz = 0, iPara = 0;
    double2 iPR = make_double2(0.0, 0.0); // {1/zeta_perp, 1/zeta_rot}
    if ((!MCOL || WRITE_F) && any && !ghost) {
        ld256(mob.rec + 8 * (size_t)r, qx, qy, qz, iPara); // one 64-byte line per rod
        iPR = ldGather2(reinterpret_cast<const double2 *>(mob.rec + 8 * (size_t)r + 4));
    }
    if (act && !ghost) {
        double2 u0 = make_double2(0.0, 0.0), u1 = u0, u2 = u0;
        if (MCOL && !WRITE_F) { // the sums ARE the velocity
            u0 = make_double2(f[0], f[1]);
            u1 = make_double2(f[2], f[3]);
            u2 = make_double2(f[4], f[5]);
        } else if (any) {
            const double qf = qx * f[0] + qy * f[1] + qz * f[2];
            const double px = qf * qx, py = qf * qy, pz = qf * qz;
            u0 = make_double2(iPara * px + iPR.x * (f[0] - px), iPara * py + iPR.x * (f[1] - py));
            u1 = make_double2(iPara * pz + iPR.x * (f[2] - pz), iPR.y * f[3]);
            u2 = make_double2(iPR.y * f[4], iPR.y * f[5]);
        }
        // (not storing a row that is zero and stays zero -- 45 % of the rods -- was measured: 2 us slower, the state bit costs
        // more than the 22 MB of stores it saves)
        double2 *Up = reinterpret_cast<double2 *>(U + 6 * (size_t)r);
        Up[0] = u0; Up[1] = u1; Up[2] = u2;
        if (WRITE_F) {
            double2 *Fp = reinterpret_cast<double2 *>(F + 6 * (size_t)r);
            Fp[0] = make_double2(f[0], f[1]);
            Fp[1] = make_double2(f[2], f[3]);
            Fp[2] = make_double2(f[4], f[5]);
        }
        if (HALO && hp.on && boundary && !(hp.debug & 1)) { // where this rod is mirrored (boundary tiles only: 8 % of the rods' tiles)
            const int mirL = hp.mir[0] ? __ldg(hp.mir[0] + r) : -1, mirR = hp.mir[1] ? __ldg(hp.mir[1] + r) : -1;
            if (mirL >= 0) {
                double2 *Rp = reinterpret_cast<double2 *>(hp.rem[0] + 6 * (size_t)mirL);
                Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
                pushed = true;
            }
            if (mirR >= 0) {
                double2 *Rp = reinterpret_cast<double2 *>(hp.rem[1] + 6 * (size_t)mirR);
                Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
                pushed = true;
            }
        }
    }
    if (HALO && hp.on && hp.ticket) { // fused multi-GPU: the last CTA of the BOUNDARY tiles releases the neighbours' halo flags
        if (boundary && nTicket > 0) {
            if (pushed && !(hp.debug & 2)) __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned t = atomicAdd(hp.ticket, 1u);
                if ((int)t == nTicket - 1) {
                    *hp.ticket = 0;
                    __threadfence_system();
                    if (hp.flag[0]) stReleaseSys(hp.flag[0], hp.seq);
                    if (hp.flag[1]) stReleaseSys(hp.flag[1], hp.seq);
                    if (in.stamp) in.stamp[1] = globalNs();
                }
            }
        } else if (nTicket == 0 && blockIdx.x == 0 && threadIdx.x == 0) { // nothing is mirrored: the flags still advance
            if (hp.flag[0]) stReleaseSys(hp.flag[0], hp.seq);
            if (hp.flag[1]) stReleaseSys(hp.flag[1], hp.seq);
        }
    }
    if (in.stamp && threadIdx.x == 0) atomicMax(in.stamp + 2, globalNs());
}

// ------------------------------------------------------------------------------------------------
// The same f = D x over live slots as two light kernels (force_kernel = 2).  k_force_vel_act keeps one 32-rod group per
// warp in ~128 registers and its time is the serial latency of a warp's groups (ids -> mask -> {x,g} + columns -> sums,
// measured ~10 700 cycles per group at 16 resident warps per SM).  Split by data dependence instead:
//   k_slot_x      thread per incidence SLOT: id (coalesced), mask bit, {x_prev, g_prev} gather and projected step only
//                 for live rows; writes the multiplier of a slot whose x != 0 into a slot-indexed array and one bit per
//                 slot (ballot) into a slot-ordered bitmap.  ~40 registers: 48 warps per SM hide the two gathers.
//   k_rod_sum     thread per ROD: the bits of its (contiguous) slot range, and for every set bit the multiplier and the
//                 48-byte column record; sums in ascending slot order (the order of every other force kernel), u = M f.
// Identical arithmetic (product rounded, then added): bit-identical results.
struct SlotX {
    const int *incCon;
    long long nInc;
    double *slotX;        // [nInc] multiplier of a slot (only slots whose bit is set are written)
    unsigned *slotLive;   // [nInc / 32 + 1] bit p = slot p has a non-zero multiplier
    int keepXG;
};

template <int XMODE>
__global__ void __launch_bounds__(256) k_slot_x(SlotX in, XIn xin, const SolverScalars *__restrict__ scal) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int code = p < in.nInc ? ldStream(in.incCon + p) : -1; // constant during a solve: before the wait
    pdlWait();
    if (scal && scal->done) return;
    bool live = code >= 0;
    if (live && xin.mask) live = (__ldg(xin.mask + (code >> 7)) >> ((code >> 2) & 31)) & 1u;
    double x = 0.0;
    if (live) {
        const bool bi = (code & 2) != 0;
        if (XMODE == 2) {
            const double2 xg = in.keepXG ? ldGather2Keep(xin.xg + (code >> 2), policyEvictLast()) : ldGather2(xin.xg + (code >> 2));
            x = xin.update ? bbStep(xg.x, xg.y, scal->alpha, bi) : xg.x;
        } else {
            const double xv = __ldg(xin.x + (code >> 2));
            x = XMODE == 1 ? 1.0 * xv * (bi ? 1.0 : 0.0) : xv;
        }
    }
    const bool on = live && x != 0.0;
    if (on) in.slotX[p] = x;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && p < in.nInc + 32) in.slotLive[p >> 5] = m;
}

struct RodSum {
    const int *incStart;
    const double *incCol6, *slotX;
    const unsigned *slotLive;
    int nRods;
};

template <bool WRITE_F, bool HALO>
__global__ void __launch_bounds__(128) k_rod_sum(RodSum in, MobIn mob, double *__restrict__ U, double *__restrict__ F,
                                                 const SolverScalars *__restrict__ scal, HaloPush hp) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = r < in.nRods;
    int b = 0, e = 0;
    double qx = 0, qy = 0, qz = 0, iPara = 0, iPerp = 0, iRot = 0;
    unsigned ghost = 1;
    int mirL = -1, mirR = -1;
    if (act) { // incidence structure and rod data: constant during a solve, requested before the wait
        b = __ldg(in.incStart + r);
        e = __ldg(in.incStart + r + 1);
        qx = ldStream(mob.dx + r); qy = ldStream(mob.dy + r); qz = ldStream(mob.dz + r);
        iPara = ldStream(mob.invDrag + r); iPerp = ldStream(mob.invDrag + mob.stride + r);
        iRot = ldStream(mob.invDrag + 2 * mob.stride + r);
        ghost = mob.ghost[r];
        if (HALO && hp.on) {
            if (hp.mir[0]) mirL = hp.mir[0][r];
            if (hp.mir[1]) mirR = hp.mir[1][r];
        }
    }
    pdlWait();
    if (scal && scal->done) return;
    double f[6] = {0, 0, 0, 0, 0, 0};
    bool pushed = false;
    if (act && e > b) {
        for (int wd = b >> 5; wd <= (e - 1) >> 5; wd++) {
            unsigned bits = __ldg(in.slotLive + wd);
            const int lo = wd << 5;
            if (b > lo) bits &= ~((1u << (b - lo)) - 1u);
            if (e < lo + 32) bits &= (1u << (e - lo)) - 1u;
            while (bits) { // ascending slot order
                const int p = lo + __ffs(bits) - 1;
                bits &= bits - 1u;
                const double x = __ldg(in.slotX + p);
                const double2 *cp = reinterpret_cast<const double2 *>(in.incCol6 + kColRec * (size_t)p);
                const double2 c01 = ldStream2(cp), c23 = ldStream2(cp + 1), c45 = ldStream2(cp + 2);
                f[0] += c01.x * x; f[1] += c01.y * x; f[2] += c23.x * x;
                f[3] += c23.y * x; f[4] += c45.x * x; f[5] += c45.y * x;
            }
        }
    }
    if (act && !ghost) {
        const double qf = qx * f[0] + qy * f[1] + qz * f[2];
        const double px = qf * qx, py = qf * qy, pz = qf * qz;
        const double2 u0 = make_double2(iPara * px + iPerp * (f[0] - px), iPara * py + iPerp * (f[1] - py));
        const double2 u1 = make_double2(iPara * pz + iPerp * (f[2] - pz), iRot * f[3]);
        const double2 u2 = make_double2(iRot * f[4], iRot * f[5]);
        double2 *Up = reinterpret_cast<double2 *>(U + 6 * (size_t)r);
        Up[0] = u0; Up[1] = u1; Up[2] = u2;
        if (WRITE_F) {
            double2 *Fp = reinterpret_cast<double2 *>(F + 6 * (size_t)r);
            Fp[0] = make_double2(f[0], f[1]);
            Fp[1] = make_double2(f[2], f[3]);
            Fp[2] = make_double2(f[4], f[5]);
        }
        if (HALO && mirL >= 0) {
            double2 *Rp = reinterpret_cast<double2 *>(hp.rem[0] + 6 * (size_t)mirL);
            Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
            pushed = true;
        }
        if (HALO && mirR >= 0) {
            double2 *Rp = reinterpret_cast<double2 *>(hp.rem[1] + 6 * (size_t)mirR);
            Rp[0] = u0; Rp[1] = u1; Rp[2] = u2;
            pushed = true;
        }
    }
    if (HALO && hp.on && hp.ticket) { // fused multi-GPU: the last CTA of the grid releases the neighbours' halo flags
        if (pushed) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(hp.ticket, 1u);
            if (t == gridDim.x - 1) {
                *hp.ticket = 0;
                __threadfence_system();
                if (hp.flag[0]) stReleaseSys(hp.flag[0], hp.seq);
                if (hp.flag[1]) stReleaseSys(hp.flag[1], hp.seq);
            }
        }
    }
}

// row k of D^T times u
__device__ __forceinline__ double dtransRow(const ConGeom &g, size_t k, const double *__restrict__ U) {
    const double gx = g.n[k], gy = g.n[k + g.stride], gz = g.n[k + 2 * g.stride];
    double y;
    {
        const double2 *u = reinterpret_cast<const double2 *>(U + 6 * (size_t)g.idxI[k]);
        const double2 a = u[0], b = u[1], c = u[2];
        const double px = g.pI[k], py = g.pI[k + g.stride], pz = g.pI[k + 2 * g.stride];
        y = gx * a.x;
        y += gy * a.y;
        y += gz * b.x;
        y += (gz * py - gy * pz) * b.y;
        y += (gx * pz - gz * px) * c.x;
        y += (gy * px - gx * py) * c.y;
    }
    const int j = g.idxJ[k];
    if (j >= 0) {
        const double2 *u = reinterpret_cast<const double2 *>(U + 6 * (size_t)j);
        const double2 a = u[0], b = u[1], c = u[2];
        const double px = g.pJ[k], py = g.pJ[k + g.stride], pz = g.pJ[k + 2 * g.stride];
        const double hx = -gx, hy = -gy, hz = -gz;
        y += hx * a.x;
        y += hy * a.y;
        y += hz * b.x;
        y += (hz * py - hy * pz) * b.y;
        y += (hx * pz - hz * px) * c.x;
        y += (hy * px - hx * py) * c.y;
    }
    return y;
}

// plain operator tail: y = D^T u + (K^-1/dt) x   (ConstraintOperator.cpp:54-69)
__global__ void k_dtrans(long long nc, ConGeom g, const double *__restrict__ U, const double *__restrict__ x,
                         const double *__restrict__ invKdt, double *__restrict__ y) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    double v = dtransRow(g, (size_t)k, U);
    v += 1.0 * invKdt[k] * x[k];
    y[k] = v;
}

// ------------------------------------------------------------------------------------------------
// deterministic CTA reduction of (sum0, sum1, sum2, max) -> out[4]
__device__ __forceinline__ void blockReduce4(double s0, double s1, double s2, double mx, double out[4]) {
    __shared__ double red[kVecBlock / 32][4];
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[w][0] = s0; red[w][1] = s1; red[w][2] = s2; red[w][3] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, m = 0;
        for (int i = 0; i < kVecBlock / 32; i++) {
            a += red[i][0]; b += red[i][1]; c += red[i][2]; m = fmax(m, red[i][3]);
        }
        out[0] = a; out[1] = b; out[2] = c; out[3] = m;
    }
    __syncthreads();
}

// Dai & Fletcher eq. 2.2 projected gradient (BCQPSolver.cpp:461-497); bounds: lb = -0.1*DBL_MAX*biFlag
// (ConstraintSolver.cpp:69), ub = DBL_MAX/10 (BCQPSolver.cpp:499-510)
__device__ __forceinline__ double projGrad(double x, double g, double lbFlag, int &err) {
    const double eps = DBL_EPSILON * 100;
    const double lb = (-DBL_MAX * .1) * lbFlag, ub = kHuge;
    if (x < lb + eps) return g < 0.0 ? g : 0.0;
    if (x > ub - eps) return g > 0.0 ? g : 0.0;
    if (x > lb && x < ub) return g;
    err = 1;
    return 0.0;
}

struct ReduceArgs {
    double *mailPeer[kMaxRanks];             // my slot in each peer's mailbox of this parity
    unsigned long long *seqPeer[kMaxRanks];
    const double *mailMine;                  // [R][4] of this parity in my window
    const unsigned long long *seqMine;       // [R]
    int *err;
    int R;
    unsigned long long seq;
};
struct BbTail {
    long long nc;
    ConGeom g;
    const double *U, *b, *invKdt;
    const double2 *xgPrev; // {x, g} of the previous iteration (iteration 0: {x0, anything})
    double2 *xgOut;        // {x, g} of this iteration (iteration 0: written in place)
    const unsigned char *bi;
    double *partial; // [gridDim][4]
    SolverScalars *scal;
    double *hist;
    int histCap;
    double tol;
    int ite; // iteration number of this launch (0 = initial gradient)
    const int *tileOrder;     // multi-GPU: [nTiles] tile indices, tiles without a ghost-reading row first; [nTiles] = their number (nullptr: natural order, wait at the start)
    int pdlTrig;              // see FvAct
    int *prog;                // pinned host words {completed applies, done}: the host throttles its launches on them
    int keepXG;               // store {x, g} with the L2 evict_last policy (the force kernel gathers it next)
    unsigned *maskOut;        // bit k = 1 unless the NEXT iterate's x_k is certainly 0 (see k_bb_tail); nc/32 words
    // force_kernel = 3 (k_force_vel_rec): the tail flips a row's bits in the slot-ordered bitmap when the row's bit differs
    // from the one of the previous iteration (maskOut is read before it is overwritten; slotLive = nullptr: off) and, in
    // rec_mode 0 (rec != nullptr), refreshes {x, g} in the slot records of the rows whose bit is set
    double *rec;
    const int2 *cSlot;
    unsigned *slotLive;
    int2 *head;               // rod headers: the live bits of a rod's first 32 slots are flipped there as well
    // fused multi-GPU, tail_push: the rows of U mirrored on the neighbours are copied into their windows HERE, by every CTA
    // before its first row (two entries per thread), instead of by the force kernel that computes them: remote stores
    // issued from the force kernel cost it 8 us, here they overlap the 70 us of rows that need no halo
    const int *pushSrc[2];    // my sorted rod rows that are mirrored on the left / right neighbour
    int pushBase[2];          // first staging row over there: entry e goes to row pushBase + e (contiguous remote writes)
    int pushN[2];
    double *pushRem[2];       // the neighbours' U
    unsigned long long *pushFlag[2];
    unsigned int *pushTicket;
    unsigned long long *stamp; // 8 words of this iteration (nullptr: off)
    const unsigned char *own; // multi-rank: 1 = this rank counts the row in the dot products (nullptr = all)
    double *redOut;           // multi-rank: the reduced partials go here, k_bb_reduce finishes the step
    // fused multi-GPU variant (one rank per device): the kernel itself waits for the neighbours' ghost rows of U
    // before its first gather, and its last CTA runs the mailbox allreduce
    const unsigned long long *waitFlag[2];
    unsigned long long waitSeq;
    int fusedReduce;
    ReduceArgs red;
};

// last-CTA epilogue shared by the BBPGD tail kernels: fixed-order reduction of the per-CTA partials, then the
// scalar logic of BCQPSolver.cpp:195-233 (residual test, BB1/BB2 step, stagnation)
__device__ __forceinline__ void bbScalarStep(const BbTail &p, const double out[4]) {
    SolverScalars *sc = p.scal;
    sc->ticket = 0;
    sc->maybeRows = atomicExch(&sc->maybeAcc, 0ull); // all CTAs have added theirs (they fence before taking a ticket)
    sc->maybeSum += sc->maybeRows;
    sc->mv += 1;
    sc->ite = p.ite;
    const double res = out[3];
    const double alphaUsed = p.ite == 0 ? 0.0 : sc->alpha;
    if (sc->nhist < p.histCap) {
        double *h = p.hist + 6 * (size_t)sc->nhist;
        h[0] = 1.0 * p.ite; h[1] = 0; h[2] = 0; h[3] = alphaUsed; h[4] = res; h[5] = 1.0 * sc->mv;
    }
    sc->nhist += 1;
    sc->res = res;
    if (isinf(res) || isnan(res)) {
        sc->done = 3; // projection error
    } else if (fabs(res) < p.tol) {
        sc->done = 1;
    } else if (p.ite == 0) {
        sc->alpha = 1.0 / res; // Dai & Fletcher 2005 section 5 (BCQPSolver.cpp:183)
    } else {
        double a, b;
        if (p.ite % 2 == 0) { a = out[0]; b = out[1]; } // BB1
        else { a = out[1]; b = out[2]; }                // BB2
        if (fabs(b) < 10 * DBL_EPSILON) b += 10 * DBL_EPSILON;
        const double alpha = a / b;
        sc->dotA = a; sc->dotB = b;
        sc->alpha = alpha;
        if (alpha < DBL_EPSILON * 10) sc->done = 2; // stagnation (BCQPSolver.cpp:229-233)
    }
    if (p.stamp) p.stamp[7] = globalNs();
    if (p.prog) { // host flow control: no stream synchronisation inside the loop
        volatile int *pg = p.prog;
        pg[1] = sc->done;
        __threadfence_system();
        pg[0] = p.ite + 1;
    }
}

// all streaming operands of constraint row k: 2 ids + 9 geometry doubles + {x_prev, g_prev} + b (+ K^-1/dt) + flag
// = 113 B (121 B with K^-1), plus the 16-byte {x, g} store
struct TailRow {
    int iI, iJ;
    double gx, gy, gz, pIx, pIy, pIz, pJx, pJy, pJz, invK, b;
    double2 xg;
    unsigned char bi;
    int2 sl;     // slot records of the row's two sides (force_kernel = 3)
    unsigned char own; // multi-rank: this rank counts the row in the dot products (streamed with the row, not at use)
};
__device__ __forceinline__ unsigned char ldStreamU8(const unsigned char *p) {
    unsigned v;
    asm volatile("ld.global.cs.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return (unsigned char)v;
}
template <bool HASK>
__device__ __forceinline__ void loadTailRow(const BbTail &p, size_t k, TailRow &r) {
    const size_t S = p.g.stride;
    r.iI = ldStream(p.g.idxI + k); r.iJ = ldStream(p.g.idxJ + k);
    r.gx = ldStream(p.g.n + k); r.gy = ldStream(p.g.n + k + S); r.gz = ldStream(p.g.n + k + 2 * S);
    r.pIx = ldStream(p.g.pI + k); r.pIy = ldStream(p.g.pI + k + S); r.pIz = ldStream(p.g.pI + k + 2 * S);
    r.pJx = ldStream(p.g.pJ + k); r.pJy = ldStream(p.g.pJ + k + S); r.pJz = ldStream(p.g.pJ + k + 2 * S);
    r.xg = ldStream2(p.xgPrev + k);
    r.b = ldStream(p.b + k);
    r.invK = HASK ? ldStream(p.invKdt + k) : 0.0;
    r.bi = ldStreamU8(p.bi + k);
    r.own = p.own ? ldStreamU8(p.own + k) : (unsigned char)1;
    if (p.slotLive) {
        if (p.rec) { // rec_mode 0: the slots are needed for every row that may be non-zero: streamed with the row
            const int2 *sp = p.cSlot + k;
            asm volatile("ld.global.cs.v2.s32 {%0, %1}, [%2];" : "=r"(r.sl.x), "=r"(r.sl.y) : "l"(sp));
        }
    }
}

// x = P(x_prev - alpha g_prev), g = A x + b, residual, BB dots; the last CTA to finish turns the partials into the
// next step size (BCQPSolver.cpp:191-233) -- one launch replaces ~10 vector passes and 3 allreduces.
// Persistent grid-stride kernel, software-pipelined: the streaming operands of the NEXT row are requested
// before the rod velocities of the CURRENT row are gathered, so a thread always has one DRAM round trip
// and one L2 round trip (96 B) in flight and never waits on an index before issuing loads.
// HASK = false: no row has a finite stiffness (K^-1 = 0 everywhere: collision-only pools), the K^-1 x term and its
// 8 B/row are skipped.
// one row of the tail: x = P(x_prev - alpha g_prev), y = row k of D^T times U (+ K^-1 x), g = y + b, the row's share of the
// residual and of the BB dot products; stores {x, g}; returns the row's "may be non-zero next time" bit
// a row's liveness flipped: its two slot bits in the slot-ordered bitmap and, for a slot among the first 32 of its rod, in
// the rod's header.  Out of line: a few thousand rows per launch take this path, and k_bb_tail's loop is at its register limit.
// (pointers, not the parameter struct: a reference to it would put a copy of all kernel parameters on the stack)
__device__ __noinline__ void tailFlipBits(int2 *head, unsigned *slotLive, int sI, int sJ, int iI, int iJ, bool on) {
    const int jI = sI >= 0 ? sI - head[iI].x : 32, jJ = sJ >= 0 ? sJ - head[iJ].x : 32;
    unsigned *hI = reinterpret_cast<unsigned *>(&head[iI].y);
    unsigned *hJ = reinterpret_cast<unsigned *>(&head[sJ >= 0 ? iJ : iI].y);
    if (on) {
        if (sI >= 0) atomicOr(slotLive + (sI >> 5), 1u << (sI & 31));
        if (sJ >= 0) atomicOr(slotLive + (sJ >> 5), 1u << (sJ & 31));
        if (jI < 32) atomicOr(hI, 1u << jI);
        if (jJ < 32) atomicOr(hJ, 1u << jJ);
    } else {
        if (sI >= 0) atomicAnd(slotLive + (sI >> 5), ~(1u << (sI & 31)));
        if (sJ >= 0) atomicAnd(slotLive + (sJ >> 5), ~(1u << (sJ & 31)));
        if (jI < 32) atomicAnd(hI, ~(1u << jI));
        if (jJ < 32) atomicAnd(hJ, ~(1u << jJ));
    }
}

__device__ __noinline__ void tailFlipRow(int2 *head, unsigned *slotLive, const int2 *cSlot, long long k, int iI, int iJ, bool on) {
    const int2 sl = cSlot[k];
    tailFlipBits(head, slotLive, sl.x, sl.y, iI, iJ, on);
}

template <bool HASK>
__device__ __forceinline__ bool tailRowMath(const BbTail &p, const TailRow &cur, bool two, long long k, double alpha,
                                            const double2 &a, const double2 &b, const double2 &c, const double2 &d,
                                            const double2 &e, const double2 &f, unsigned long long keepPol, double &s0,
                                            double &s1, double &s2, double &mx) {
    const double xp = cur.xg.x, gp = cur.xg.y;
    const double x = p.ite > 0 ? bbStep(xp, gp, alpha, cur.bi != 0) : xp;
    const double gx = cur.gx, gy = cur.gy, gz = cur.gz;
    double y = gx * a.x;
    y += gy * a.y;
    y += gz * b.x;
    y += (gz * cur.pIy - gy * cur.pIz) * b.y;
    y += (gx * cur.pIz - gz * cur.pIx) * c.x;
    y += (gy * cur.pIx - gx * cur.pIy) * c.y;
    if (two) { // (the gathers do not wait for this test)
        const double hx = -gx, hy = -gy, hz = -gz;
        y += hx * d.x;
        y += hy * d.y;
        y += hz * e.x;
        y += (hz * cur.pJy - hy * cur.pJz) * e.y;
        y += (hx * cur.pJz - hz * cur.pJx) * f.x;
        y += (hy * cur.pJx - hx * cur.pJy) * f.y;
    }
    if (HASK) y += 1.0 * cur.invK * x;
    const double gk = 1.0 * cur.b + 1.0 * y;
    if (p.keepXG) stKeep2(p.xgOut + k, make_double2(x, gk), keepPol);
    else p.xgOut[k] = make_double2(x, gk);
    int err = 0;
    const double q = projGrad(x, gk, cur.bi ? 1.0 : 0.0, err);
    mx = fmax(mx, err ? INFINITY : fabs(q));
    if (p.ite > 0 && cur.own) { // a row mirrored on two ranks is counted by the owner of rod I
        const double dx = 1.0 * x + (-1.0) * xp;
        const double dg = 1.0 * gk + (-1.0) * gp;
        s0 += dx * dx;
        s1 += dx * dg;
        s2 += dg * dg;
    }
    // x_next = P(x - alpha_next g) with alpha_next > 0 (the loop stops on alpha < 10 eps): a unilateral row with x = 0 and
    // g >= 0 stays exactly 0 whatever alpha_next turns out to be -- its bit is 0.  NaN counts as "may be non-zero".
    const bool on = cur.bi != 0 || !(x == 0.0) || !(gk >= 0.0);
    if (p.slotLive) { // for k_force_vel_rec: the slot bitmap, and (rec_mode 0) the pair it will take its multiplier from
        // the row's bit of the previous iteration is a function of what this row has just read: {x_prev, g_prev} (k_bb_init
        // applies the same rule to the initial guess, with g = 0)
        const bool was0 = cur.bi != 0 || !(xp == 0.0) || !(gp >= 0.0);
        if (p.rec) { // rec_mode 0: the slots came with the row
            const int sI = cur.sl.x, sJ = cur.sl.y;
            if (on) {
                // the whole first 32-byte sector of the record {x, g, col[0], col[1]} = {x, g, +-n_x, +-n_y} in ONE 256-bit
                // store (sm_100): a full-sector write needs no read-for-merge in L2, a 16-byte one would
                if (sI >= 0) st256(p.rec + 8 * (size_t)sI, x, gk, gx, gy);
                if (sJ >= 0) st256(p.rec + 8 * (size_t)sJ, x, gk, -gx, -gy);
            }
            if (on != was0) tailFlipBits(p.head, p.slotLive, sI, sJ, cur.iI, cur.iJ, on);
        } else if (on != was0) { // rec_mode 1 / 2: the slots of the few rows whose bit flips are gathered out of line
            tailFlipRow(p.head, p.slotLive, p.cSlot, k, cur.iI, cur.iJ, on);
        }
    }
    return on;
}

// end of a tail kernel: CTA partials, last-CTA election, fixed-order reduction, (multi-rank allreduce,) scalar step
__device__ __forceinline__ void tailEpilogue(const BbTail &p, double s0, double s1, double s2, double mx, int nMaybe) {
    __shared__ double out[4];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0 && nMaybe) atomicAdd(&p.scal->maybeAcc, (unsigned long long)nMaybe);
    blockReduce4(s0, s1, s2, mx, out);
    if (threadIdx.x == 0) {
        double *dst = p.partial + 4 * (size_t)blockIdx.x;
        dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; dst[3] = out[3];
        __threadfence();
        const unsigned t = atomicAdd(&p.scal->ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (p.stamp && threadIdx.x == 0) p.stamp[6] = globalNs();
    // fixed-order reduction of the per-CTA partials
    s0 = s1 = s2 = mx = 0;
    const volatile double *pp = p.partial;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kVecBlock) {
        s0 += pp[4 * i]; s1 += pp[4 * i + 1]; s2 += pp[4 * i + 2]; mx = fmax(mx, pp[4 * i + 3]);
    }
    blockReduce4(s0, s1, s2, mx, out);
    if (p.fusedReduce) { // allreduce over the ranks' mailboxes: thread q talks to rank q, thread 0 sums in rank order
        __shared__ double tot4[kMaxRanks][4];
        __shared__ int failed;
        const ReduceArgs &a = p.red;
        if (threadIdx.x == 0) failed = 0;
        __syncthreads();
        if ((int)threadIdx.x < a.R) { // R remote stores + releases in parallel (one NVLink round trip instead of R)
            const int q = threadIdx.x;
            double *dst = a.mailPeer[q];
            dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; dst[3] = out[3];
            stReleaseSys(a.seqPeer[q], a.seq); // (release: this thread's four stores are ordered before the flag; no extra fence)
            if (!waitSeq(a.seqMine + q, a.seq, a.err)) {
                failed = 1;
            } else {
                const volatile double *m = a.mailMine + 4 * q;
                tot4[q][0] = m[0]; tot4[q][1] = m[1]; tot4[q][2] = m[2]; tot4[q][3] = m[3];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // a peer never arrived -- at the mailbox, or earlier at the halo wait of ANY CTA of this kernel (waitSeq sets the
            // header's error word before that CTA takes its ticket): the rows behind that wait used stale ghost velocities,
            // so stop the loop and tell the host (solveBBPGD throws ALENS_ERR_COMM on done == 4)
            if (failed || (a.err && *(volatile int *)a.err)) {
                p.scal->ticket = 0;
                p.scal->done = 4;
                if (p.prog) {
                    volatile int *pg = p.prog;
                    pg[1] = 4;
                    __threadfence_system();
                    pg[0] = p.ite + 1;
                }
                return;
            }
            double tot[4] = {0, 0, 0, 0};
            for (int q = 0; q < a.R; q++) { // identical order on every rank: identical bits
                tot[0] += tot4[q][0]; tot[1] += tot4[q][1]; tot[2] += tot4[q][2]; tot[3] = fmax(tot[3], tot4[q][3]);
            }
            bbScalarStep(p, tot);
        }
        return;
    }
    if (threadIdx.x == 0) {
        if (p.redOut) { // multi-rank, unfused: k_bb_reduce combines the ranks and takes the scalar step
            p.redOut[0] = out[0]; p.redOut[1] = out[1]; p.redOut[2] = out[2]; p.redOut[3] = out[3];
            p.scal->ticket = 0;
        } else {
            bbScalarStep(p, out);
        }
    }
}

template <bool HASK>
__global__ void __launch_bounds__(kVecBlock, 2) k_bb_tail(BbTail p) {
    const int done = p.scal->done;
    const double alpha = p.scal->alpha; // plain loads: every thread reads the same two words, which L1 broadcasts
    // Rows are walked in tiles of 256 (one per CTA trip, grid-stride).  Multi-GPU: the tiles that contain a row reading a
    // ghost rod's velocity come LAST; the wait for the neighbours' halo sits in front of the first such tile, behind all
    // the work that needs no halo (the neighbours' force kernels push their mirrored rows FIRST, see k_force_vel_rec).
    const int nTiles = (int)((p.nc + kVecBlock - 1) / kVecBlock);
    // p.tileOrder (k_tile_order, built at setup): the tiles without a ghost-reading row first (nClean of them), then the
    // others -- whatever the slab axis
    // (the tile id of trip t + 2 is requested in trip t: the streaming loads of trip t + 1 never wait for an index)
    const bool ordered = p.waitSeq && p.tileOrder;
    const int nClean = ordered ? __ldg(p.tileOrder + nTiles) : 0;
    auto tileOf = [&](int tau) -> int { return (ordered && tau < nTiles) ? __ldg(p.tileOrder + tau) : tau; };
    int tau = blockIdx.x, tauN = tau + gridDim.x;
    int tileN = tileOf(tauN);
    long long k = tau < nTiles ? (long long)tileOf(tau) * kVecBlock + threadIdx.x : p.nc;
    double s0 = 0, s1 = 0, s2 = 0, mx = 0;
    TailRow cur, nxt;
    if (k < p.nc) loadTailRow<HASK>(p, (size_t)k, cur); // all of it at least two kernels old (see pdlWait)
    pdlWait(); // U comes from the force kernel right in front
    if (p.pdlTrig) pdlLaunchDependents();
    if (done) return;
    if (p.stamp && threadIdx.x == 0) atomicMin(p.stamp + 3, globalNs());
    // Mirrored rows of U -> the neighbours' windows: the remote stores are issued now and drain while this CTA works on its
    // first rows; the fence + ticket that lets the last CTA release the neighbours' halo flags comes kPushTrips trips later
    // (a fence right here waits ~25 us for the burst of remote writes), and in any case before this CTA waits for ITS halo.
    bool pushed = false;
    int pushTrip = p.pushTicket ? 0 : -1; // -1: nothing (more) to finish
    if (p.pushTicket) {
#pragma unroll
        for (int d = 0; d < 2; d++)
            for (int e = blockIdx.x * kVecBlock + threadIdx.x; e < p.pushN[d]; e += gridDim.x * kVecBlock) {
                const double2 *src = reinterpret_cast<const double2 *>(p.U + 6 * (size_t)__ldg(p.pushSrc[d] + e));
                double2 *dst = reinterpret_cast<double2 *>(p.pushRem[d] + 6 * (size_t)(p.pushBase[d] + e));
                const double2 a = src[0], b = src[1], c = src[2];
                dst[0] = a; dst[1] = b; dst[2] = c;
                pushed = true;
            }
    }
    auto finishPush = [&]() { // CTA-uniform
        if (pushed) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned tk = atomicAdd(p.pushTicket, 1u);
            if (tk == gridDim.x - 1) {
                *p.pushTicket = 0;
                __threadfence_system();
                if (p.pushFlag[0]) stReleaseSys(p.pushFlag[0], p.waitSeq);
                if (p.pushFlag[1]) stReleaseSys(p.pushFlag[1], p.waitSeq);
                if (p.stamp) p.stamp[1] = globalNs();
            }
        }
        pushTrip = -1;
    };
    constexpr int kPushTrips = 10;
    bool waited = p.waitSeq == 0;
    const int lane = threadIdx.x & 31;
    int nMaybe = 0; // rows of this warp whose bit is set (statistics for the roofline accounting of the force kernel)
    const unsigned long long keepPol = policyEvictLast();
    while (tau < nTiles) {
        if (pushTrip >= 0 && (++pushTrip > kPushTrips || (!waited && tau >= nClean))) finishPush();
        if (!waited && tau >= nClean) { // ghost rows of U: pushed by the neighbours (CTA-uniform test)
            if (threadIdx.x < 2) { // one thread per neighbour: the two polls overlap
                if (p.stamp) atomicMin(p.stamp + 4, globalNs());
                if (p.waitFlag[threadIdx.x]) waitSeq(p.waitFlag[threadIdx.x], p.waitSeq, p.red.err);
                if (p.stamp) atomicMax(p.stamp + 5, globalNs());
            }
            __syncthreads();
            waited = true;
        }
        const bool valid = k < p.nc;
        const int tauNN = tauN + gridDim.x;
        const int tileNN = tileOf(tauNN);
        const long long kn = tauN < nTiles ? (long long)tileN * kVecBlock + threadIdx.x : p.nc;
        // (1) the six 16-byte gathers of this row's two U rows (L2 hits, needed first), unconditional and back to back:
        // a one-sided row gathers rod I twice, a lane past the end gathers row 0
        const int iI = valid ? cur.iI : 0;
        const bool two = valid && cur.iJ >= 0;
        const int iJ = two ? cur.iJ : iI;
        const double2 *uI = reinterpret_cast<const double2 *>(p.U + 6 * (size_t)iI);
        const double2 *uJ = reinterpret_cast<const double2 *>(p.U + 6 * (size_t)iJ);
        const double2 a = ldGather2(uI), b = ldGather2(uI + 1), c = ldGather2(uI + 2);
        const double2 d = ldGather2(uJ), e = ldGather2(uJ + 1), f = ldGather2(uJ + 2);
        // (2) the streaming operands of the next row (DRAM, needed one trip later)
        if (kn < p.nc) loadTailRow<HASK>(p, (size_t)kn, nxt);
        bool on = false;
        if (valid) on = tailRowMath<HASK>(p, cur, two, k, alpha, a, b, c, d, e, f, keepPol, s0, s1, s2, mx);
        // one ballot per 32 rows publishes, for the NEXT iteration's force kernel, which rows can be non-zero
        const unsigned mbits = __ballot_sync(0xffffffffu, on);
        if (lane == 0 && p.maskOut && k < p.nc) p.maskOut[k >> 5] = mbits;
        nMaybe += __popc(mbits);
        cur = nxt;
        k = kn;
        tau = tauN;
        tauN = tauNN;
        tileN = tileNN;
    }
    if (pushTrip >= 0) finishPush(); // (a CTA with fewer trips, or none)
    tailEpilogue(p, s0, s1, s2, mx, nMaybe);
}

// ------------------------------------------------------------------------------------------------
// The same tail with its 15 streaming operand arrays staged through shared memory by the TMA engine: tiles of 256
// consecutive rows, one cp.async.bulk per array and tile (SASS UBLKCP) into a ring of NST stages, completion through one
// mbarrier per stage (expect_tx = bytes of the tile).  One elected thread issues the copies of tile i + NST - 1 before
// the CTA works on tile i, so NST - 1 tiles (26 KB each) per CTA are in flight independently of what the warps do:
// the register-staged version above has at most one row per thread (113 B) in flight and pays for it in registers.
static constexpr int kTailTile = 256;
template <bool HASK>
struct TailStage {
    double n[3][kTailTile], pI[3][kTailTile], pJ[3][kTailTile];
    double2 xg[kTailTile];
    double b[kTailTile];
    double invK[HASK ? kTailTile : 2];
    int iI[kTailTile], iJ[kTailTile];
    unsigned char bi[kTailTile];
};

template <bool HASK, int NST>
__global__ void __launch_bounds__(kTailTile, 2) k_bb_tail_ring(BbTail p, int nTiles) {
    extern __shared__ __align__(128) unsigned char smRaw[];
    using Stage = TailStage<HASK>;
    Stage *stg = reinterpret_cast<Stage *>(smRaw);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smRaw + NST * sizeof(Stage));
    const int tid = threadIdx.x, lane = tid & 31;
    const int done = p.scal->done; // two kernels old (see pdlWait)
    const double alpha = p.scal->alpha;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) mbarInit(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t S = p.g.stride;
    auto issue = [&](int i) { // bulk copies of my i-th tile into stage i % NST (thread 0 only)
        const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
        if (tile >= nTiles) return;
        Stage &st = stg[i % NST];
        unsigned long long *br = &bar[i % NST];
        const size_t k0 = (size_t)tile * kTailTile;
        const unsigned cnt = (unsigned)min((long long)kTailTile, p.nc - (long long)k0);
        const unsigned c16 = (cnt + 15u) & ~15u; // sizes in multiples of 16 B; the arrays are padded accordingly
        mbarExpectTx(br, c16 * (9u * 8u + 16u + 8u + (HASK ? 8u : 0u) + 4u + 4u + 1u));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            bulkLoad(st.n[c], p.g.n + k0 + c * S, c16 * 8u, br);
            bulkLoad(st.pI[c], p.g.pI + k0 + c * S, c16 * 8u, br);
            bulkLoad(st.pJ[c], p.g.pJ + k0 + c * S, c16 * 8u, br);
        }
        bulkLoad(st.xg, p.xgPrev + k0, c16 * 16u, br);
        bulkLoad(st.b, p.b + k0, c16 * 8u, br);
        if (HASK) bulkLoad(st.invK, p.invKdt + k0, c16 * 8u, br);
        bulkLoad(st.iI, p.g.idxI + k0, c16 * 4u, br);
        bulkLoad(st.iJ, p.g.idxJ + k0, c16 * 4u, br);
        bulkLoad(st.bi, p.bi + k0, c16, br);
    };
    // everything the copies read is at least two kernels old: they may start before the force kernel in front is done
    if (tid == 0 && !done)
        for (int i = 0; i < NST - 1; i++) issue(i);
    pdlWait(); // U comes from the force kernel right in front
    if (p.pdlTrig) pdlLaunchDependents();
    if (done) return;
    if (p.waitSeq) {
        if (tid == 0) {
            if (p.waitFlag[0]) waitSeq(p.waitFlag[0], p.waitSeq, p.red.err);
            if (p.waitFlag[1]) waitSeq(p.waitFlag[1], p.waitSeq, p.red.err);
        }
        __syncthreads();
    }
    double s0 = 0, s1 = 0, s2 = 0, mx = 0;
    int nMaybe = 0;
    const unsigned long long keepPol = policyEvictLast();
    // The six U gathers of a row (L2 round trip) are requested one trip ahead, as soon as the ids of the next tile have
    // landed: a warp never waits for a gather it has just issued.
    int iI = 0, iJ = -1;
    double2 a, b, c, d, e, f;
    auto gather = [&](int i) { // ids of my row of tile i from its stage, then the gathers
        const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
        mbarWait(&bar[i % NST], (unsigned)((i / NST) & 1));
        const Stage &st = stg[i % NST];
        const bool valid = tile * kTailTile + tid < p.nc;
        iI = valid ? st.iI[tid] : 0;
        iJ = valid ? st.iJ[tid] : -1;
        const double2 *uI = reinterpret_cast<const double2 *>(p.U + 6 * (size_t)iI);
        const double2 *uJ = reinterpret_cast<const double2 *>(p.U + 6 * (size_t)(iJ >= 0 ? iJ : iI));
        a = ldGather2(uI); b = ldGather2(uI + 1); c = ldGather2(uI + 2);
        d = ldGather2(uJ); e = ldGather2(uJ + 1); f = ldGather2(uJ + 2);
    };
    if ((long long)blockIdx.x < nTiles) gather(0);
    for (int i = 0;; i++) {
        const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
        if (tile >= nTiles) break;
        if (tid == 0) issue(i + NST - 1); // into the stage the CTA left behind the barrier at the end of trip i - 1
        const Stage &st = stg[i % NST];
        const long long k = tile * kTailTile + tid;
        const bool valid = k < p.nc;
        TailRow cur;
        cur.iI = iI; cur.iJ = iJ;
        const bool two = iJ >= 0;
        const double2 a0 = a, b0 = b, c0 = c, d0 = d, e0 = e, f0 = f;
        if (tile + gridDim.x < nTiles) gather(i + 1); // next trip's gathers: in flight during this trip's arithmetic
        bool on = false;
        if (valid) {
            cur.gx = st.n[0][tid]; cur.gy = st.n[1][tid]; cur.gz = st.n[2][tid];
            cur.pIx = st.pI[0][tid]; cur.pIy = st.pI[1][tid]; cur.pIz = st.pI[2][tid];
            cur.pJx = st.pJ[0][tid]; cur.pJy = st.pJ[1][tid]; cur.pJz = st.pJ[2][tid];
            cur.xg = st.xg[tid];
            cur.b = st.b[tid];
            cur.invK = HASK ? st.invK[tid] : 0.0;
            cur.bi = st.bi[tid];
            on = tailRowMath<HASK>(p, cur, two, k, alpha, a0, b0, c0, d0, e0, f0, keepPol, s0, s1, s2, mx);
        }
        const unsigned mbits = __ballot_sync(0xffffffffu, on);
        if (lane == 0 && p.maskOut) p.maskOut[k >> 5] = mbits;
        nMaybe += __popc(mbits);
        __syncthreads(); // the stage may be refilled
    }
    tailEpilogue(p, s0, s1, s2, mx, nMaybe);
}

// bit k = x_k can contribute to D x (k_force_vel_act on a plain vector); BIONLY: gamma_b = gamma o biFlag
template <bool BIONLY>
__global__ void k_mask_from_x(long long nc, const double *__restrict__ x, const unsigned char *__restrict__ bi,
                              unsigned *__restrict__ mask) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool on = false;
    if (k < nc) on = !(x[k] == 0.0) && (!BIONLY || bi[k] != 0);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && k < nc) mask[k >> 5] = m;
}

// BBPGD keeps its iterates as interleaved {x, g} pairs: start from x0, and unpack the two newest iterates afterwards
__global__ void k_bb_init(long long nc, const double *__restrict__ x0, double2 *__restrict__ xg,
                          unsigned *__restrict__ mask, const unsigned char *__restrict__ bi) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool on = false;
    if (k < nc) {
        const double x = x0[k];
        xg[k] = make_double2(x, 0.0);
        on = !(x == 0.0) || bi[k] != 0; // k_bb_tail's rule for "may be non-zero next time" at {x0, g = 0}
    }
    const unsigned m = __ballot_sync(0xffffffffu, on); // rows of the initial guess that can contribute to D x0
    if ((threadIdx.x & 31) == 0 && k < nc && mask) mask[k >> 5] = m;
}
__global__ void k_bb_extract(long long nc, const double2 *__restrict__ xgNew, const double2 *__restrict__ xgOld,
                             double *__restrict__ xNew, double *__restrict__ xOld) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    xNew[k] = xgNew[k].x;
    if (xgOld) xOld[k] = xgOld[k].x;
}

// multi-rank end of a BBPGD iteration: every rank drops its 4 partials into every peer's mailbox (remote
// stores + system-scope release of the sequence number), then sums all mailboxes in rank order -- each rank
// performs the same additions in the same order, so alpha, the residual and `done` agree bit for bit.
// Replaces the 3 MPI allreduces per iteration of BCQPSolver.cpp:200-233.
__global__ void k_bb_reduce(ReduceArgs a, BbTail p) {
    if (p.scal->done) return;
    const int t = threadIdx.x;
    if (t < a.R) {
        double *dst = a.mailPeer[t];
        dst[0] = p.redOut[0]; dst[1] = p.redOut[1]; dst[2] = p.redOut[2]; dst[3] = p.redOut[3];
        __threadfence_system();
        stReleaseSys(a.seqPeer[t], a.seq);
    }
    __syncwarp();
    if (t != 0) return;
    double out[4] = {0, 0, 0, 0};
    for (int q = 0; q < a.R; q++) {
        if (!waitSeq(a.seqMine + q, a.seq, a.err)) {
            p.scal->done = 4;
            return;
        }
        const volatile double *m = a.mailMine + 4 * q;
        out[0] += m[0]; out[1] += m[1]; out[2] += m[2]; out[3] = fmax(out[3], m[3]);
    }
    bbScalarStep(p, out);
}

// ------------------------------------------------------------------------------------------------
// generic vector kernels for APGD (BCQPSolver.cpp:249-389)
__global__ void k_update2(long long n, double *y, double a, const double *A, double b, const double *B, double g) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = (g == 0.0 ? 0.0 : g * y[i]) + a * A[i] + b * B[i];
}
__global__ void k_fill(long long n, double *y, double v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = v;
}
__global__ void k_project(long long n, double *x, const double *lbFlag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lb = (-DBL_MAX * .1) * lbFlag[i];
    double v = x[i];
    v = v > lb ? v : lb;
    v = v < kHuge ? v : kHuge;
    x[i] = v;
}
// up to three dot products + one projected-gradient max in one pass: out = {a.b, c.d, e.f, max|q(x,g)|}
struct Dot3 {
    const double *a, *b, *c, *d, *e, *f, *x, *g, *lbFlag;
    double addB; // q uses g + addB*bvec ... unused
};
__global__ void __launch_bounds__(kVecBlock)
k_dot3(long long n, Dot3 p, double *partial, unsigned int *ticket, double *result) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double s0 = 0, s1 = 0, s2 = 0, mx = 0;
    int err = 0;
    if (i < n) {
        if (p.a) s0 = p.a[i] * p.b[i];
        if (p.c) s1 = p.c[i] * p.d[i];
        if (p.e) s2 = p.e[i] * p.f[i];
        if (p.x) {
            mx = fabs(projGrad(p.x[i], p.g[i], p.lbFlag[i], err));
            if (err) mx = INFINITY;
        }
    }
    __shared__ double out[4];
    __shared__ bool last;
    blockReduce4(s0, s1, s2, mx, out);
    if (threadIdx.x == 0) {
        double *dst = partial + 4 * (size_t)blockIdx.x;
        dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2]; dst[3] = out[3];
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    s0 = s1 = s2 = mx = 0;
    const volatile double *pp = partial;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += kVecBlock) {
        s0 += pp[4 * j]; s1 += pp[4 * j + 1]; s2 += pp[4 * j + 2]; mx = fmax(mx, pp[4 * j + 3]);
    }
    blockReduce4(s0, s1, s2, mx, out);
    if (threadIdx.x == 0) {
        result[0] = out[0]; result[1] = out[1]; result[2] = out[2]; result[3] = out[3];
        *ticket = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// results: uni/bi split (ConstraintSolver.cpp:95-106) + permutation back to the caller's rod order
// One thread per (local rod in the caller's order, component): the sorted rows are gathered (48-byte runs, every
// sector fully used by the six threads of a rod), the four result arrays are written fully coalesced.  Without
// bilateral rows the bilateral outputs are plain memsets.
__global__ void k_split_out(int nLocal, const int *__restrict__ userToSorted, const double *__restrict__ F,
                            const double *__restrict__ U, const double *__restrict__ Fb,
                            const double *__restrict__ Ub, double *__restrict__ oFU, double *__restrict__ oVU,
                            double *__restrict__ oFB, double *__restrict__ oVB) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 6LL * nLocal) return;
    const int u = (int)(e / 6), c = (int)(e - 6LL * u);
    const size_t s = 6 * (size_t)userToSorted[u] + c;
    if (Fb) {
        const double fb = Fb[s], ub = Ub[s];
        oFU[e] = 1.0 * F[s] + (-1.0) * fb;
        oVU[e] = 1.0 * U[s] + (-1.0) * ub;
        oFB[e] = fb;
        oVB[e] = ub;
    } else {
        oFU[e] = 1.0 * F[s] + (-1.0) * 0.0;
        oVU[e] = 1.0 * U[s] + (-1.0) * 0.0;
    }
}
__global__ void k_permute6_to_user(int n, const int *__restrict__ sUser, const double *__restrict__ in,
                                   double *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const size_t u = 6 * (size_t)sUser[s], r = 6 * (size_t)s;
    for (int c = 0; c < 6; c++) out[u + c] = in[r + c];
}

// sumForceVelocity + stepEuler (SylinderSystem.cpp:802-827; Sylinder.cpp:91-99; EquatnHelper.hpp:74-90)
__global__ void k_step_euler(int n, double dt, const double *__restrict__ velNC, const double *__restrict__ vU,
                             const double *__restrict__ vB, double *__restrict__ pos, double *__restrict__ quat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v[6];
    for (int c = 0; c < 6; c++) {
        const double nb = velNC ? velNC[6 * (size_t)i + c] : 0.0;
        v[c] = ((nb + 0.0) + vU[6 * (size_t)i + c]) + vB[6 * (size_t)i + c]; // velNonB + velBrown(=0) + velCol + velBi
    }
    for (int c = 0; c < 3; c++) pos[3 * (size_t)i + c] += v[c] * dt;
    const double ox = v[3], oy = v[4], oz = v[5];
    const double w = sqrt(ox * ox + oy * oy + oz * oz);
    if (w < (double)FLT_EPSILON) return;
    double *q = quat + 4 * (size_t)i; // (x,y,z,w)
    const double winv = 1 / w, sw = sin(w * dt / 2), cw = cos(w * dt / 2);
    const double s = q[3], px = q[0], py = q[1], pz = q[2];
    const double cx = oy * pz - oz * py, cy = oz * px - ox * pz, cz = ox * py - oy * px; // omega x p
    double nx = s * sw * ox * winv + cw * px + sw * winv * cx;
    double ny = s * sw * oy * winv + cw * py + sw * winv * cy;
    double nz = s * sw * oz * winv + cw * pz + sw * winv * cz;
    double nw = s * cw - (px * ox + py * oy + pz * oz) * sw * winv;
    const double nn = sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
    q[0] = nx / nn; q[1] = ny / nn; q[2] = nz / nn; q[3] = nw / nn;
}

// =================================================================================================
// host side
// =================================================================================================
static size_t mobStride(int n) { return ((size_t)n + 3) & ~(size_t)1; }
static MobIn mobIn(Context &c) {
    return MobIn{c.sDx.p, c.sDy.p, c.sDz.p, c.sInvDrag.p, c.sMobRec.p, c.nRods, mobStride(c.nRods), c.sGhost.p};
}
// fused multi-GPU with tail_push: the tail kernel gathers a ghost rod's velocity from the staging rows behind the sorted rods
// (row nRods + g for ghost g = user index nLocal + g), where the neighbour's tail kernel writes them contiguously
__global__ void k_remap_ghost_rows(long long nc, const int *__restrict__ idxI, const int *__restrict__ idxJ,
                                   const unsigned char *__restrict__ ghost, const int *__restrict__ sUser, int nLocal, int nRods,
                                   int *__restrict__ outI, int *__restrict__ outJ) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    const int i = idxI[k], j = idxJ[k];
    outI[k] = ghost[i] ? nRods + (sUser[i] - nLocal) : i;
    outJ[k] = (j >= 0 && ghost[j]) ? nRods + (sUser[j] - nLocal) : j;
}
static ConGeom conGeom(Context &c) { return ConGeom{c.cIdxI.p, c.cIdxJ.p, c.cN.p, c.cPI.p, c.cPJ.p, c.conCap}; }
static FvIn fvIn(Context &c) {
    return FvIn{c.incStart.p, c.incCon.p, c.incCol.p, (size_t)c.incStride, c.nRods};
}

void calcMobility(Context &c, double mu) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_calc_mobility: call alens_set_rods first"};
    c.viscosity = mu;
    const int n = c.nRods;
    c.sInvDrag.reserve(3 * mobStride(n) + 4);
    if (n > 0) {
        c.sMobRec.reserve(8 * (size_t)n + 8);
        k_mob_coeff<<<gridFor(n, 256), 256, 0, c.stream>>>(n, c.sLen.p, c.sRad.p, c.sImm.p, mu, c.sInvDrag.p,
                                                           mobStride(n), c.sDx.p, c.sDy.p, c.sDz.p, c.sMobRec.p);
        c.launches++;
    }
    ALENS_CUDA(cudaGetLastError());
    c.haveMob = true;
    if (c.haveSetup && c.incLayout == 3 && c.recMode == 2) c.recDirty = true; // the slot records hold M * column
}

void mobilityApply(Context &c, const double *x, double *y) {
    if (!c.haveMob) throw ArgError{ALENS_ERR_STATE, "alens_mobility_apply: call alens_calc_mobility first"};
    const int n = c.nRods;
    if (n == 0) return;
    c.vTmp0.reserve(6 * (size_t)n);
    c.vTmp1.reserve(6 * (size_t)n);
    ALENS_CUDA(cudaMemcpyAsync(c.vTmp0.p, x, 48 * (size_t)n, cudaMemcpyHostToDevice, c.stream));
    k_mob_apply_user<<<gridFor(n, 256), 256, 0, c.stream>>>(mobIn(c), c.userToSorted.p, c.vTmp0.p, c.vTmp1.p);
    c.launches++;
    ALENS_CUDA(cudaMemcpyAsync(y, c.vTmp1.p, 48 * (size_t)n, cudaMemcpyDeviceToHost, c.stream));
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
}

void calcVelocityBrown(Context &c, double kBT, double dt, const double *normals12, unsigned long long seed,
                       unsigned long long step, double *out) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_calc_velocity_brown: call alens_set_rods first"};
    if (!c.haveMob) throw ArgError{ALENS_ERR_STATE, "alens_calc_velocity_brown: call alens_calc_mobility first"};
    if (!(dt > 0) || !(kBT >= 0)) throw ArgError{ALENS_ERR_ARG, "alens_calc_velocity_brown: dt > 0 and kBT >= 0 required"};
    if (!out) throw ArgError{ALENS_ERR_ARG, "alens_calc_velocity_brown: NULL output"};
    cudaStream_t st = c.stream;
    const int n = c.nLocal;
    if (n == 0) return;
    c.vTmp0.reserve(12 * (size_t)n + 12);
    c.vTmp1.reserve(6 * (size_t)n + 6);
    if (normals12) ALENS_CUDA(cudaMemcpyAsync(c.vTmp0.p, normals12, 96 * (size_t)n, cudaMemcpyHostToDevice, st));
    k_velocity_brown<<<gridFor(n, 128), 128, 0, st>>>(n, c.userToSorted.p, c.uGid.p, c.uQuat.p, c.sInvDrag.p,
                                                      mobStride(c.nRods), kBT, dt, normals12 ? c.vTmp0.p : nullptr, seed,
                                                      step, c.vTmp1.p);
    c.launches++;
    ALENS_CUDA(cudaMemcpyAsync(out, c.vTmp1.p, 48 * (size_t)n, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaStreamSynchronize(st));
}

void calcVelocityNonCon(Context &c, const double *force, const double *velNB, const double *velB, int monolayer,
                        double *velNonBOut) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "alens_calc_velocity_noncon: call alens_set_rods first"};
    if (force && !c.haveMob) throw ArgError{ALENS_ERR_STATE, "alens_calc_velocity_noncon: call alens_calc_mobility first"};
    cudaStream_t st = c.stream;
    const int n = c.nLocal;
    waitVelNC(c);
    c.uVelNC.reserve(6 * (size_t)c.nRods + 6);
    const size_t bytes = 48 * (size_t)n;
    c.vTmp0.reserve(6 * (size_t)n + 6); c.vTmp1.reserve(6 * (size_t)n + 6); c.vTmp2.reserve(6 * (size_t)n + 6);
    c.vTmp3.reserve(6 * (size_t)n + 6);
    if (n > 0) {
        if (force) ALENS_CUDA(cudaMemcpyAsync(c.vTmp0.p, force, bytes, cudaMemcpyHostToDevice, st));
        if (velNB) ALENS_CUDA(cudaMemcpyAsync(c.vTmp1.p, velNB, bytes, cudaMemcpyHostToDevice, st));
        if (velB) ALENS_CUDA(cudaMemcpyAsync(c.vTmp2.p, velB, bytes, cudaMemcpyHostToDevice, st));
        k_velocity_noncon<<<gridFor(n, 128), 128, 0, st>>>(mobIn(c), n, c.userToSorted.p, force ? c.vTmp0.p : nullptr,
                                                           velNB ? c.vTmp1.p : nullptr, velB ? c.vTmp2.p : nullptr,
                                                           monolayer, c.uVelNC.p, velNonBOut ? c.vTmp3.p : nullptr);
        c.launches++;
        if (velNonBOut) ALENS_CUDA(cudaMemcpyAsync(velNonBOut, c.vTmp3.p, bytes, cudaMemcpyDeviceToHost, st));
    }
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.haveVelNC = true;
}

void setupConstraints(Context &c, const double *velNC, double dt) {
    if (!c.sorted) throw ArgError{ALENS_ERR_STATE, "setup: call alens_set_rods first"};
    if (!c.haveMob) throw ArgError{ALENS_ERR_STATE, "setup: call alens_calc_mobility first"};
    if (!(dt > 0)) throw ArgError{ALENS_ERR_ARG, "setup: dt must be > 0"};
    cudaStream_t st = c.stream;
    const int n = c.nRods;
    const long long nc = c.nCon;
    c.dt = dt;
    waitVelNC(c);
    // velNonCon (user order)
    if (velNC) { // (also on a rank that owns no rods: the ghost-row exchange below is collective)
        c.uVelNC.reserve(6 * (size_t)n + 6);
        if (c.nLocal > 0)
            ALENS_CUDA(cudaMemcpyAsync(c.uVelNC.p, velNC, 48 * (size_t)c.nLocal, cudaMemcpyHostToDevice, st));
        c.haveVelNC = true;
    }
    if (c.comm.active) { // ghost rows of velNonCon (collective: every rank must pass the same NULL / non-NULL)
        if (c.haveVelNC) {
            c.uVelNC.reserve(6 * (size_t)n + 6, st, true, 6 * (size_t)c.nLocal);
            commHaloVelNC(c);
        }
    }
    const bool useV = c.haveVelNC && n > 0;
    // incidence
    c.incDeg.reserve(n + 1);
    c.incStart.reserve(n + 8);
    c.incFill.reserve(n + 1);
    ALENS_CUDA(cudaMemsetAsync(c.incDeg.p, 0, sizeof(int) * (n + 1), st));
    ALENS_CUDA(cudaMemsetAsync(c.incFill.p, 0, sizeof(int) * (n + 1), st));
    if (nc > 0) {
        k_inc_count<<<gridFor(nc, 256), 256, 0, st>>>(nc, c.cIdxI.p, c.cIdxJ.p, c.sGhost.p, c.incDeg.p);
        c.launches++;
    }
    launchScanInt(c, c.incDeg.p, c.incStart.p, n);
    // slots = 2 per two-sided + 1 per one-sided constraint: known on the host, no readback (with ghost rods
    // the sides that fall on a ghost are skipped: read the total back)
    long long nInc = 2 * nc - c.nOneSide;
    if (c.comm.active) {
        int tot = 0;
        ALENS_CUDA(cudaMemcpyAsync(&tot, c.incStart.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        nInc = tot;
    }
    if (nInc > 0x7fffffffLL - 16 || nc >= (1LL << 29))
        throw ArgError{ALENS_ERR_UNSUPPORTED, "setup: more than 2^29 constraints / 2^31 incidence slots on one GPU"};
    c.nInc = nInc;
    c.incLayout = c.optForceKernel; // fixed for this setup: the force kernels follow the layout that was built
    c.recMode = c.optRecMode;
    c.incStride = ((nInc + 3) & ~3LL) + 4; // component stride: 16-byte aligned bulk copies may over-read < 4 slots
    c.incCon.reserve((size_t)c.incStride + 4);
    c.incRaw.reserve((size_t)nInc + 4);
    if (c.incLayout != 3) c.incCol.reserve(kColRec * (size_t)c.incStride + 8);
    const size_t vcap = (size_t)nc + 32; // (+ padding: bulk copies read whole 16-row groups)
    c.vX0.reserve(vcap); c.vX1.reserve(vcap); c.vG0.reserve(vcap); c.vG1.reserve(vcap);
    c.vB.reserve(vcap); c.vLbFlag.reserve(vcap); c.vTmp5.reserve(vcap); // vTmp5 = invKdt
    c.rU.reserve(6 * (size_t)n + 6); c.rF.reserve(6 * (size_t)n + 6);
    c.rUb.reserve(6 * (size_t)n + 6); c.rFb.reserve(6 * (size_t)n + 6);
    c.outFU.reserve(6 * (size_t)n + 6); c.outVU.reserve(6 * (size_t)n + 6);
    c.outFB.reserve(6 * (size_t)n + 6); c.outVB.reserve(6 * (size_t)n + 6);
    c.redPartial.reserve(4 * (size_t)(gridFor(std::max<long long>(nc, 1), kVecBlock) + 1));
    if (c.incLayout == 3) { // slot records (64 bytes each), per-row slot pairs, slot bitmaps
        c.incRec.reserve(8 * ((size_t)nInc + 8));
        c.cSlot.reserve((size_t)nc + 32);
        c.slotBi.reserve((size_t)(nInc >> 5) + 4);
        c.slotLive.reserve((size_t)(nInc >> 5) + 4);
        ALENS_CUDA(cudaMemsetAsync(c.cSlot.p, 0xff, sizeof(int2) * ((size_t)nc + 32), st));
        ALENS_CUDA(cudaMemsetAsync(c.slotBi.p, 0, sizeof(unsigned) * ((size_t)(nInc >> 5) + 4), st));
        ALENS_CUDA(cudaMemsetAsync(c.slotLive.p, 0, sizeof(unsigned) * ((size_t)(nInc >> 5) + 4), st));
        c.rodHead.reserve((size_t)n + 2);
        k_rod_head_build<<<gridFor(n + 1, 256), 256, 0, st>>>(n, c.incStart.p, c.rodHead.p);
        c.launches++;
    }
    if (nc > 0) {
        if (c.incLayout == 3) {
            k_inc_fill<<<gridFor(nc, 256), 256, 0, st>>>(nc, c.cIdxI.p, c.cIdxJ.p, c.cBi.p, c.sGhost.p, c.incStart.p,
                                                         c.incFill.p, c.incCon.p);
            // rec_mode 2 needs the rods' mobility for M * column; without it (setup before alens_calc_mobility) this setup
            // falls back to plain columns
            if (c.recMode == 2 && !c.haveMob) c.recMode = 1;
            k_inc_emit_rec<<<gridFor(n, 128), 128, 0, st>>>(n, c.incStart.p, c.incCon.p, conGeom(c), c.incRec.p, c.cSlot.p,
                                                            c.slotBi.p, c.recMode == 2 ? c.sMobRec.p : nullptr);
            c.recDirty = false;
        } else if (c.incLayout == 1) { // rod-major: the raw slot lists are sorted in place and ARE incCon
            k_inc_fill<<<gridFor(nc, 256), 256, 0, st>>>(nc, c.cIdxI.p, c.cIdxJ.p, c.cBi.p, c.sGhost.p, c.incStart.p,
                                                         c.incFill.p, c.incCon.p);
            k_inc_emit_rm<<<gridFor(n, 128), 128, 0, st>>>(n, c.incStart.p, c.incCon.p, conGeom(c), c.incCol.p);
        } else {
            k_inc_fill<<<gridFor(nc, 256), 256, 0, st>>>(nc, c.cIdxI.p, c.cIdxJ.p, c.cBi.p, c.sGhost.p, c.incStart.p,
                                                         c.incFill.p, c.incRaw.p);
            k_inc_emit<<<gridFor(n, 128), 128, 0, st>>>(n, c.incStart.p, c.incRaw.p, conGeom(c), c.incCon.p,
                                                        c.incCol.p, (size_t)c.incStride);
        }
        int *tileFlag = nullptr;
        const int nTiles = gridFor(nc, kVecBlock), nRodTiles = gridFor(std::max(n, 1), 128);
        if (c.comm.active) { // which tiles of rows read ghost velocities (k_bb_tail keeps them for last), which tiles of
                             // rods are mirrored on a neighbour (k_force_vel_rec computes and pushes them first)
            c.tailFlag.reserve((size_t)nTiles + 1);
            c.tailOrder.reserve((size_t)nTiles + 2);
            c.rodFlag.reserve((size_t)nRodTiles + 1);
            c.rodOrder.reserve((size_t)nRodTiles + 2);
            ALENS_CUDA(cudaMemsetAsync(c.tailFlag.p, 0, sizeof(int) * ((size_t)nTiles + 1), st));
            ALENS_CUDA(cudaMemsetAsync(c.rodFlag.p, 0, sizeof(int) * ((size_t)nRodTiles + 1), st));
            tileFlag = c.tailFlag.p;
        }
        k_setup<<<gridFor(nc, 256), 256, 0, st>>>(nc, conGeom(c), c.sUser.p, useV ? c.uVelNC.p : nullptr,
                                                  c.cDelta0.p, c.cGamma0.p, c.cInvKappa.p, c.cBi.p, 1.0 / dt,
                                                  c.vB.p, c.vTmp5.p, c.vLbFlag.p, c.vX0.p, c.sGhost.p, tileFlag);
        if (c.comm.active) {
            const Comm &m = c.comm;
            if (n > 0)
                k_rod_tile_flag<<<gridFor(n, 256), 256, 0, st>>>(n, m.left >= 0 ? m.mirror[0].p : nullptr,
                                                                 m.right >= 0 ? m.mirror[1].p : nullptr, c.rodFlag.p);
            k_tile_order<<<1, 1024, 0, st>>>(nTiles, c.tailFlag.p, 0, c.tailOrder.p);
            k_tile_order<<<1, 1024, 0, st>>>(nRodTiles, c.rodFlag.p, 1, c.rodOrder.p);
            c.cIdxIU.reserve((size_t)nc + 32);
            c.cIdxJU.reserve((size_t)nc + 32);
            k_remap_ghost_rows<<<gridFor(nc, 256), 256, 0, st>>>(nc, c.cIdxI.p, c.cIdxJ.p, c.sGhost.p, c.sUser.p, c.nLocal,
                                                               c.nRods, c.cIdxIU.p, c.cIdxJU.p);
            c.launches += 4;
        }
        c.launches += 3;
    }
    ALENS_CUDA(cudaGetLastError());
    c.haveSetup = true;
    c.haveSolution = false;
    c.xLastApplied = nullptr;
    c.xSolution = nullptr;
}

// ---- optional per-kernel timing (alens_set_profiling): event pair around a launch, summed in profFlush
static void profBegin(Context &c, int kind) {
    if (!c.profiling || c.profMute) return;
    if ((size_t)c.profUsed + 2 > c.profEv.size()) {
        const size_t old = c.profEv.size();
        c.profEv.resize(old + 64);
        for (size_t i = old; i < c.profEv.size(); i++) ALENS_CUDA(cudaEventCreate(&c.profEv[i]));
    }
    c.profKind.resize(c.profEv.size() / 2);
    c.profKind[c.profUsed / 2] = kind;
    ALENS_CUDA(cudaEventRecord(c.profEv[c.profUsed], c.stream));
}
static void profEnd(Context &c) {
    if (!c.profiling || c.profMute) return;
    ALENS_CUDA(cudaEventRecord(c.profEv[c.profUsed + 1], c.stream));
    c.profUsed += 2;
}
void profFlush(Context &c, int maxEvents) { // call after a stream synchronisation
    for (int i = 0; i < c.profUsed && i < maxEvents; i += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c.profEv[i], c.profEv[i + 1]);
        switch (c.profKind[i / 2]) {
        case 0: c.timers.op_force_vel_ms += ms; c.timers.op_force_vel_n++; break;
        case 1: c.timers.op_dtrans_ms += ms; c.timers.op_dtrans_n++; break;
        default: c.timers.op_update_ms += ms; c.timers.op_update_n++; break;
        }
    }
    c.profUsed = 0;
}

// persistent grid: as many CTAs as stay resident (each keeps 16 KB of warp queues: ask for the large shared-memory
// split), every warp walks its groups with stride gridDim * kActWarps
template <int XMODE, bool WF, int MINB, bool HALO>
static void launchForceActT(Context &c, const XIn &xin, double *U, double *F, const SolverScalars *scal,
                            const HaloPush &hp) {
    static int perSM = 0;
    if (perSM == 0) {
        cudaFuncSetAttribute((k_force_vel_act<XMODE, WF, MINB, HALO>), cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        ALENS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, (k_force_vel_act<XMODE, WF, MINB, HALO>),
                                                                 kActWarps * 32, 0));
        perSM = std::max(perSM, 1);
    }
    const int n = c.nRods;
    const FvAct fa{c.incStart.p, c.incCon.p, c.incCol.p, n, c.optKeepXG, c.optPdl == 2};
    const int grid = std::max(1, std::min(gridFor(gridFor(n, 32), kActWarps), c.numSMs * perSM * c.optForceWaves));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kActWarps * 32);
    cfg.stream = c.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c.pdlNow ? 1 : 0; // inside the BBPGD loop: overlap with the drain of the tail kernel in front
    ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_act<XMODE, WF, MINB, HALO>), fa, mobIn(c), xin, U, F, scal, hp));
}
// persistent grid: as many CTAs as stay resident (each keeps 18 KB of warp queues: ask for the large shared-memory
// split), every warp walks its groups with stride gridDim * kActWarps.  The halo code (mirror indices, remote stores,
// ticket) is only compiled into the variant the fused multi-GPU loop launches.
template <int XMODE, bool WF, int MINB>
static void launchForceAct(Context &c, const XIn &xin, double *U, double *F, const SolverScalars *scal,
                           const HaloPush &hp) {
    if (XMODE == 2 && !WF && hp.on) launchForceActT<XMODE, WF, MINB, (XMODE == 2 && !WF)>(c, xin, U, F, scal, hp);
    else launchForceActT<XMODE, WF, MINB, false>(c, xin, U, F, scal, hp);
}

// force_kernel = 2: k_slot_x (thread per slot) + k_rod_sum (thread per rod), chained with programmatic dependent launch
template <int XMODE, bool WF>
static void launchForceSplit(Context &c, const XIn &xin, double *U, double *F, const SolverScalars *scal,
                             const HaloPush &hp) {
    const int n = c.nRods;
    const long long nInc = c.nInc;
    c.slotX.reserve((size_t)nInc + 64);
    c.slotLive.reserve((size_t)(nInc >> 5) + 4);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = c.stream;
    cfg.attrs = at;
    // one warp more than needed: the bitmap word behind the last slot is written too (k_rod_sum may read it)
    cfg.gridDim = dim3((unsigned)std::max(1, gridFor(nInc + 32, 256)));
    cfg.blockDim = dim3(256);
    cfg.numAttrs = c.pdlNow ? 1 : 0;
    const SlotX sx{c.incCon.p, nInc, c.slotX.p, c.slotLive.p, c.optKeepXG};
    ALENS_CUDA(cudaLaunchKernelEx(&cfg, k_slot_x<XMODE>, sx, xin, scal));
    const RodSum rs{c.incStart.p, c.incCol.p, c.slotX.p, c.slotLive.p, n};
    cfg.gridDim = dim3((unsigned)std::max(1, gridFor(n, 128)));
    cfg.blockDim = dim3(128);
    cfg.numAttrs = c.optPdl ? 1 : 0; // its predecessor is always k_slot_x
    if (XMODE == 2 && !WF && hp.on)
        ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_rod_sum<WF, (XMODE == 2 && !WF)>), rs, mobIn(c), U, F, scal, hp));
    else ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_rod_sum<WF, false>), rs, mobIn(c), U, F, scal, hp));
}

template <int XMODE, bool WF>
static void launchForceVel(Context &c, const XIn &xin, double *U, double *F, const SolverScalars *scal,
                           const HaloPush *push = nullptr) {
    const int n = c.nRods;
    const HaloPush hp = push ? *push : HaloPush{};
    if (n == 0 && !push) return;
    profBegin(c, 0);
    if (c.incLayout == 3) { // slot records: plain vectors and iteration 0 of BBPGD first fill {x, g} and the slot bitmap
        const bool init = XMODE != 2 || !xin.update;
        const long long nInc = c.nInc;
        if (init) {
            XIn xm = xin;
            if (XMODE != 2 && c.nCon > 0) { // rows of a plain vector that can contribute
                c.vMask2.reserve((size_t)(c.nCon >> 5) + 2);
                const int g = gridFor(c.nCon, kVecBlock);
                if (XMODE == 1) k_mask_from_x<true><<<g, kVecBlock, 0, c.stream>>>(c.nCon, xin.x, c.cBi.p, c.vMask2.p);
                else k_mask_from_x<false><<<g, kVecBlock, 0, c.stream>>>(c.nCon, xin.x, c.cBi.p, c.vMask2.p);
                c.launches++;
                xm.mask = c.vMask2.p;
            }
            const SlotInit si{c.incCon.p, nInc, c.recMode == 0 ? c.incRec.p : nullptr, c.slotLive.p};
            k_slot_init<XMODE><<<std::max(1, gridFor(nInc + 32, 256)), 256, 0, c.stream>>>(si, xm);
            if (n > 0) k_rod_head_init<<<gridFor(n, 256), 256, 0, c.stream>>>(n, c.rodHead.p, c.slotLive.p);
            c.launches += 2;
            c.timers.op_launches += 2;
        }
        if (c.recDirty) { // alens_calc_mobility ran after the setup: M * column again (the sort inside is a no-op now)
            k_inc_emit_rec<<<gridFor(n, 128), 128, 0, c.stream>>>(n, c.incStart.p, c.incCon.p, conGeom(c), c.incRec.p,
                                                                  c.cSlot.p, c.slotBi.p, c.sMobRec.p);
            c.launches++;
            c.recDirty = false;
        }
        const ConGeom cg = conGeom(c);
        FvRec fr{c.incStart.p, c.incRec.p, c.slotLive.p, c.slotBi.p, n, init ? 0 : 1, nullptr, nullptr, XMODE,
                 XMODE == 2 ? c.stampNow : nullptr, cg.n, cg.pI, cg.pJ, cg.stride, c.rodHead.p};
        if (c.recMode != 0) {
            if (XMODE == 2) fr.xg = xin.xg;
            else fr.x = xin.x;
        }
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.stream = c.stream;
        cfg.attrs = at;
        cfg.gridDim = dim3((unsigned)std::max(1, gridFor(n, 128)));
        cfg.blockDim = dim3(128);
        cfg.numAttrs = c.pdlNow ? 1 : 0;
        constexpr int SRC1 = XMODE == 2 ? 1 : 2; // rec_mode 1: gather from the {x, g} pairs / from the plain vector
        if (c.recMode == 0) {
            if (XMODE == 2 && !WF && hp.on)
                ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_rec<WF, (XMODE == 2 && !WF), 0>), fr, mobIn(c), U, F, scal, hp));
            else ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_rec<WF, false, 0>), fr, mobIn(c), U, F, scal, hp));
        } else if (c.recMode == 2) {
            if (XMODE == 2 && !WF && hp.on)
                ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_rec<WF, (XMODE == 2 && !WF), SRC1, true>), fr, mobIn(c), U, F, scal, hp));
            else ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_rec<WF, false, SRC1, true>), fr, mobIn(c), U, F, scal, hp));
        } else {
            if (XMODE == 2 && !WF && hp.on)
                ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_rec<WF, (XMODE == 2 && !WF), SRC1>), fr, mobIn(c), U, F, scal, hp));
            else ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_force_vel_rec<WF, false, SRC1>), fr, mobIn(c), U, F, scal, hp));
        }
        profEnd(c);
        c.launches++;
        c.timers.op_launches++;
        return;
    }
    if (c.incLayout == 1) { // rod-major slots: active-set kernels
        XIn xm = xin;
        if (XMODE != 2 && c.optForceMask && c.nCon > 0) { // plain vector: one cheap pass marks its non-zero rows
            c.vMask2.reserve((size_t)(c.nCon >> 5) + 2);
            const int g = gridFor(c.nCon, kVecBlock);
            if (XMODE == 1) k_mask_from_x<true><<<g, kVecBlock, 0, c.stream>>>(c.nCon, xin.x, c.cBi.p, c.vMask2.p);
            else k_mask_from_x<false><<<g, kVecBlock, 0, c.stream>>>(c.nCon, xin.x, c.cBi.p, c.vMask2.p);
            c.launches++;
            xm.mask = c.vMask2.p;
        }
        if (c.optForceSplit) {
            launchForceSplit<XMODE, WF>(c, xm, U, F, scal, hp);
            profEnd(c);
            c.launches += 2;
            c.timers.op_launches += 2;
            return;
        }
        if (c.optForceMinB == 3) launchForceAct<XMODE, WF, 3>(c, xm, U, F, scal, hp);
        else if (c.optForceMinB == 5) launchForceAct<XMODE, WF, 5>(c, xm, U, F, scal, hp);
        else launchForceAct<XMODE, WF, 4>(c, xm, U, F, scal, hp);
        profEnd(c);
        c.launches++;
        c.timers.op_launches++;
        return;
    }
    const int block = c.optForceBlock;
    const int grid = std::max(1, gridFor((long long)gridFor(n, 32) * 32, block));
    if (c.optForceChunk == 4)
        k_force_vel_lm<4, XMODE, WF><<<grid, block, 0, c.stream>>>(fvIn(c), mobIn(c), xin, U, F, scal, hp);
    else
        k_force_vel_lm<2, XMODE, WF><<<grid, block, 0, c.stream>>>(fvIn(c), mobIn(c), xin, U, F, scal, hp);
    profEnd(c);
    c.launches++;
    c.timers.op_launches++;
}
static XIn xPlain(const double *x) { return XIn{x, nullptr, 0, nullptr}; }

void operatorApply(Context &c, const double *x, double *y, double *force, double *vel) {
    if (!c.haveSetup) throw ArgError{ALENS_ERR_STATE, "alens_operator_apply: call alens_setup_constraints first"};
    if (c.comm.active) throw ArgError{ALENS_ERR_UNSUPPORTED, "alens_operator_apply: single-rank test entry"};
    cudaStream_t st = c.stream;
    const long long nc = c.nCon;
    const int n = c.nRods;
    c.vTmp0.reserve((size_t)nc + 1);
    c.vTmp1.reserve((size_t)nc + 1);
    if (nc > 0) ALENS_CUDA(cudaMemcpyAsync(c.vTmp0.p, x, 8 * (size_t)nc, cudaMemcpyHostToDevice, st));
    launchForceVel<0, true>(c, xPlain(c.vTmp0.p), c.rU.p, c.rF.p, nullptr);
    if (nc > 0) {
        k_dtrans<<<gridFor(nc, kVecBlock), kVecBlock, 0, st>>>(nc, conGeom(c), c.rU.p, c.vTmp0.p, c.vTmp5.p,
                                                               c.vTmp1.p);
        c.launches++;
        ALENS_CUDA(cudaMemcpyAsync(y, c.vTmp1.p, 8 * (size_t)nc, cudaMemcpyDeviceToHost, st));
    }
    if (n > 0 && (force || vel)) {
        if (force) {
            k_permute6_to_user<<<gridFor(n, 256), 256, 0, st>>>(n, c.sUser.p, c.rF.p, c.outFU.p);
            ALENS_CUDA(cudaMemcpyAsync(force, c.outFU.p, 48 * (size_t)n, cudaMemcpyDeviceToHost, st));
        }
        if (vel) {
            k_permute6_to_user<<<gridFor(n, 256), 256, 0, st>>>(n, c.sUser.p, c.rU.p, c.outVU.p);
            ALENS_CUDA(cudaMemcpyAsync(vel, c.outVU.p, 48 * (size_t)n, cudaMemcpyDeviceToHost, st));
        }
    }
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaStreamSynchronize(st));
}

static void syncScalars(Context &c) {
    ALENS_CUDA(cudaMemcpyAsync(c.hScal, c.dScal.p, sizeof(SolverScalars), cudaMemcpyDeviceToHost, c.stream));
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
}

// BCQPSolver::solveBBPGD (BCQPSolver.cpp:134-247).  Iterations are enqueued in batches without host
// synchronisation; every kernel is a no-op once the device-side `done` flag is set, so the iterate and
// the history are exactly those of the sequential loop.
static ReduceArgs reduceArgs(Context &c) { // next mailbox round
    Comm &m = c.comm;
    const unsigned long long seq = ++m.seqMail;
    const int par = (int)(seq & 1);
    ReduceArgs a{};
    for (int q = 0; q < c.nranks; q++) {
        CommHeader *h = reinterpret_cast<CommHeader *>(m.peerWin[q]);
        a.mailPeer[q] = &h->mail[par][c.rank][0];
        a.seqPeer[q] = &h->mailSeq[par][c.rank];
    }
    CommHeader *me = reinterpret_cast<CommHeader *>(m.win);
    a.mailMine = &me->mail[par][0][0];
    a.seqMine = &me->mailSeq[par][0];
    a.err = &me->error;
    a.R = c.nranks;
    a.seq = seq;
    return a;
}

// ring depth: 4 stages of 26.3 KB (two resident CTAs: 210 KB); with the K^-1 array a stage is 28.3 KB: 3 stages
template <bool HASK>
static void launchTailRing(cudaLaunchConfig_t &cfg, const BbTail &t, int nTiles) {
    constexpr int NST = HASK ? 3 : 4;
    const size_t smem = NST * sizeof(TailStage<HASK>) + 64;
    static bool once = false;
    if (!once) {
        ALENS_CUDA(cudaFuncSetAttribute((k_bb_tail_ring<HASK, NST>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        once = true;
    }
    cfg.dynamicSmemBytes = smem;
    ALENS_CUDA(cudaLaunchKernelEx(&cfg, (k_bb_tail_ring<HASK, NST>), t, nTiles));
}

static void launchTail(Context &c, const BbTail &t, int gridTail) {
    profBegin(c, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)gridTail);
    cfg.blockDim = dim3(kVecBlock);
    cfg.stream = c.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c.pdlNow ? 1 : 0;
    if (c.optTailRing && t.nc >= 4 * kTailTile) { // TMA-staged variant: tiles of 256 rows through a shared-memory ring
        const int nTiles = (int)((t.nc + kTailTile - 1) / kTailTile);
        cfg.gridDim = dim3((unsigned)std::min(nTiles, c.numSMs * 2));
        cfg.blockDim = dim3(kTailTile);
        if (c.nBilateral > 0) launchTailRing<true>(cfg, t, nTiles);
        else launchTailRing<false>(cfg, t, nTiles);
    } else if (c.nBilateral > 0) ALENS_CUDA(cudaLaunchKernelEx(&cfg, k_bb_tail<true>, t));
    else ALENS_CUDA(cudaLaunchKernelEx(&cfg, k_bb_tail<false>, t));
    profEnd(c);
    c.launches++;
    c.timers.op_launches++;
}

static int solveBBPGD(Context &c, double tol, int maxIte) {
    cudaStream_t st = c.stream;
    const long long nc = c.nCon;
    const bool multi = c.comm.active;
    c.vXG0.reserve((size_t)nc + 32); // (+ padding: the bulk copies of k_bb_tail_ring read whole 16-row groups)
    c.vXG1.reserve((size_t)nc + 32);
    double2 *XG[2] = {c.vXG0.p, c.vXG1.p};
    c.vMask.reserve((size_t)(nc >> 5) + 2);
    const unsigned *mask = (c.optForceMask || c.incLayout == 3) ? c.vMask.p : nullptr;
    const int grid = std::max(1, gridFor(nc, kVecBlock));
    const int gridTail = std::min(grid, c.numSMs * c.optTailCtasPerSM); // persistent (2 resident CTAs per SM)
    BbTail t{};
    t.nc = nc; t.g = conGeom(c); t.U = c.rU.p; t.b = c.vB.p; t.invKdt = c.vTmp5.p; t.bi = c.cBi.p;
    t.partial = c.redPartial.p; t.scal = c.dScal.p; t.hist = c.dHist.p; t.histCap = c.histCap; t.tol = tol;
    t.maskOut = (c.optForceMask || c.incLayout == 3) ? c.vMask.p : nullptr;
    t.keepXG = c.incLayout == 3 ? 0 : c.optKeepXG; // (nothing gathers the {x, g} array when the slot records are in use)
    if (c.incLayout == 3 && nc > 0) {
        t.rec = c.recMode == 0 ? c.incRec.p : nullptr;
        t.cSlot = c.cSlot.p;
        t.slotLive = c.slotLive.p;
        t.head = c.rodHead.p;
    }
    t.pdlTrig = c.optPdl == 2;
    t.tileOrder = (multi && c.optLateHalo && c.tailOrder.p && nc > 0) ? c.tailOrder.p : nullptr;
    if (multi) {
        t.own = c.cOwn.p;
        t.redOut = reinterpret_cast<double *>(c.dCounters.p); // 4 doubles of scratch
    }
    // one BBPGD iteration = force kernel (x recomputed from {x_prev, g_prev} on the fly, f = D x, u = M f) + tail;
    // multi-rank: ghost rows of U are pushed to / awaited from the neighbours between the two kernels, and
    // k_bb_reduce replaces the last-CTA scalar step
    const bool fused = multi && c.comm.fused;
    if (c.optStamps) { // 8 words per iteration: min-stamps start at ~0ull, the others at 0
        c.stampCap = std::min(maxIte, 4096) + 2;
        c.dStamps.reserve(8 * (size_t)c.stampCap);
        std::vector<unsigned long long> init(8 * (size_t)c.stampCap, 0ull);
        for (int i = 0; i < c.stampCap; i++) init[8 * i] = init[8 * i + 3] = init[8 * i + 4] = ~0ull;
        ALENS_CUDA(cudaMemcpyAsync(c.dStamps.p, init.data(), 8 * init.size(), cudaMemcpyHostToDevice, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
    }
    auto applyAndTail = [&](const XIn &x) {
        t.stamp = (c.optStamps && t.ite < c.stampCap) ? c.dStamps.p + 8 * (size_t)t.ite : nullptr;
        c.stampNow = t.stamp;
        if (fused) {
            // the force kernel also stores the mirrored rows of U into the neighbours' windows, a one-thread kernel
            // releases their halo flags; the tail waits for its own flags before the first gather and finishes
            // with the mailbox allreduce: no separate wait / reduce launches, nothing goes through the host
            Comm &m = c.comm;
            const unsigned long long seq = ++m.seqHalo;
            CommHeader *me = reinterpret_cast<CommHeader *>(m.win);
            HaloPush hp{};
            for (int d = 0; d < 2; d++) {
                const int q = d == 0 ? m.left : m.right;
                if (q < 0) continue;
                hp.mir[d] = m.mirror[d].p;
                hp.rem[d] = reinterpret_cast<double *>(m.peerWin[q] + m.offU);
            }
            hp.on = 1;
            hp.debug = c.optHaloDebug;
            const bool tailPush = c.incLayout == 3 && c.optTailPush;
            if (tailPush) { // the tail kernel copies the mirrored rows and releases the flags (see BbTail::pushSrc)
                for (int d = 0; d < 2; d++) {
                    const int q = d == 0 ? m.left : m.right;
                    t.pushN[d] = q < 0 ? 0 : m.nSend[d];
                    t.pushSrc[d] = m.sendSorted[d].p;
                    t.pushBase[d] = m.pushBase[d];
                    t.pushRem[d] = q < 0 ? nullptr : reinterpret_cast<double *>(m.peerWin[q] + m.offU);
                    t.pushFlag[d] = q < 0 ? nullptr : &reinterpret_cast<CommHeader *>(m.peerWin[q])->haloSeq[1 - d];
                }
                t.pushTicket = &c.dScal.p->ticketFv;
                t.g.idxI = c.cIdxIU.p; // ghost rods: the staging rows
                t.g.idxJ = c.cIdxJU.p;
                launchForceVel<2, false>(c, x, c.rU.p, nullptr, c.dScal.p);
            } else if (c.incLayout != 0) { // the force kernel releases the neighbours' halo flags itself (last CTA)
                for (int d = 0; d < 2; d++) {
                    const int q = d == 0 ? m.left : m.right;
                    hp.flag[d] = q < 0 ? nullptr : &reinterpret_cast<CommHeader *>(m.peerWin[q])->haloSeq[1 - d];
                }
                hp.seq = seq;
                hp.ticket = &c.dScal.p->ticketFv;
                if (c.incLayout == 3 && c.optLateHalo && c.rodOrder.p && c.nRods > 0) {
                    hp.tileFlag = c.rodFlag.p;
                    hp.nBoundary = c.rodOrder.p + gridFor(c.nRods, 128); // k_tile_order: order[nTiles] = flagged tiles
                }
            }
            if (!tailPush) launchForceVel<2, false>(c, x, c.rU.p, nullptr, c.dScal.p, &hp);
            if (c.incLayout == 0) commSignalHalo(c, seq);
            t.waitFlag[0] = m.left >= 0 ? &me->haloSeq[0] : nullptr;
            t.waitFlag[1] = m.right >= 0 ? &me->haloSeq[1] : nullptr;
            t.waitSeq = seq;
            t.fusedReduce = 1;
            t.red = reduceArgs(c);
            launchTail(c, t, gridTail);
            return;
        }
        launchForceVel<2, false>(c, x, c.rU.p, nullptr, c.dScal.p);
        if (multi) commPushU(c, ++c.comm.seqHalo);
        launchTail(c, t, gridTail);
        if (multi) {
            k_bb_reduce<<<1, 32, 0, st>>>(reduceArgs(c), t);
            c.launches++;
        }
    };
    // host flow control through the progress words: single rank, or every rank on its own device with the fused kernels
    // (all ranks take bit-identical scalar steps, so they stop after the same iteration)
    const bool poll = (!multi || (fused && c.incLayout != 0)) && c.optPoll && c.hProg != nullptr;
    const unsigned long long halo0 = c.comm.seqHalo, mail0 = c.comm.seqMail;
    if (poll) {
        c.hProg[0] = 0; c.hProg[1] = 0;
        t.prog = c.hProgDev;
    }
    c.pdlNow = (!multi || fused) && c.optPdl && c.incLayout != 0;
    // iteration 0: g0 = A x0 + b, {x0, g0} written in place
    if (nc > 0) {
        k_bb_init<<<grid, kVecBlock, 0, st>>>(nc, c.vX0.p, XG[0], c.vMask.p, c.cBi.p);
        c.launches++;
    }
    t.ite = 0; t.xgPrev = XG[0]; t.xgOut = XG[0];
    c.profMute = poll; // iteration 0 (x0: few non-zero rows) is not part of the sample
    applyAndTail(XIn{nullptr, XG[0], 0, mask});
    c.profMute = false;
    int ite = 0;
    if (poll) {
        // Single rank: the host never synchronises inside the loop.  The last CTA of every tail kernel publishes
        // {completed applies, done} in pinned host memory; the host keeps at most `look` iterations queued behind
        // the running one and stops as soon as it sees `done` (kernels queued past that point exit at once).
        const int look = std::max(1, c.optLookahead);
        volatile int *pg = c.hProg;
        while (ite < maxIte) {
            long long spins = 0;
            bool stuck = false;
            while (pg[0] < ite + 1 - look && !pg[1]) {
                cpuRelax();
                if ((++spins & 0xfffff) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) { // the stream ran dry or failed
                    stuck = pg[0] < ite + 1 - look && !pg[1];
                    break;
                }
            }
            if (stuck) break; // syncScalars below reports the state (or the CUDA error)
            if (pg[1]) break;
            ite++;
            const int cur = (ite - 1) & 1, nxt = ite & 1;
            t.ite = ite; t.xgPrev = XG[cur]; t.xgOut = XG[nxt];
            c.profMute = (ite % kProfEvery) != 1; // event pairs only around a sample: they break the PDL overlap
            applyAndTail(XIn{nullptr, XG[cur], 1, mask});
        }
        c.profMute = false;
        syncScalars(c);
        if (multi) { // launches queued past the last executed iteration were no-ops: every rank continues from the
                     // sequence numbers of the applies that really ran (the ranks may have queued different numbers)
            c.comm.seqHalo = halo0 + (unsigned long long)c.hScal->mv;
            c.comm.seqMail = mail0 + (unsigned long long)c.hScal->mv;
        }
        // 2 launches x 2 events per sampled iteration; sampled iterations past the last executed one were no-ops
        profFlush(c, c.hScal->ite >= 1 ? 4 * ((c.hScal->ite - 1) / kProfEvery + 1) : 0);
    } else {
        const int batch = c.optBatch > 0 ? c.optBatch : (nc > 200000 || multi ? 8 : 32);
        syncScalars(c);
        profFlush(c);
        while (!c.hScal->done && ite < maxIte) {
            const int nb = std::min(batch, maxIte - ite);
            for (int b = 0; b < nb; b++) {
                ite++;
                const int cur = (ite - 1) & 1, nxt = ite & 1;
                t.ite = ite; t.xgPrev = XG[cur]; t.xgOut = XG[nxt];
                c.profMute = (ite % kProfEvery) != 1;
                applyAndTail(XIn{nullptr, XG[cur], 1, mask});
            }
            c.profMute = false;
            syncScalars(c);
            if (c.hScal->done) c.profUsed = 0; // this batch contains early-exit no-ops: not representative
            else profFlush(c);
        }
    }
    c.pdlNow = false;
    c.stampNow = nullptr;
    c.stampIters = c.optStamps ? std::min(c.hScal->ite + 1, c.stampCap) : 0;
    c.timers.op_rows_live = (long long)c.hScal->maybeSum;
    c.timers.op_applies = c.hScal->mv;
    ALENS_CUDA(cudaGetLastError());
    if (c.hScal->done == 4) throw ArgError{ALENS_ERR_COMM, "solve: timed out waiting for a peer rank"};
    if (c.comm.active) checkCommError(c); // a halo wait that timed out in any kernel of the loop (stale ghost velocities)
    const int n = c.hScal->ite; // iterations actually executed
    // unpack: vX0 = the iterate the operator last saw, vX1 = the one before it
    const bool older = !(c.hScal->done || n == 0);
    if (nc > 0) {
        k_bb_extract<<<grid, kVecBlock, 0, st>>>(nc, XG[n & 1], older ? XG[(n - 1) & 1] : nullptr, c.vX0.p, c.vX1.p);
        c.launches++;
    }
    c.lastXG = XG[(std::max(n, 1) - 1) & 1]; // {x, g} the last force kernel started from (alens_time_kernel)
    c.xLastApplied = c.vX0.p;
    c.xSolution = older ? c.vX1.p : c.vX0.p; // iteMax exit returns the older iterate (BCQPSolver.cpp:237-241)
    return c.hScal->done == 2 ? 1 : 0;
}

// instrumentation (bench.py's roofline accounting of k_force_vel_rec): slots whose bit is set in the slot bitmap and rods
// with at least one such slot, in the state the last solve / apply left behind
__global__ void k_live_stats(int nRods, const int *__restrict__ incStart, const unsigned *__restrict__ slotLive,
                             unsigned long long *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (r < nRods) {
        const int b = incStart[r], e = incStart[r + 1];
        for (int wd = b >> 5; e > b && wd <= (e - 1) >> 5; wd++) {
            unsigned bits = slotLive[wd];
            const int lo = wd << 5;
            if (b > lo) bits &= ~((1u << (b - lo)) - 1u);
            if (e < lo + 32) bits &= (1u << (e - lo)) - 1u;
            cnt += __popc(bits);
        }
    }
    const unsigned any = __ballot_sync(0xffffffffu, cnt > 0);
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(out, (unsigned long long)cnt);
        atomicAdd(out + 1, (unsigned long long)__popc(any));
    }
}
void liveStats(Context &c, long long *slots, long long *rods) {
    *slots = *rods = 0;
    if (c.incLayout != 3 || !c.haveSetup || c.nRods == 0 || c.nInc == 0) return;
    cudaStream_t st = c.stream;
    ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, 2 * sizeof(unsigned long long), st));
    k_live_stats<<<gridFor(c.nRods, 256), 256, 0, st>>>(c.nRods, c.incStart.p, c.slotLive.p, c.dCounters.p);
    unsigned long long h[2] = {0, 0};
    ALENS_CUDA(cudaMemcpyAsync(h, c.dCounters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    *slots = (long long)h[0];
    *rods = (long long)h[1];
}

// instrumentation: average device time of one BBPGD kernel on the current setup (alens_time_kernel).
// Leaves the solver scalars / iterates in an unspecified state: run a setup or solve afterwards.
double timeKernel(Context &c, int which, int reps) {
    if (!c.haveSetup || c.nCon == 0) throw ArgError{ALENS_ERR_STATE, "alens_time_kernel: call alens_setup_constraints first"};
    cudaStream_t st = c.stream;
    const long long nc = c.nCon;
    const int grid = gridFor(nc, kVecBlock);
    const int gridTail = std::min(grid, c.numSMs * c.optTailCtasPerSM);
    if (which == 3) { // the BBPGD force kernel on the iterate / mask / step size the last solve left behind
        if (!c.haveSolution || !c.lastXG) throw ArgError{ALENS_ERR_STATE, "alens_time_kernel: force_vel_last needs a BBPGD solve"};
        const bool prof = c.profiling;
        c.profiling = false;
        SolverScalars sc;
        ALENS_CUDA(cudaMemcpyAsync(&sc, c.dScal.p, sizeof(sc), cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        SolverScalars run = sc;
        run.done = 0;
        ALENS_CUDA(cudaMemcpyAsync(c.dScal.p, &run, sizeof(run), cudaMemcpyHostToDevice, st));
        const XIn xi{nullptr, c.lastXG, 1, c.optForceMask ? c.vMask.p : nullptr};
        c.rUb.reserve(6 * (size_t)c.nRods + 6);
        for (int i = 0; i < 3; i++) launchForceVel<2, false>(c, xi, c.rUb.p, nullptr, c.dScal.p);
        ALENS_CUDA(cudaEventRecord(c.ev[5], st));
        for (int i = 0; i < reps; i++) launchForceVel<2, false>(c, xi, c.rUb.p, nullptr, c.dScal.p);
        ALENS_CUDA(cudaEventRecord(c.ev[6], st));
        ALENS_CUDA(cudaMemcpyAsync(c.dScal.p, &sc, sizeof(sc), cudaMemcpyHostToDevice, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        ALENS_CUDA(cudaGetLastError());
        c.profiling = prof;
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev[5], c.ev[6]);
        return 1e3 * ms / std::max(reps, 1);
    }
    ALENS_CUDA(cudaMemsetAsync(c.dScal.p, 0, sizeof(SolverScalars), st));
    c.vXG0.reserve((size_t)nc + 1);
    c.vXG1.reserve((size_t)nc + 1);
    c.vMask.reserve((size_t)(nc >> 5) + 2);
    k_bb_init<<<grid, kVecBlock, 0, st>>>(nc, c.vX0.p, c.vXG0.p, c.vMask.p, c.cBi.p);
    BbTail t{};
    t.nc = nc; t.g = conGeom(c); t.U = c.rU.p; t.b = c.vB.p; t.invKdt = c.vTmp5.p; t.bi = c.cBi.p;
    t.partial = c.redPartial.p; t.scal = c.dScal.p; t.hist = c.dHist.p; t.histCap = 0; t.tol = -1.0;
    t.ite = 1; t.xgPrev = c.vXG0.p; t.xgOut = c.vXG1.p;
    ALENS_CUDA(cudaMemsetAsync(c.rU.p, 0, 48 * (size_t)c.nRods, st));
    const bool prof = c.profiling;
    c.profiling = false;
    if (c.incLayout == 3) { // slot records and bitmap of x0, and a tail that maintains them
        launchForceVel<2, false>(c, XIn{nullptr, c.vXG0.p, 0, c.vMask.p}, c.rU.p, nullptr, c.dScal.p);
        t.maskOut = c.vMask.p; t.rec = c.recMode == 0 ? c.incRec.p : nullptr; t.cSlot = c.cSlot.p; t.slotLive = c.slotLive.p; t.head = c.rodHead.p;
    }
    auto one = [&]() {
        if (which == 0)
            launchForceVel<2, false>(c, XIn{nullptr, c.vXG0.p, 1, c.optForceMask ? c.vMask.p : nullptr}, c.rU.p, nullptr,
                                     c.dScal.p);
        else if (which == 1) launchTail(c, t, gridTail);
        else launchForceVel<0, false>(c, xPlain(c.vX0.p), c.rU.p, nullptr, c.dScal.p);
    };
    for (int i = 0; i < 3; i++) one();
    ALENS_CUDA(cudaEventRecord(c.ev[5], st));
    for (int i = 0; i < reps; i++) one();
    ALENS_CUDA(cudaEventRecord(c.ev[6], st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    ALENS_CUDA(cudaGetLastError());
    c.profiling = prof;
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[5], c.ev[6]);
    c.haveSetup = false;
    c.haveSolution = false;
    return 1e3 * ms / std::max(reps, 1);
}

// host-driven helpers for APGD
static void dot3(Context &c, long long n, const Dot3 &p, double out[4]) {
    const int grid = gridFor(std::max<long long>(n, 1), kVecBlock);
    double *res = reinterpret_cast<double *>(c.dCounters.p); // 4 doubles of scratch
    k_dot3<<<grid, kVecBlock, 0, c.stream>>>(n, p, c.redPartial.p, &c.dScal.p->ticket, res);
    c.launches++;
    ALENS_CUDA(cudaMemcpyAsync(out, res, 32, cudaMemcpyDeviceToHost, c.stream));
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
}

static void applyA(Context &c, const double *x, double *y) { // y = A x, caches U
    launchForceVel<0, false>(c, xPlain(x), c.rU.p, nullptr, nullptr);
    k_dtrans<<<gridFor(c.nCon, kVecBlock), kVecBlock, 0, c.stream>>>(c.nCon, conGeom(c), c.rU.p, x, c.vTmp5.p, y);
    c.launches++; c.timers.op_launches++;
    c.xLastApplied = const_cast<double *>(x);
}

// for the general BCQP front end (bcqp.cu): y = A x on device vectors with the operator of the last setup
void operatorApplyDevice(Context &c, const double *x, double *y) { applyA(c, x, y); }

// BCQPSolver::solveAPGD (BCQPSolver.cpp:249-389); scalar control flow stays on the host.
static int solveAPGD(Context &c, double tol, int maxIte) {
    cudaStream_t st = c.stream;
    const long long n = c.nCon;
    const size_t vcap = (size_t)n + 1;
    c.vTmp0.reserve(vcap); c.vTmp1.reserve(vcap); c.vTmp2.reserve(vcap); c.vTmp3.reserve(vcap); c.vTmp4.reserve(vcap);
    DevBuf<double> bXk1, bYk1, bXhat, bAxb1, bXdiff; // extra work vectors (APGD is not the default path)
    bXk1.reserve(vcap); bYk1.reserve(vcap); bXhat.reserve(vcap); bAxb1.reserve(vcap); bXdiff.reserve(vcap);
    const int grid = gridFor(n, kVecBlock);
    double *xk = c.vX0.p, *yk = c.vX1.p, *xkp1 = bXk1.p, *ykp1 = bYk1.p, *gVec = c.vG0.p, *tempVec = c.vG1.p;
    double *xhatk = bXhat.p, *xkdiff = bXdiff.p, *Axb = c.vTmp0.p, *Axbkp1 = bAxb1.p;
    const double *b = c.vB.p, *lbf = c.vLbFlag.p;
    auto upd2 = [&](double *y, double a, const double *A, double bb, const double *B) {
        k_update2<<<grid, kVecBlock, 0, st>>>(n, y, a, A, bb, B, 0.0);
        c.launches++;
    };
    std::vector<double> &H = c.hist;
    H.clear();
    int mv = 0;
    ALENS_CUDA(cudaMemcpyAsync(yk, xk, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    k_fill<<<grid, kVecBlock, 0, st>>>(n, xhatk, 1.0);
    upd2(xkdiff, -1.0, xhatk, 1.0, xk);
    applyA(c, xkdiff, tempVec);
    mv++;
    double r[4];
    dot3(c, n, Dot3{tempVec, tempVec, xkdiff, xkdiff, nullptr, nullptr, nullptr, nullptr, nullptr, 0}, r);
    double Lk = sqrt(r[0]) / sqrt(r[1]);
    double tk = 1.0 / Lk;
    H.insert(H.end(), {0, 0, 0, tk, 0, 1.0 * mv});
    int ite = 0, stag = 0, perr = 0;
    double thetak = 1, thetakp1 = 1, resmin = DBL_MAX, resPhi = 0;
    while (ite < maxIte) {
        ite++;
        applyA(c, yk, Axb);
        mv++;
        upd2(gVec, 1.0, b, 1.0, Axb);
        upd2(xkp1, 1.0, yk, -tk, gVec);
        k_project<<<grid, kVecBlock, 0, st>>>(n, xkp1, lbf);
        dot3(c, n, Dot3{yk, Axb, yk, b, nullptr, nullptr, nullptr, nullptr, nullptr, 0}, r);
        const double right1 = r[0] * 0.5, right2 = r[1];
        while (true) {
            upd2(xkdiff, 1.0, xkp1, -1.0, yk);
            applyA(c, xkp1, Axbkp1);
            mv++;
            dot3(c, n, Dot3{xkp1, Axbkp1, xkp1, b, gVec, xkdiff, nullptr, nullptr, nullptr, 0}, r);
            const double left1 = r[0] * 0.5, left2 = r[1], right3 = r[2];
            double r2[4];
            dot3(c, n, Dot3{xkdiff, xkdiff, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0}, r2);
            const double right4 = 0.5 * Lk * r2[0];
            if ((left1 + left2) <= (right1 + right2 + right3 + right4)) break;
            Lk *= 2;
            tk = 1 / Lk;
            upd2(xkp1, 1.0, yk, -tk, gVec);
            k_project<<<grid, kVecBlock, 0, st>>>(n, xkp1, lbf);
        }
        if (tk < DBL_EPSILON * 10) { stag = 1; break; }
        thetakp1 = (-thetak * thetak + thetak * sqrt(4 + thetak * thetak)) / 2;
        const double betakp1 = thetak * (1 - thetak) / (thetak * thetak + thetakp1);
        upd2(ykp1, (1 + betakp1), xkp1, -betakp1, xk);
        k_update2<<<grid, kVecBlock, 0, st>>>(n, Axbkp1, 1.0, b, 0.0, b, 1.0); // Axbkp1 += b
        dot3(c, n, Dot3{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, xkp1, Axbkp1, lbf, 0}, r);
        resPhi = fabs(r[3]);
        if (std::isinf(resPhi) || std::isnan(resPhi)) { perr = 1; break; }
        if (resPhi < resmin) {
            resmin = resPhi;
            ALENS_CUDA(cudaMemcpyAsync(xhatk, xkp1, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
        }
        H.insert(H.end(), {1.0 * ite, 0, 0, tk, resPhi, 1.0 * mv});
        if (resPhi < tol) break;
        upd2(tempVec, 1.0, xkp1, -1.0, xk);
        dot3(c, n, Dot3{gVec, tempVec, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0}, r);
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
    // the solution is copied into vX0 so that the work vectors can be released
    ALENS_CUDA(cudaMemcpyAsync(c.vTmp1.p, xhatk, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    // the operator's cached force/vel belong to the last apply: keep that vector too
    ALENS_CUDA(cudaMemcpyAsync(c.vTmp2.p, c.xLastApplied, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    c.xSolution = c.vTmp1.p;
    c.xLastApplied = c.vTmp2.p;
    c.hScal->ite = ite;
    c.hScal->mv = mv;
    c.hScal->res = resPhi;
    c.hScal->alpha = tk;
    c.hScal->done = perr ? 3 : (stag ? 2 : 1);
    c.hScal->nhist = (int)(H.size() / 6);
    ALENS_CUDA(cudaGetLastError());
    return stag;
}

void solveConstraints(Context &c, double res, int maxIte, int choice) {
    if (!c.haveSetup) throw ArgError{ALENS_ERR_STATE, "solve: setup has not been run"};
    solveCore(c, res * (1.0 / c.dt), maxIte, choice); // ConstraintSolver.cpp:76
}

void solveCore(Context &c, double tol, int maxIte, int choice) {
    if (!c.haveSetup) throw ArgError{ALENS_ERR_STATE, "solve: setup has not been run"};
    cudaStream_t st = c.stream;
    const long long nc = c.nCon;
    const int n = c.nRods;
    alens_solve_report &rep = c.lastReport;
    memset(&rep, 0, sizeof(rep));
    rep.n_constraints = nc;
    rep.n_rods = c.nLocal;
    c.timers.op_launches = 0;
    c.profUsed = 0;
    c.hist.clear();
    // history capacity
    const int wantHist = (int)std::min<long long>((long long)maxIte + 2, 1 << 20);
    if (wantHist > c.histCap) {
        c.dHist.reserve(6 * (size_t)wantHist);
        c.histCap = wantHist;
    }
    ALENS_CUDA(cudaMemsetAsync(c.dScal.p, 0, sizeof(SolverScalars), st));
    ALENS_CUDA(cudaEventRecord(c.ev[2], st));
    int status = 0;
    const bool multi = c.comm.active;
    if (multi && choice == ALENS_SOLVER_APGD)
        throw ArgError{ALENS_ERR_UNSUPPORTED, "solve: APGD is single-rank only (use BBPGD with the slab decomposition)"};
    if (nc == 0 && !multi) {
        // empty problem: residual 0 < tol, zero forces (the reference returns after the first check)
        if (n > 0) {
            ALENS_CUDA(cudaMemsetAsync(c.rU.p, 0, 48 * (size_t)n, st));
            ALENS_CUDA(cudaMemsetAsync(c.rF.p, 0, 48 * (size_t)n, st));
        }
        c.hist = {0, 0, 0, 0, 0, 1};
        memset(c.hScal, 0, sizeof(SolverScalars));
        c.hScal->mv = 1; c.hScal->nhist = 1; c.hScal->done = 1;
    } else if (choice == ALENS_SOLVER_APGD) {
        status = solveAPGD(c, tol, maxIte);
    } else {
        status = solveBBPGD(c, tol, maxIte);
        const int rows = std::min(c.hScal->nhist, c.histCap);
        c.hist.resize(6 * (size_t)rows);
        if (rows) ALENS_CUDA(cudaMemcpyAsync(c.hist.data(), c.dHist.p, 48 * (size_t)rows, cudaMemcpyDeviceToHost, st));
    }
    ALENS_CUDA(cudaEventRecord(c.ev[3], st));
    // split (ConstraintSolver.cpp:95-106): force/vel of the LAST apply minus the bilateral part
    if (nc > 0 || multi) {
        launchForceVel<0, true>(c, xPlain(c.xLastApplied), c.rU.p, c.rF.p, nullptr);
        // no bilateral block in the pool: gamma_b = 0, the bilateral force/velocity are exactly zero
        if (c.nBilateral > 0) launchForceVel<1, true>(c, xPlain(c.xSolution), c.rUb.p, c.rFb.p, nullptr);
    }
    if (n > 0) {
        const bool bi = nc > 0 && c.nBilateral > 0;
        if (c.nLocal > 0) {
            if (!bi) {
                ALENS_CUDA(cudaMemsetAsync(c.outFB.p, 0, 48 * (size_t)c.nLocal, st));
                ALENS_CUDA(cudaMemsetAsync(c.outVB.p, 0, 48 * (size_t)c.nLocal, st));
            }
            k_split_out<<<gridFor(6LL * c.nLocal, 256), 256, 0, st>>>(c.nLocal, c.userToSorted.p, c.rF.p, c.rU.p,
                                                                      bi ? c.rFb.p : nullptr, bi ? c.rUb.p : nullptr,
                                                                      c.outFU.p, c.outVU.p, c.outFB.p, c.outVB.p);
            c.launches++;
        }
    }
    ALENS_CUDA(cudaEventRecord(c.ev[4], st));
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]);
    c.timers.solve_ms = ms;
    cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]);
    c.timers.split_ms = ms;
    rep.status = status;
    rep.iterations = c.hScal->ite;
    rep.matvecs = c.hScal->mv;
    rep.history_rows = (int)(c.hist.size() / 6);
    rep.residual = c.hScal->res;
    rep.step = c.hScal->alpha;
    c.haveSolution = true;
    if (c.hScal->done == 3) throw ArgError{ALENS_ERR_PROJECTION, "projection error occured (BCQPSolver.cpp:484-494)"};
}

void stepEuler(Context &c, double dt) {
    if (!c.haveSolution) throw ArgError{ALENS_ERR_STATE, "alens_step_euler: no solution available"};
    const int n = c.nLocal;
    if (n == 0) return;
    waitVelNC(c);
    k_step_euler<<<gridFor(n, 256), 256, 0, c.stream>>>(n, dt, c.haveVelNC ? c.uVelNC.p : nullptr, c.outVU.p,
                                                        c.outVB.p, c.uPos.p, c.uQuat.p);
    c.launches++;
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
}

// Force the (lazily loaded) kernels of this file into the context now: loading a kernel at its first launch can
// synchronise the context, which deadlocks against a peer rank's waiting kernel when two ranks share one GPU.
void preloadSolverKernels() {
    cudaFuncAttributes a;
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_mob_coeff));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_mob_apply_user));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_inc_count));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_inc_fill));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_inc_emit));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_setup));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_dtrans));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_bb_init));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_bb_extract));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_bb_tail_ring<true, 3>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_bb_tail_ring<false, 4>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_bb_tail<true>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_bb_tail<false>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_bb_reduce));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_update2));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_fill));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_project));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_dot3));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_split_out));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_permute6_to_user));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_step_euler));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<2, 0, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<2, 0, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<2, 1, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<2, 2, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<4, 0, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<4, 0, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<4, 1, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_lm<4, 2, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<0, false, 5, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<0, true, 5, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<1, true, 5, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<2, false, 5, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<0, false, 4, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<0, true, 4, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<1, true, 4, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<2, false, 4, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<2, false, 5, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<2, false, 4, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<0, false, 3, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<0, true, 3, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<1, true, 3, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<2, false, 3, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_act<2, false, 3, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_slot_x<0>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_slot_x<1>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_slot_x<2>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_rod_sum<false, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_rod_sum<true, false>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_rod_sum<false, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_velocity_noncon));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_velocity_brown));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_inc_emit_rm));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_mask_from_x<true>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_mask_from_x<false>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_inc_emit_rec));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_rod_head_build));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_rod_head_init));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_slot_init<0>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_slot_init<1>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_slot_init<2>));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, false, 0>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<true, false, 0>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, true, 0>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, false, 1>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, true, 1>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, false, 2>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<true, false, 2>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, false, 1, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, true, 1, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<false, false, 2, true>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_force_vel_rec<true, false, 2, true>)));
}

} // namespace alens
