// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin driver around the reference's OWN sources, compiled in place from
// /root/reference (never copied into this repo):
//   * SimToolbox/Collision/DCPQuery.hpp      -- segment/segment + point/segment closest point
//   * SimToolbox/FDPS/particle_simulator.hpp -- vendored FDPS 6.0b2 (tree, periodic images,
//                                               TreeForForceShort<>::Symmetry neighbour walk)
// The pair functor (SimToolbox/Sylinder/SylinderNear.hpp:197-414) cannot be compiled here
// (it drags in mpi.h, Eigen and Tpetra), so the functor body below re-derives the block
// fields with a 3-vector type of our own while calling the reference's DCPQuery verbatim.
//
// Output goes to oracle/_ref/libalens_ref.so (git-ignored, travels to the GPU box).
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    V3 operator+(const V3 &o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3 &o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator*(double s) const { return V3(x * s, y * s, z * s); }
    double dot(const V3 &o) const { return x * o.x + y * o.y + z * o.z; }
    double norm() const { return sqrt(x * x + y * y + z * z); }
    void normalize() {
        double n = norm(); // Eigen >= 3.3 normalized(): unchanged when the squared norm is not > 0
        if (n > 0) { x /= n; y /= n; z /= n; }
    }
};
static inline V3 operator*(double s, const V3 &v) { return V3(v.x * s, v.y * s, v.z * s); }

#include "Collision/DCPQuery.hpp"
#include "FDPS/particle_simulator.hpp"

extern "C" {

// plain mirrors of the oracle's C structs (oracle/alens_oracle.h); layouts must match
struct ref_rod {
    int gid, globalIndex, rank, pad_;
    double radius, length, radiusCollision, lengthCollision, colBuf;
    double pos[3];
    double direction[3];
};

struct ref_pair {
    int gidI, gidJ;
    double delta0;
    double normI[3];
    double posI[3], posJ[3];
    double labI[3], labJ[3];
};

double ref_dcp_segseg(const double *P0, const double *P1, const double *Q0, const double *Q1, double *Ploc,
                      double *Qloc, double *s, double *t) {
    DCPQuery<3, double, V3> q;
    V3 p, qq;
    double d = q(V3(P0[0], P0[1], P0[2]), V3(P1[0], P1[1], P1[2]), V3(Q0[0], Q0[1], Q0[2]), V3(Q1[0], Q1[1], Q1[2]), p,
                 qq, *s, *t);
    Ploc[0] = p.x; Ploc[1] = p.y; Ploc[2] = p.z;
    Qloc[0] = qq.x; Qloc[1] = qq.y; Qloc[2] = qq.z;
    return d;
}

double ref_dist_point_seg(const double *pt, const double *m, const double *p, double *perp) {
    V3 out;
    double d = DistPointSeg<V3>(V3(pt[0], pt[1], pt[2]), V3(m[0], m[1], m[2]), V3(p[0], p[1], p[2]), out);
    perp[0] = out.x; perp[1] = out.y; perp[2] = out.z;
    return d;
}
}

// ---------------------------------------------------------------- FDPS types
namespace {

struct RodFP { // "full particle": FDPS only needs getPos/setPos
    ref_rod r;
    PS::F64vec getPos() const { return PS::F64vec(r.pos[0], r.pos[1], r.pos[2]); }
    void setPos(const PS::F64vec &p) { r.pos[0] = p.x; r.pos[1] = p.y; r.pos[2] = p.z; }
};

struct RodEP { // essential particle, same search radius rule as SylinderNear.hpp:108-113
    ref_rod r;
    PS::F64vec getPos() const { return PS::F64vec(r.pos[0], r.pos[1], r.pos[2]); }
    void setPos(const PS::F64vec &p) { r.pos[0] = p.x; r.pos[1] = p.y; r.pos[2] = p.z; }
    void copyFromFP(const RodFP &fp) { r = fp.r; }
    PS::F64 getRSearch() const {
        const double b = .5 * std::max(r.length + 2. * r.radius, r.lengthCollision + 2. * r.radiusCollision);
        return b + r.colBuf;
    }
};

struct ForceDummy {
    double f[3];
    void clear() { f[0] = f[1] = f[2] = 0; }
};

inline bool isSphere(const ref_rod &s) { return s.lengthCollision < 2 * s.radiusCollision; }

// one candidate pair -> block fields; arithmetic order follows SylinderNear.hpp:253-414
inline bool pairBlock(const ref_rod &a, const ref_rod &b, ref_pair &out) {
    const bool sa = isSphere(a), sb = isSphere(b);
    V3 cI(a.pos[0], a.pos[1], a.pos[2]), cJ(b.pos[0], b.pos[1], b.pos[2]);
    V3 Ploc, Qloc;
    double sep;
    bool reverse = false;
    if (sa && sb) {
        const double radI = a.lengthCollision * 0.5 + a.radiusCollision;
        const double radJ = b.lengthCollision * 0.5 + b.radiusCollision;
        const V3 rIJ = cJ - cI;
        sep = rIJ.norm() - (radI + radJ);
        Ploc = cI; Qloc = cJ;
    } else if (sa || sb) {
        // sp_sy(sphere, sylinder); when I is the sylinder the reference calls sp_sy(J,I,...,reverse)
        const ref_rod &sp = sa ? a : b;
        const ref_rod &sy = sa ? b : a;
        reverse = !sa;
        const V3 cs(sp.pos[0], sp.pos[1], sp.pos[2]);
        const V3 cy(sy.pos[0], sy.pos[1], sy.pos[2]);
        const V3 dy(sy.direction[0], sy.direction[1], sy.direction[2]);
        const double radI = sp.lengthCollision * 0.5 + sp.radiusCollision;
        const V3 Qm = cy - dy * (0.5 * sy.lengthCollision);
        const V3 Qp = cy + dy * (0.5 * sy.lengthCollision);
        V3 q;
        const double dist = DistPointSeg<V3>(cs, Qm, Qp, q);
        sep = dist - (radI + sy.radiusCollision);
        Ploc = cs; Qloc = q; // (sphere side, sylinder side)
        cI = cs; cJ = cy;
    } else {
        DCPQuery<3, double, V3> dcp;
        const V3 dI(a.direction[0], a.direction[1], a.direction[2]);
        const V3 dJ(b.direction[0], b.direction[1], b.direction[2]);
        const V3 Pm = cI - dI * (0.5 * a.lengthCollision);
        const V3 Pp = cI + dI * (0.5 * a.lengthCollision);
        const V3 Qm = cJ - dJ * (0.5 * b.lengthCollision);
        const V3 Qp = cJ + dJ * (0.5 * b.lengthCollision);
        double s, t = 0;
        const double dist = dcp(Pm, Pp, Qm, Qp, Ploc, Qloc, s, t);
        sep = dist - (a.radiusCollision + b.radiusCollision);
    }
    const double buffer = std::max(a.colBuf, b.colBuf);
    if (!(sep < buffer))
        return false;
    V3 nI = Ploc - Qloc;
    nI.normalize();
    V3 pI = Ploc - cI, pJ = Qloc - cJ;
    out.delta0 = sep;
    if (!reverse) {
        out.gidI = a.gid; out.gidJ = b.gid;
        out.normI[0] = nI.x; out.normI[1] = nI.y; out.normI[2] = nI.z;
        out.posI[0] = pI.x; out.posI[1] = pI.y; out.posI[2] = pI.z;
        out.posJ[0] = pJ.x; out.posJ[1] = pJ.y; out.posJ[2] = pJ.z;
        out.labI[0] = Ploc.x; out.labI[1] = Ploc.y; out.labI[2] = Ploc.z;
        out.labJ[0] = Qloc.x; out.labJ[1] = Qloc.y; out.labJ[2] = Qloc.z;
    } else { // ConstraintBlock::reverseIJ(): swap I<->J, normJ = -normI becomes normI
        out.gidI = a.gid; out.gidJ = b.gid;
        out.normI[0] = -nI.x; out.normI[1] = -nI.y; out.normI[2] = -nI.z;
        out.posI[0] = pJ.x; out.posI[1] = pJ.y; out.posI[2] = pJ.z;
        out.posJ[0] = pI.x; out.posJ[1] = pI.y; out.posJ[2] = pI.z;
        out.labI[0] = Qloc.x; out.labI[1] = Qloc.y; out.labI[2] = Qloc.z;
        out.labJ[0] = Ploc.x; out.labJ[1] = Ploc.y; out.labJ[2] = Ploc.z;
    }
    return true;
}

struct Functor {
    std::vector<std::vector<ref_pair>> *pool;
    long long *ncand; // per-thread candidate counters (gidI<gidJ tested pairs)
    void operator()(const RodEP *const ep_i, const PS::S32 Nip, const RodEP *const ep_j, const PS::S32 Njp,
                    ForceDummy *const force) {
        const int tid = omp_get_thread_num();
        auto &que = (*pool)[tid];
        long long c = 0;
        for (PS::S32 i = 0; i < Nip; ++i) {
            force[i].clear();
            for (PS::S32 j = 0; j < Njp; ++j) {
                if (ep_i[i].r.gid >= ep_j[j].r.gid)
                    continue;
                c++;
                ref_pair blk;
                if (pairBlock(ep_i[i].r, ep_j[j].r, blk))
                    que.push_back(blk);
            }
        }
        ncand[tid * 8] += c;
    }
};

using Tree = PS::TreeForForceShort<ForceDummy, RodEP, RodEP>::Symmetry;

struct RefState {
    bool psInit = false;
    PS::DomainInfo *dinfo = nullptr;
    PS::ParticleSystem<RodFP> *psys = nullptr;
    Tree *tree = nullptr;
    int treeN = 0;
    std::vector<std::vector<ref_pair>> pool;
    std::vector<ref_pair> flat;
    double tLast = 0;
    long long candLast = 0;
};
RefState G;

} // namespace

extern "C" {

// one candidate pair through the functor restatement that uses the reference DCPQuery
int ref_pair_block(const ref_rod *a, const ref_rod *b, ref_pair *out) { return pairBlock(*a, *b, *out) ? 1 : 0; }

// Runs the reference neighbour search exactly as SylinderSystem does on one rank:
// setDomainInfo (SylinderSystem.cpp:569-610) -> adjustPositionIntoRootDomain + decomposeDomainAll
// (:612-615) -> exchangeParticle (:617) -> TreeSylinderNear::calcForceAll (:1152-1160).
// rods[] positions are updated in place by the box wrap (as applyBoxBC does).
// Returns number of blocks found; fetch them with ref_fdps_get().
long long ref_fdps_collect(int n, ref_rod *rods, const double *boxLow, const double *boxHigh, const int *pbc,
                           int nthreads, int rebuild) {
    if (!G.psInit) {
        // FDPS sizes its per-thread buffers once, from omp_get_max_threads() at Initialize: do that with every
        // core so that later calls may ask for any smaller thread count
        int argc = 0;
        char **argv = nullptr;
        omp_set_num_threads(omp_get_num_procs());
        PS::Initialize(argc, argv);
        G.psInit = true;
    }
    if (nthreads > 0)
        omp_set_num_threads(std::min(nthreads, omp_get_num_procs()));
    if (rebuild || !G.dinfo) {
        delete G.tree; G.tree = nullptr; G.treeN = 0;
        delete G.psys; delete G.dinfo;
        G.dinfo = new PS::DomainInfo();
        G.dinfo->initialize();
        const int flag = 100 * (pbc[0] ? 1 : 0) + 10 * (pbc[1] ? 1 : 0) + (pbc[2] ? 1 : 0);
        switch (flag) {
        case 0: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_OPEN); break;
        case 1: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_Z); break;
        case 10: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_Y); break;
        case 100: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_X); break;
        case 11: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_YZ); break;
        case 101: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_XZ); break;
        case 110: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_XY); break;
        case 111: G.dinfo->setBoundaryCondition(PS::BOUNDARY_CONDITION_PERIODIC_XYZ); break;
        }
        G.dinfo->setPosRootDomain(PS::F64vec(boxLow[0], boxLow[1], boxLow[2]),
                                  PS::F64vec(boxHigh[0], boxHigh[1], boxHigh[2]));
        G.psys = new PS::ParticleSystem<RodFP>();
        G.psys->initialize();
        G.psys->setAverageTargetNumberOfSampleParticlePerProcess(200);
    }
    G.psys->setNumberOfParticleLocal(n);
    for (int i = 0; i < n; i++)
        (*G.psys)[i].r = rods[i];
    G.psys->adjustPositionIntoRootDomain(*G.dinfo);
    if (rebuild || G.treeN == 0)
        G.dinfo->decomposeDomainAll(*G.psys);
    G.psys->exchangeParticle(*G.dinfo);
    for (int i = 0; i < n; i++)
        rods[i] = (*G.psys)[i].r; // single rank: order is preserved
    if (n > 1.5 * G.treeN || !G.tree) {
        delete G.tree;
        G.tree = new Tree();
        G.tree->initialize(2 * (n > 0 ? n : 1));
        G.treeN = n;
    }
    const int nt = omp_get_max_threads();
    G.pool.assign(nt, {});
    std::vector<long long> ncand(nt * 8, 0);
    Functor f;
    f.pool = &G.pool;
    f.ncand = ncand.data();
    const double t0 = omp_get_wtime();
    G.tree->calcForceAll(f, *G.psys, *G.dinfo);
    G.tLast = omp_get_wtime() - t0;
    G.candLast = 0;
    for (int t = 0; t < nt; t++) G.candLast += ncand[t * 8];
    G.flat.clear();
    for (auto &q : G.pool)
        G.flat.insert(G.flat.end(), q.begin(), q.end());
    return (long long)G.flat.size();
}

void ref_fdps_get(ref_pair *out) {
    if (!G.flat.empty())
        memcpy(out, G.flat.data(), G.flat.size() * sizeof(ref_pair));
}
double ref_fdps_last_seconds() { return G.tLast; }
long long ref_fdps_last_candidates() { return G.candLast; }
int ref_sizeof_rod() { return (int)sizeof(ref_rod); }
int ref_sizeof_pair() { return (int)sizeof(ref_pair); }
}
