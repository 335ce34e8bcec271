/*
 * oracle/alens_oracle.c -- CPU restatement of the aLENS/SimToolbox collision-constraint hot path.
 * TEST INFRASTRUCTURE ONLY -- see alens_oracle.h for the rules and the parity status.
 *
 * Build: gcc -O2 -std=c11 -fopenmp -ffp-contract=off -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off so that every expression rounds exactly as written (the GPU pair kernel is
 * compiled with -fmad=false for the same reason): the integer pair list must not depend on FMAs.
 *
 * All citations are relative to /root/reference/.
 */
#include "alens_oracle.h"

#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int orc_sizeof_rod(void) { return (int)sizeof(orc_rod); }
int orc_sizeof_block(void) { return (int)sizeof(orc_block); }

/* ------------------------------------------------------------------ small 3-vector helpers */
static inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void sub3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2];
}
static inline double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------ DCPQuery.hpp:310-340 */
static double clamped_root(double slope, double h0, double h1) {
    const double eps = DBL_EPSILON;
    double r;
    if (fabs(h0) < eps && fabs(h1) < eps) {
        r = 0.5;
    } else if (h0 < 0) {
        if (h1 > 0) {
            r = -h0 / slope;
            r = r > 0.0 ? r : 0.0; /* std::max(-h0/slope, 0) */
            r = r < 1.0 ? r : 1.0; /* std::min(.,1) */
        } else {
            r = 1;
        }
    } else {
        r = 0;
    }
    return r;
}

typedef struct {
    double mA, mB, mC, mD, mE;
    double mF00, mF10, mF01, mF11;
    double mG00, mG10, mG01, mG11;
} dcp_state;

/* DCPQuery.hpp:343-424 */
static void dcp_intersection(const dcp_state *q, const double sValue[2], const int classify[2], int edge[2],
                             double end[2][2]) {
    if (classify[0] < 0) {
        edge[0] = 0;
        end[0][0] = 0;
        end[0][1] = q->mF00 / q->mB;
        if (end[0][1] < 0 || end[0][1] > 1) end[0][1] = 0.5;
        if (classify[1] == 0) {
            edge[1] = 3;
            end[1][0] = sValue[1];
            end[1][1] = 1;
        } else {
            edge[1] = 1;
            end[1][0] = 1;
            end[1][1] = q->mF10 / q->mB;
            if (end[1][1] < 0 || end[1][1] > 1) end[1][1] = 0.5;
        }
    } else if (classify[0] == 0) {
        edge[0] = 2;
        end[0][0] = sValue[0];
        end[0][1] = 0;
        if (classify[1] < 0) {
            edge[1] = 0;
            end[1][0] = 0;
            end[1][1] = q->mF00 / q->mB;
            if (end[1][1] < 0 || end[1][1] > 1) end[1][1] = 0.5;
        } else if (classify[1] == 0) {
            edge[1] = 3;
            end[1][0] = sValue[1];
            end[1][1] = 1;
        } else {
            edge[1] = 1;
            end[1][0] = 1;
            end[1][1] = q->mF10 / q->mB;
            if (end[1][1] < 0 || end[1][1] > 1) end[1][1] = 0.5;
        }
    } else {
        edge[0] = 1;
        end[0][0] = 1;
        end[0][1] = q->mF10 / q->mB;
        if (end[0][1] < 0 || end[0][1] > 1) end[0][1] = 0.5;
        if (classify[1] == 0) {
            edge[1] = 3;
            end[1][0] = sValue[1];
            end[1][1] = 1;
        } else {
            edge[1] = 0;
            end[1][0] = 0;
            end[1][1] = q->mF00 / q->mB;
            if (end[1][1] < 0 || end[1][1] > 1) end[1][1] = 0.5;
        }
    }
}

/* DCPQuery.hpp:427-472 */
static void dcp_min_params(const dcp_state *q, const int edge[2], const double end[2][2], double parameter[2]) {
    const double eps = DBL_EPSILON;
    const double delta = end[1][1] - end[0][1];
    const double h0 = delta * ((-q->mB * end[0][0] - q->mE) + q->mC * end[0][1]);
    const double h1 = delta * ((-q->mB * end[1][0] - q->mE) + q->mC * end[1][1]);
    if (fabs(h0) < fabs(q->mC) * eps && fabs(h1) < fabs(q->mC) * eps) {
        const double z = 0.5, omz = 1.0 - z;
        parameter[0] = omz * end[0][0] + z * end[1][0];
        parameter[1] = omz * end[0][1] + z * end[1][1];
    } else if (h0 >= 0) {
        if (edge[0] == 0) {
            parameter[0] = 0;
            parameter[1] = clamped_root(q->mC, q->mG00, q->mG01);
        } else if (edge[0] == 1) {
            parameter[0] = 1;
            parameter[1] = clamped_root(q->mC, q->mG10, q->mG11);
        } else {
            parameter[0] = end[0][0];
            parameter[1] = end[0][1];
        }
    } else {
        if (h1 <= 0) {
            if (edge[1] == 0) {
                parameter[0] = 0;
                parameter[1] = clamped_root(q->mC, q->mG00, q->mG01);
            } else if (edge[1] == 1) {
                parameter[0] = 1;
                parameter[1] = clamped_root(q->mC, q->mG10, q->mG11);
            } else {
                parameter[0] = end[1][0];
                parameter[1] = end[1][1];
            }
        } else {
            const double z = clamped_root(h1 - h0, h0, h1);
            const double omz = 1.0 - z;
            parameter[0] = omz * end[0][0] + z * end[1][0];
            parameter[1] = omz * end[0][1] + z * end[1][1];
        }
    }
}

/* DCPQuery.hpp:199-308 */
double orc_dcp_segseg(const double P0[3], const double P1[3], const double Q0[3], const double Q1[3], double Ploc[3],
                      double Qloc[3], double *s, double *t) {
    dcp_state q;
    double P1mP0[3], Q1mQ0[3], P0mQ0[3], par[2];
    sub3(P1, P0, P1mP0);
    sub3(Q1, Q0, Q1mQ0);
    sub3(P0, Q0, P0mQ0);
    q.mA = dot3(P1mP0, P1mP0);
    q.mB = dot3(P1mP0, Q1mQ0);
    q.mC = dot3(Q1mQ0, Q1mQ0);
    q.mD = dot3(P1mP0, P0mQ0);
    q.mE = dot3(Q1mQ0, P0mQ0);
    q.mF00 = q.mD;
    q.mF10 = q.mF00 + q.mA;
    q.mF01 = q.mF00 - q.mB;
    q.mF11 = q.mF10 - q.mB;
    q.mG00 = -q.mE;
    q.mG10 = q.mG00 - q.mB;
    q.mG01 = q.mG00 + q.mC;
    q.mG11 = q.mG10 + q.mC;

    if (q.mA > 0 && q.mC > 0) {
        double sValue[2];
        int classify[2];
        sValue[0] = clamped_root(q.mA, q.mF00, q.mF10);
        sValue[1] = clamped_root(q.mA, q.mF01, q.mF11);
        for (int i = 0; i < 2; ++i) {
            if (sValue[i] <= 0) classify[i] = -1;
            else if (sValue[i] >= 1) classify[i] = +1;
            else classify[i] = 0;
        }
        if (classify[0] == -1 && classify[1] == -1) {
            par[0] = 0;
            par[1] = clamped_root(q.mC, q.mG00, q.mG01);
        } else if (classify[0] == +1 && classify[1] == +1) {
            par[0] = 1;
            par[1] = clamped_root(q.mC, q.mG10, q.mG11);
        } else {
            int edge[2];
            double end[2][2];
            dcp_intersection(&q, sValue, classify, edge, end);
            dcp_min_params(&q, edge, (const double(*)[2])end, par);
        }
    } else {
        if (q.mA > 0) {
            par[0] = clamped_root(q.mA, q.mF00, q.mF10);
            par[1] = 0;
        } else if (q.mC > 0) {
            par[0] = 0;
            par[1] = clamped_root(q.mC, q.mG00, q.mG01);
        } else {
            par[0] = 0;
            par[1] = 0;
        }
    }
    double diff[3];
    for (int k = 0; k < 3; k++) {
        Ploc[k] = (1.0 - par[0]) * P0[k] + par[0] * P1[k];
        Qloc[k] = (1.0 - par[1]) * Q0[k] + par[1] * Q1[k];
    }
    sub3(Ploc, Qloc, diff);
    *s = par[0];
    *t = par[1];
    return sqrt(dot3(diff, diff));
}

/* DCPQuery.hpp:91-128 */
double orc_dist_point_seg(const double pt[3], const double minus[3], const double plus[3], double perp[3]) {
    double direction[3], diff[3], closest[3];
    sub3(plus, minus, direction);
    sub3(pt, plus, diff);
    double t = dot3(direction, diff);
    if (t >= 0) {
        memcpy(closest, plus, sizeof(closest));
    } else {
        sub3(pt, minus, diff);
        t = dot3(direction, diff);
        if (t <= 0) {
            memcpy(closest, minus, sizeof(closest));
        } else {
            const double sqrLength = dot3(direction, direction);
            if (sqrLength > 0) {
                t /= sqrLength;
                for (int k = 0; k < 3; k++) closest[k] = minus[k] + t * direction[k];
            } else {
                memcpy(closest, minus, sizeof(closest));
            }
        }
    }
    sub3(pt, closest, diff);
    memcpy(perp, closest, sizeof(closest));
    return sqrt(dot3(diff, diff));
}

/* ------------------------------------------------------------------ rod preparation */
/* Eigen Quaternion * Vector3 with v = (0,0,1) (SylinderNear.hpp:86; Eigen 3.3 _transformVector:
 * uv = q.vec x v; uv += uv; v + q.w*uv + q.vec x uv), memory order of coeffs is (x,y,z,w) */
void orc_quat_to_dir(const double q[4], double dir[3]) {
    const double v[3] = {0, 0, 1};
    double uv[3], c[3];
    cross3(q, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(q, uv, c);
    for (int k = 0; k < 3; k++) dir[k] = (v[k] + q[3] * uv[k]) + c[k];
}

/* SylinderSystem.cpp:897-905 (collision geometry, colBuf) + SylinderNear.hpp:74-90 (copyFromFP)
 * + SylinderSystem.cpp:868-880 (globalIndex = base + i) */
void orc_make_rods(int n, const int *gid, const double *radius, const double *length, const double *pos,
                   const double *quat, double dRatio, double lRatio, double colBuf, int base, orc_rod *out) {
    for (int i = 0; i < n; i++) {
        orc_rod r;
        memset(&r, 0, sizeof(r));
        r.gid = gid[i];
        r.globalIndex = base + i;
        r.rank = 0;
        r.radius = radius[i];
        r.length = length[i];
        r.radiusCollision = radius[i] * dRatio;
        r.lengthCollision = length[i] * lRatio;
        r.colBuf = colBuf;
        memcpy(r.pos, pos + 3 * i, 3 * sizeof(double));
        orc_quat_to_dir(quat + 4 * i, r.direction);
        out[i] = r;
    }
}

/* FDPS/particle_system.hpp:798-843 adjustPositionIntoRootDomain wraps into the ROOT domain, and
 * setPosRootDomain (FDPS/domain_info.hpp:1191-1212) only narrows the periodic axes: an open axis keeps
 * +-LARGE_FLOAT, so a rod never moves along it.  pbc == NULL: all three axes periodic. */
void orc_wrap_positions(int n, double *pos, const double lo[3], const double hi[3], const int *pbc) {
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            if (pbc && !pbc[k]) continue;
            double x = pos[3 * i + k];
            const double len = hi[k] - lo[k];
            while (x < lo[k]) x += len;
            while (x >= hi[k]) x -= len;
            if (x == hi[k]) x = lo[k];
            pos[3 * i + k] = x;
        }
}

/* ------------------------------------------------------------------ collideStress, SylinderNear.hpp:432-519 */
static void init_syN(double N[3][3], double r, double h, double rho) {
    memset(N, 0, 9 * sizeof(double));
    const double beta = h / 2.0 / r;
    N[0][0] = 1.0 / 30.0 * (15.0 * beta + 8);
    N[1][1] = N[0][0];
    N[2][2] = 1.0 / 15.0 * (10.0 * beta * beta * beta + 20.0 * beta * beta + 15.0 * beta + 4.0);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) N[i][j] = N[i][j] * rho * r * r * r * r * r * M_PI;
}
static void init_syGA(double G[3][3], double r, double h, double rho) {
    memset(G, 0, 9 * sizeof(double));
    const double beta = h / 2.0 / r;
    G[0][0] = 1.0 / 30.0 * (20.0 * beta * beta * beta + 40.0 * beta * beta + 45.0 * beta + 16.0);
    G[1][1] = G[0][0];
    G[2][2] = 1.0 / 15.0 * (15 * beta + 8);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) G[i][j] = G[i][j] * M_PI * r * r * r * r * r * rho;
}
static void iso_plus_dyad(double a, double b, const double d[3], double out[3][3]) {
    /* a*I + (b-a)*(d d^T) */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i][j] = a * (i == j ? 1.0 : 0.0) + (b - a) * (d[i] * d[j]);
}

void orc_collide_stress(const double dirI[3], const double dirJ[3], const double centerI[3], const double centerJ[3],
                        double hI, double hJ, double rI, double rJ, double rho, const double Ploc[3],
                        const double Qloc[3], double stress[9]) {
    static const double epsilon[3][3][3] = {{{0, 0, 0}, {0, 0, 1}, {0, -1, 0}},
                                            {{0, 0, -1}, {0, 0, 0}, {1, 0, 0}},
                                            {{0, 1, 0}, {-1, 0, 0}, {0, 0, 0}}};
    double NI[3][3], GI[3][3], NJ[3][3], GJ[3][3], iGI[3][3], iGJ[3][3];
    init_syN(NI, rI, hI, rho);
    init_syGA(GI, rI, hI, rho);
    init_syN(NJ, rJ, hJ, rho);
    init_syGA(GJ, rJ, hJ, rho);
    double aI = NI[0][0], bI = NI[2][2], aJ = NJ[0][0], bJ = NJ[2][2];
    iso_plus_dyad(aI, bI, dirI, NI);
    iso_plus_dyad(aJ, bJ, dirJ, NJ);
    aI = 1.0 / GI[0][0]; bI = 1.0 / GI[2][2];
    aJ = 1.0 / GJ[0][0]; bJ = 1.0 / GJ[2][2];
    iso_plus_dyad(aI, bI, dirI, iGI);
    iso_plus_dyad(aJ, bJ, dirJ, iGJ);

    double F1[3], mF1[3], tmp[3], xICf[3], xJCf[3];
    sub3(Qloc, Ploc, F1);
    {
        const double n2 = dot3(F1, F1);
        if (n2 > 0) {
            const double n = sqrt(n2);
            F1[0] /= n; F1[1] /= n; F1[2] /= n;
        }
    }
    mF1[0] = -F1[0]; mF1[1] = -F1[1]; mF1[2] = -F1[2];
    sub3(Ploc, centerI, tmp);
    cross3(tmp, mF1, xICf);
    sub3(Qloc, centerJ, tmp);
    cross3(tmp, F1, xJCf);

    double SGI[3][3], SGJ[3][3];
    memset(SGI, 0, sizeof(SGI));
    memset(SGJ, 0, sizeof(SGJ));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++)
                    for (int r = 0; r < 3; r++) {
                        SGI[i][j] = SGI[i][j] + NI[i][l] * epsilon[j][k][l] * iGI[k][r] * xICf[r];
                        SGJ[i][j] = SGJ[i][j] + NJ[i][l] * epsilon[j][k][l] * iGJ[k][r] * xJCf[r];
                    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double rIf = centerI[i] * mF1[j];
            const double rJf = centerJ[i] * F1[j];
            stress[3 * i + j] = ((rIf + rJf) + SGI[i][j]) + SGJ[i][j];
        }
}

/* ------------------------------------------------------------------ pair functor, SylinderNear.hpp:197-414 */
static inline int is_sphere_col(const orc_rod *s) { return s->lengthCollision < 2 * s->radiusCollision; }

static void fill_block(orc_block *b, double delta0, double gamma, const orc_rod *I, const orc_rod *J,
                       const double normI[3], const double posI[3], const double posJ[3], const double labI[3],
                       const double labJ[3]) {
    memset(b, 0, sizeof(*b));
    b->delta0 = delta0;
    b->gamma = gamma;
    b->gammaLB = 0;
    b->gidI = I->gid;
    b->gidJ = J->gid;
    b->globalIndexI = I->globalIndex;
    b->globalIndexJ = J->globalIndex;
    b->oneSide = 0;
    b->bilateral = 0;
    b->kappa = 0;
    for (int k = 0; k < 3; k++) {
        b->normI[k] = normI[k];
        b->normJ[k] = -normI[k];
        b->posI[k] = posI[k];
        b->posJ[k] = posJ[k];
        b->labI[k] = labI[k];
        b->labJ[k] = labJ[k];
    }
}

static void reverse_ij(orc_block *b) { /* ConstraintBlock.hpp:111-119 */
    int t;
    t = b->gidI; b->gidI = b->gidJ; b->gidJ = t;
    t = b->globalIndexI; b->globalIndexI = b->globalIndexJ; b->globalIndexJ = t;
    for (int k = 0; k < 3; k++) {
        double d;
        d = b->normI[k]; b->normI[k] = b->normJ[k]; b->normJ[k] = d;
        d = b->posI[k]; b->posI[k] = b->posJ[k]; b->posJ[k] = d;
        d = b->labI[k]; b->labI[k] = b->labJ[k]; b->labJ[k] = d;
    }
}

/* a = target (ep_i), b = source (ep_j); caller has already applied the gid filter if wanted.
 * Dispatch exactly as operator() :207-236 */
int orc_pair_functor(const orc_rod *a, const orc_rod *b, int withStress, orc_block *out) {
    static const double ez[3] = {0, 0, 1};
    const int sa = is_sphere_col(a), sb = is_sphere_col(b);
    double Ploc[3], Qloc[3], normI[3], posI[3], posJ[3], d[3];
    const double buffer = a->colBuf > b->colBuf ? a->colBuf : b->colBuf; /* symmetric */
    if (sa && sb) { /* sp_sp :253-294 */
        const double radI = a->lengthCollision * 0.5 + a->radiusCollision;
        const double radJ = b->lengthCollision * 0.5 + b->radiusCollision;
        double rIJ[3];
        sub3(b->pos, a->pos, rIJ);
        const double sep = norm3(rIJ) - (radI + radJ);
        if (!(sep < buffer)) return 0;
        memcpy(Ploc, a->pos, sizeof(Ploc));
        memcpy(Qloc, b->pos, sizeof(Qloc));
        sub3(Ploc, Qloc, d);
        const double n = norm3(d); /* Eigen >= 3.3 normalized(): unchanged when the squared norm is not > 0 */
        for (int k = 0; k < 3; k++) normI[k] = n > 0 ? d[k] / n : d[k];
        sub3(Ploc, a->pos, posI);
        sub3(Qloc, b->pos, posJ);
        fill_block(out, sep, sep < 0 ? -sep : 0, a, b, normI, posI, posJ, Ploc, Qloc);
        if (withStress) orc_collide_stress(ez, ez, a->pos, b->pos, 0, 0, radI, radJ, 1.0, Ploc, Qloc, out->stress);
        return 1;
    }
    if (sa || sb) { /* sp_sy :307-355; reverse when the target is the sylinder (:230) */
        const orc_rod *sp = sa ? a : b;
        const orc_rod *sy = sa ? b : a;
        const double radI = sp->lengthCollision * 0.5 + sp->radiusCollision;
        double Qm[3], Qp[3];
        for (int k = 0; k < 3; k++) {
            Qm[k] = sy->pos[k] - sy->direction[k] * (0.5 * sy->lengthCollision);
            Qp[k] = sy->pos[k] + sy->direction[k] * (0.5 * sy->lengthCollision);
        }
        memcpy(Ploc, sp->pos, sizeof(Ploc));
        const double distMin = orc_dist_point_seg(sp->pos, Qm, Qp, Qloc);
        const double sep = distMin - (radI + sy->radiusCollision);
        if (!(sep < buffer)) return 0;
        sub3(Ploc, Qloc, d);
        const double n = norm3(d); /* Eigen >= 3.3 normalized(): unchanged when the squared norm is not > 0 */
        for (int k = 0; k < 3; k++) normI[k] = n > 0 ? d[k] / n : d[k];
        sub3(Ploc, sp->pos, posI);
        sub3(Qloc, sy->pos, posJ);
        fill_block(out, sep, sep < 0 ? -sep : 0, sp, sy, normI, posI, posJ, Ploc, Qloc);
        if (!sa) reverse_ij(out);
        if (withStress)
            orc_collide_stress(ez, sy->direction, sp->pos, sy->pos, 0, sy->lengthCollision, radI, sy->radiusCollision,
                               1.0, Ploc, Qloc, out->stress);
        return 1;
    }
    /* sy_sy :367-414 */
    double Pm[3], Pp[3], Qm[3], Qp[3], s, t;
    for (int k = 0; k < 3; k++) {
        Pm[k] = a->pos[k] - a->direction[k] * (0.5 * a->lengthCollision);
        Pp[k] = a->pos[k] + a->direction[k] * (0.5 * a->lengthCollision);
        Qm[k] = b->pos[k] - b->direction[k] * (0.5 * b->lengthCollision);
        Qp[k] = b->pos[k] + b->direction[k] * (0.5 * b->lengthCollision);
    }
    const double distMin = orc_dcp_segseg(Pm, Pp, Qm, Qp, Ploc, Qloc, &s, &t);
    const double sep = distMin - (a->radiusCollision + b->radiusCollision);
    if (!(sep < buffer)) return 0;
    sub3(Ploc, Qloc, d);
    const double n = norm3(d); /* Eigen >= 3.3 normalized(): unchanged when the squared norm is not > 0 */
    for (int k = 0; k < 3; k++) normI[k] = n > 0 ? d[k] / n : d[k];
    sub3(Ploc, a->pos, posI);
    sub3(Qloc, b->pos, posJ);
    fill_block(out, sep, sep < 0 ? -sep : 0, a, b, normI, posI, posJ, Ploc, Qloc);
    if (withStress)
        orc_collide_stress(a->direction, b->direction, a->pos, b->pos, a->lengthCollision, b->lengthCollision,
                           a->radiusCollision, b->radiusCollision, 1.0, Ploc, Qloc, out->stress);
    return 1;
}

/* ------------------------------------------------------------------ geometric pair list P_geo */
static int cmp_block(const void *pa, const void *pb) {
    const orc_block *a = (const orc_block *)pa, *b = (const orc_block *)pb;
    if (a->gidI != b->gidI) return a->gidI < b->gidI ? -1 : 1;
    if (a->gidJ != b->gidJ) return a->gidJ < b->gidJ ? -1 : 1;
    /* same pair through two different periodic images (tiny boxes only): order by labJ */
    for (int k = 0; k < 3; k++)
        if (a->labJ[k] != b->labJ[k]) return a->labJ[k] < b->labJ[k] ? -1 : 1;
    return 0;
}

static inline double bound_radius(const orc_rod *r) { return 0.5 * r->lengthCollision + r->radiusCollision; }

/* I = lower gid at its own position, J = higher gid shifted by k*boxLen (FDPS image rule:
 * FDPS/tree_for_force_utils.hpp:256-262 pos_new = pos + shift, shift = ix*size_root_domain) */
static int try_pair(const orc_rod *I, const orc_rod *J, const double shift[3], int withStress, orc_block *blk) {
    orc_rod Js = *J;
    for (int k = 0; k < 3; k++) Js.pos[k] = J->pos[k] + shift[k];
    /* conservative centre-distance prefilter (never rejects a true contact) */
    double d[3];
    sub3(Js.pos, I->pos, d);
    const double cut = bound_radius(I) + bound_radius(J) + (I->colBuf > J->colBuf ? I->colBuf : J->colBuf);
    if (dot3(d, d) > cut * cut * (1.0 + 1e-10)) return 0;
    return orc_pair_functor(I, &Js, withStress, blk);
}

long long orc_collect_pairs_brute(int n, const orc_rod *rods, const double lo[3], const double hi[3], const int pbc[3],
                                  int withStress, orc_block *out, long long cap) {
    long long cnt = 0;
    const double len[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const int kx = pbc[0] ? 1 : 0, ky = pbc[1] ? 1 : 0, kz = pbc[2] ? 1 : 0;
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) {
            if (rods[a].gid >= rods[b].gid) continue;
            for (int ix = -kx; ix <= kx; ix++)
                for (int iy = -ky; iy <= ky; iy++)
                    for (int iz = -kz; iz <= kz; iz++) {
                        const double shift[3] = {ix * len[0], iy * len[1], iz * len[2]};
                        orc_block blk;
                        if (try_pair(&rods[a], &rods[b], shift, withStress, &blk)) {
                            if (cnt < cap) out[cnt] = blk;
                            cnt++;
                        }
                    }
        }
    if (cnt <= cap) qsort(out, (size_t)cnt, sizeof(orc_block), cmp_block);
    return cnt;
}

/* uniform cell list, OpenMP over cells; used as the CPU baseline at sizes brute force cannot reach */
long long orc_collect_pairs_cells(int n, const orc_rod *rods, const double lo[3], const double hi[3], const int pbc[3],
                                  int withStress, orc_block *out, long long cap, int nthreads) {
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double maxR = 0, maxBuf = 0;
    for (int i = 0; i < n; i++) {
        const double R = bound_radius(&rods[i]);
        if (R > maxR) maxR = R;
        if (rods[i].colBuf > maxBuf) maxBuf = rods[i].colBuf;
    }
    const double cutoff = (2 * maxR + maxBuf) * (1.0 + 1e-9);
    const double len[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    int nc[3];
    for (int k = 0; k < 3; k++) {
        double m = floor(len[k] / cutoff);
        if (m < 1) m = 1;
        if (m > 512) m = 512;
        nc[k] = (int)m;
    }
    const long long ncell = (long long)nc[0] * nc[1] * nc[2];
    int *cellOf = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *start = (int *)calloc((size_t)ncell + 1, sizeof(int));
    int *order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int c[3];
        for (int k = 0; k < 3; k++) {
            int ci = (int)floor((rods[i].pos[k] - lo[k]) / len[k] * nc[k]);
            if (ci < 0) ci = 0;
            if (ci >= nc[k]) ci = nc[k] - 1;
            c[k] = ci;
        }
        cellOf[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
        start[cellOf[i] + 1]++;
    }
    for (long long c = 0; c < ncell; c++) start[c + 1] += start[c];
    int *fill = (int *)malloc(sizeof(int) * (size_t)ncell);
    memcpy(fill, start, sizeof(int) * (size_t)ncell);
    for (int i = 0; i < n; i++) order[fill[cellOf[i]]++] = i;
    free(fill);

    const int nt = omp_get_max_threads();
    orc_block **tb = (orc_block **)calloc((size_t)nt, sizeof(orc_block *));
    long long *tn = (long long *)calloc((size_t)nt, sizeof(long long));
    long long *tc = (long long *)calloc((size_t)nt, sizeof(long long));
#pragma omp parallel
    {
        const int tid = omp_get_thread_num();
#pragma omp for schedule(dynamic, 8)
        for (long long c = 0; c < ncell; c++) {
            const int cx = (int)(c % nc[0]), cy = (int)((c / nc[0]) % nc[1]), cz = (int)(c / ((long long)nc[0] * nc[1]));
            for (int dz = -1; dz <= 1; dz++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        int o[3] = {cx + dx, cy + dy, cz + dz};
                        double shift[3] = {0, 0, 0};
                        int ok = 1;
                        for (int k = 0; k < 3; k++) {
                            if (o[k] < 0) {
                                if (!pbc[k]) { ok = 0; break; }
                                o[k] += nc[k];
                                shift[k] = -1 * len[k];
                            } else if (o[k] >= nc[k]) {
                                if (!pbc[k]) { ok = 0; break; }
                                o[k] -= nc[k];
                                shift[k] = 1 * len[k];
                            }
                        }
                        if (!ok) continue;
                        const long long cj = ((long long)o[2] * nc[1] + o[1]) * nc[0] + o[0];
                        for (int ii = start[c]; ii < start[c + 1]; ii++) {
                            const orc_rod *I = &rods[order[ii]];
                            for (int jj = start[cj]; jj < start[cj + 1]; jj++) {
                                const orc_rod *J = &rods[order[jj]];
                                if (I->gid >= J->gid) continue;
                                orc_block blk;
                                if (try_pair(I, J, shift, withStress, &blk)) {
                                    if (tn[tid] == tc[tid]) {
                                        tc[tid] = tc[tid] ? tc[tid] * 2 : 1024;
                                        tb[tid] = (orc_block *)realloc(tb[tid], sizeof(orc_block) * (size_t)tc[tid]);
                                    }
                                    tb[tid][tn[tid]++] = blk;
                                }
                            }
                        }
                    }
        }
    }
    long long cnt = 0;
    for (int t = 0; t < nt; t++) {
        for (long long i = 0; i < tn[t]; i++) {
            if (cnt < cap) out[cnt] = tb[t][i];
            cnt++;
        }
        free(tb[t]);
    }
    free(tb); free(tn); free(tc); free(cellOf); free(start); free(order);
    if (cnt <= cap) qsort(out, (size_t)cnt, sizeof(orc_block), cmp_block);
    return cnt;
}

/* ------------------------------------------------------------------ assembly */
/* Sylinder.cpp:69-82 (isSphere uses length < 2*radius, Sylinder.cpp:61-67) */
void orc_drag_coeff(double radius, double length, double mu, double *dragPara, double *dragPerp, double *dragRot) {
    const double Pi = 3.14159265358979323846;
    if (length < radius * 2) {
        const double rad = 0.5 * length + radius;
        *dragPara = 6 * Pi * rad * mu;
        *dragPerp = *dragPara;
        *dragRot = 8 * Pi * rad * rad * rad * mu;
    } else {
        const double b = -(1 + 2 * log(radius / (length)));
        *dragPara = 8 * Pi * length * mu / (2 * b);
        *dragPerp = 8 * Pi * length * mu / (b + 2);
        *dragRot = 2 * Pi * mu * length * length * length / (3 * (b + 2));
    }
}

void orc_csr_free(orc_csr *A) {
    free(A->rowptr); free(A->col); free(A->val);
    memset(A, 0, sizeof(*A));
}

/* ConstraintCollector.cpp:261-342 (rows/values) and :405-420 (vectors) */
int orc_build_dtrans(long long nc, const orc_block *blocks, int nRodsGlobal, orc_csr *DT, double *delta0,
                     double *invKappa, double *biFlag, double *gammaGuess) {
    (void)nRodsGlobal;
    DT->n = (int)nc;
    DT->rowptr = (long long *)malloc(sizeof(long long) * (size_t)(nc + 1));
    DT->rowptr[0] = 0;
    for (long long k = 0; k < nc; k++) DT->rowptr[k + 1] = DT->rowptr[k] + (blocks[k].oneSide ? 6 : 12);
    DT->nnz = DT->rowptr[nc];
    DT->col = (int *)malloc(sizeof(int) * (size_t)(DT->nnz > 0 ? DT->nnz : 1));
    DT->val = (double *)malloc(sizeof(double) * (size_t)(DT->nnz > 0 ? DT->nnz : 1));
#pragma omp parallel for
    for (long long k = 0; k < nc; k++) {
        const orc_block *b = &blocks[k];
        long long kk = DT->rowptr[k];
        for (int side = 0; side < (b->oneSide ? 1 : 2); side++) {
            const int gi = side == 0 ? b->globalIndexI : b->globalIndexJ;
            const double *g = side == 0 ? b->normI : b->normJ;
            const double *p = side == 0 ? b->posI : b->posJ;
            for (int c = 0; c < 6; c++) DT->col[kk + c] = 6 * gi + c;
            DT->val[kk + 0] = g[0];
            DT->val[kk + 1] = g[1];
            DT->val[kk + 2] = g[2];
            DT->val[kk + 3] = (g[2] * p[1] - g[1] * p[2]);
            DT->val[kk + 4] = (g[0] * p[2] - g[2] * p[0]);
            DT->val[kk + 5] = (g[1] * p[0] - g[0] * p[1]);
            kk += 6;
        }
        delta0[k] = b->delta0;
        gammaGuess[k] = b->gamma;
        invKappa[k] = 0;
        biFlag[k] = 0;
        if (b->bilateral) {
            invKappa[k] = b->kappa > 0 ? 1 / b->kappa : 0;
            biFlag[k] = 1;
        }
    }
    return 0;
}

/* explicit transpose as ConstraintOperator.cpp:14-20 does each step; rows of the result keep
 * ascending column (= constraint) order */
int orc_transpose(const orc_csr *A, int ncols, orc_csr *AT) {
    AT->n = ncols;
    AT->nnz = A->nnz;
    AT->rowptr = (long long *)calloc((size_t)ncols + 1, sizeof(long long));
    AT->col = (int *)malloc(sizeof(int) * (size_t)(A->nnz > 0 ? A->nnz : 1));
    AT->val = (double *)malloc(sizeof(double) * (size_t)(A->nnz > 0 ? A->nnz : 1));
    for (long long p = 0; p < A->nnz; p++) AT->rowptr[A->col[p] + 1]++;
    for (int c = 0; c < ncols; c++) AT->rowptr[c + 1] += AT->rowptr[c];
    long long *fill = (long long *)malloc(sizeof(long long) * (size_t)(ncols > 0 ? ncols : 1));
    memcpy(fill, AT->rowptr, sizeof(long long) * (size_t)ncols);
    for (int r = 0; r < A->n; r++)
        for (long long p = A->rowptr[r]; p < A->rowptr[r + 1]; p++) {
            const long long q = fill[A->col[p]]++;
            AT->col[q] = r;
            AT->val[q] = A->val[p];
        }
    free(fill);
    return 0;
}

/* SylinderSystem.cpp:622-717: 18 nnz per rod */
int orc_build_mobility(int n, const orc_rod *rods, const int *immovable, double mu, orc_csr *M) {
    M->n = 6 * n;
    M->nnz = 18LL * n;
    M->rowptr = (long long *)malloc(sizeof(long long) * (size_t)(6 * n + 1));
    M->col = (int *)malloc(sizeof(int) * (size_t)(18 * (n > 0 ? n : 1)));
    M->val = (double *)malloc(sizeof(double) * (size_t)(18 * (n > 0 ? n : 1)));
    for (int i = 0; i <= 6 * n; i++) M->rowptr[i] = 3LL * i;
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        const double *q = rods[i].direction;
        double dPara, dPerp, dRot;
        orc_drag_coeff(rods[i].radius, rods[i].length, mu, &dPara, &dPerp, &dRot);
        const int imm = immovable ? immovable[i] : 0;
        const double iPara = imm ? 0.0 : 1 / dPara;
        const double iPerp = imm ? 0.0 : 1 / dPerp;
        const double iRot = imm ? 0.0 : 1 / dRot;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                const double qq = q[r] * q[c];
                const double Imqq = (r == c ? 1.0 : 0.0) - qq;
                M->col[18 * i + 3 * r + c] = 6 * i + c;
                M->val[18 * i + 3 * r + c] = iPara * qq + iPerp * Imqq;
                M->col[18 * i + 9 + 3 * r + c] = 6 * i + 3 + c;
                M->val[18 * i + 9 + 3 * r + c] = iRot * qq + iRot * Imqq;
            }
    }
    return 0;
}

void orc_spmv(const orc_csr *A, const double *x, double *y, double alpha, double beta, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < A->n; r++) {
        double s = 0;
        for (long long p = A->rowptr[r]; p < A->rowptr[r + 1]; p++) s += A->val[p] * x[A->col[p]];
        y[r] = (beta == 0.0) ? alpha * s : beta * y[r] + alpha * s;
    }
}

/* ------------------------------------------------------------------ deterministic vector kernels */
#define ORC_CHUNK 4096
static double vdot(long long n, const double *a, const double *b) {
    const long long nch = (n + ORC_CHUNK - 1) / ORC_CHUNK;
    double *part = (double *)malloc(sizeof(double) * (size_t)(nch > 0 ? nch : 1));
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < nch; c++) {
        const long long e = (c + 1) * ORC_CHUNK < n ? (c + 1) * ORC_CHUNK : n;
        double s = 0;
        for (long long i = c * ORC_CHUNK; i < e; i++) s += a[i] * b[i];
        part[c] = s;
    }
    double s = 0;
    for (long long c = 0; c < nch; c++) s += part[c];
    free(part);
    return s;
}
/* pow(v.norm2(), 2) exactly as the reference spells it (BCQPSolver.cpp:215,220,321): the square root and the
 * squaring each round, so the result can differ from the plain dot product in the last bit */
static double vnorm2sq(long long n, const double *a) {
    /* gcc folds the reference's pow(x, 2) into x * x (so does every compiler the reference is built with at -O2/-O3;
     * glibc's pow itself differs from x * x in the last bit for about 1 argument in 1000) */
    const double nrm = sqrt(vdot(n, a, a));
    return nrm * nrm;
}
static double vnorminf(long long n, const double *a) {
    double m = 0;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (long long i = 0; i < n; i++) {
        const double v = fabs(a[i]);
        if (v > m) m = v;
    }
    return m;
}
/* y = alpha*a + beta*b + gamma*y  (Tpetra update) */
static void vupdate2(long long n, double *y, double alpha, const double *a, double beta, const double *b,
                     double gamma) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = (gamma == 0.0 ? 0.0 : gamma * y[i]) + alpha * a[i] + beta * b[i];
}
static void vupdate1(long long n, double *y, double alpha, const double *a, double beta) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * a[i];
}

/* ------------------------------------------------------------------ operator + BCQP */
typedef struct orc_op {
    /* generic A (CSR) or the constraint operator D^T M D + invKappa */
    const orc_csr *A;
    const orc_csr *DT, *D, *M;
    const double *invKappa;
    double *force, *vel;
    long long n;
} orc_op;

/* ConstraintOperator.cpp:30-71 with alpha=1, beta=0 */
static void op_apply(const orc_op *op, const double *x, double *y) {
    if (op->A) {
        orc_spmv(op->A, x, y, 1.0, 0.0, 0);
        return;
    }
    orc_spmv(op->D, x, op->force, 1.0, 0.0, 0);
    orc_spmv(op->M, op->force, op->vel, 1.0, 0.0, 0);
    orc_spmv(op->DT, op->vel, y, 1.0, 0.0, 0);
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < op->n; k++) y[k] += 1.0 * op->invKappa[k] * x[k];
}

/* BCQPSolver.cpp:431-459 */
static void bound_projection(long long n, double *v, const double *lb, const double *ub) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; i++) {
        double t = v[i];
        t = t > lb[i] ? t : lb[i]; /* std::max(temp, lb) */
        t = t < ub[i] ? t : ub[i]; /* std::min(temp, ub) */
        v[i] = t;
    }
}

/* BCQPSolver.cpp:461-497 */
static double projection_residual(long long n, const double *x, const double *y, const double *lb, const double *ub,
                                  double *q, int *err) {
    const double eps = DBL_EPSILON * 100;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (long long i = 0; i < n; i++) {
        if (x[i] < lb[i] + eps) {
            q[i] = y[i] < 0.0 ? y[i] : 0.0;
        } else if (x[i] > ub[i] - eps) {
            q[i] = y[i] > 0.0 ? y[i] : 0.0;
        } else if (x[i] > lb[i] && x[i] < ub[i]) {
            q[i] = y[i];
        } else {
            bad = 1;
        }
    }
    if (bad && err) *err = 1;
    return vnorminf(n, q);
}

static void push_hist(orc_hist *h, int cap, int *nh, double a, double b, double c, double d, double e, double f) {
    if (h && *nh < cap) {
        h[*nh].v[0] = a; h[*nh].v[1] = b; h[*nh].v[2] = c;
        h[*nh].v[3] = d; h[*nh].v[4] = e; h[*nh].v[5] = f;
    }
    (*nh)++;
}

/* BCQPSolver.cpp:134-247.  x: in = initial guess, out = returned iterate (including the iteMax quirk:
 * the swap at :237-238 has already happened, so the OLDER iterate is returned). */
static int solve_bbpgd(const orc_op *op, const double *b, const double *lb, const double *ub, double *x, double tol,
                       int iteMax, orc_hist *hist, int cap, int *nh, int *mvOut, int *iteOut, double *resOut) {
    const long long n = op->n;
    int mvCount = 0, iteCount = 0, err = 0, stag = 0;
    double *xk = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *xkm1 = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *gk = (double *)calloc((size_t)n + 1, sizeof(double));
    double *gkm1 = (double *)calloc((size_t)n + 1, sizeof(double));
    double *gkdiff = (double *)calloc((size_t)n + 1, sizeof(double));
    double *xkdiff = (double *)calloc((size_t)n + 1, sizeof(double));
    memcpy(xk, x, sizeof(double) * (size_t)n);
    memcpy(xkm1, x, sizeof(double) * (size_t)n);

    op_apply(op, xkm1, gkm1);
    mvCount++;
    vupdate1(n, gkm1, 1.0, b, 1.0);
    double resPhi = projection_residual(n, xkm1, gkm1, lb, ub, xkdiff, &err);
    push_hist(hist, cap, nh, 1.0 * iteCount, 0, 0, 0, resPhi, 1.0 * mvCount);
    double *ret = xkm1;
    if (fabs(resPhi) < tol) {
        ret = xkm1;
        goto done;
    }
    {
        double alpha = 1.0 / vnorminf(n, xkdiff);
        while (iteCount < iteMax) {
            iteCount++;
            vupdate2(n, xk, -alpha, gkm1, 1.0, xkm1, 0.0);
            bound_projection(n, xk, lb, ub);
            op_apply(op, xk, gk);
            mvCount++;
            vupdate1(n, gk, 1.0, b, 1.0);
            resPhi = projection_residual(n, xk, gk, lb, ub, xkdiff, &err);
            push_hist(hist, cap, nh, 1.0 * iteCount, 0, 0, alpha, resPhi, 1.0 * mvCount);
            if (fabs(resPhi) < tol) break;
            vupdate2(n, xkdiff, 1.0, xk, -1.0, xkm1, 0.0);
            vupdate2(n, gkdiff, 1.0, gk, -1.0, gkm1, 0.0);
            double a = 0, bb = 0;
            if (iteCount % 2 == 0) {
                a = vnorm2sq(n, xkdiff);
                bb = vdot(n, xkdiff, gkdiff);
            } else {
                a = vdot(n, xkdiff, gkdiff);
                bb = vnorm2sq(n, gkdiff);
            }
            if (fabs(bb) < 10 * DBL_EPSILON) bb += 10 * DBL_EPSILON;
            alpha = a / bb;
            if (alpha < DBL_EPSILON * 10) {
                stag = 1;
                break;
            }
            double *t;
            t = xkm1; xkm1 = xk; xk = t;
            t = gkm1; gkm1 = gk; gk = t;
        }
        ret = xk;
    }
done:
    memcpy(x, ret, sizeof(double) * (size_t)n);
    if (mvOut) *mvOut = mvCount;
    if (iteOut) *iteOut = iteCount;
    if (resOut) *resOut = resPhi;
    free(xk); free(xkm1); free(gk); free(gkm1); free(gkdiff); free(xkdiff);
    if (err) return 2;
    return stag ? 1 : 0;
}

/* BCQPSolver.cpp:249-389 */
static int solve_apgd(const orc_op *op, const double *b, const double *lb, const double *ub, double *x, double tol,
                      int iteMax, orc_hist *hist, int cap, int *nh, int *mvOut, int *iteOut, double *resOut) {
    const long long n = op->n;
    int mvCount = 0, err = 0, stag = 0;
    const size_t sz = sizeof(double) * (size_t)(n + 1);
    double *xk = (double *)malloc(sz), *yk = (double *)malloc(sz);
    double *xkp1 = (double *)calloc((size_t)n + 1, 8), *ykp1 = (double *)calloc((size_t)n + 1, 8);
    double *gVec = (double *)calloc((size_t)n + 1, 8), *tempVec = (double *)calloc((size_t)n + 1, 8);
    double *xhatk = (double *)malloc(sz), *xkdiff = (double *)calloc((size_t)n + 1, 8);
    double *Axb = (double *)calloc((size_t)n + 1, 8), *Axbkp1 = (double *)calloc((size_t)n + 1, 8);
    memcpy(xk, x, sizeof(double) * (size_t)n);
    memcpy(yk, x, sizeof(double) * (size_t)n);
    for (long long i = 0; i < n; i++) xhatk[i] = 1.0;
    double thetak = 1, thetakp1 = 1;
    vupdate2(n, xkdiff, -1.0, xhatk, 1.0, xk, 0.0);
    op_apply(op, xkdiff, tempVec);
    mvCount++;
    const double tempNorm2 = sqrt(vdot(n, tempVec, tempVec));
    const double xkdiffNorm2 = sqrt(vdot(n, xkdiff, xkdiff));
    double Lk = (tempNorm2 / xkdiffNorm2);
    double tk = 1.0 / Lk;
    push_hist(hist, cap, nh, 0, 0, 0, tk, 0, 1.0 * mvCount);
    int iteCount = 0;
    double resmin = DBL_MAX, resPhi = 0;
    while (iteCount < iteMax) {
        iteCount++;
        op_apply(op, yk, Axb);
        mvCount++;
        vupdate2(n, gVec, 1.0, b, 1.0, Axb, 0.0);
        vupdate2(n, xkp1, 1.0, yk, -tk, gVec, 0.0);
        bound_projection(n, xkp1, lb, ub);
        const double rightTerm1 = vdot(n, yk, Axb) * 0.5;
        const double rightTerm2 = vdot(n, yk, b);
        while (1) {
            vupdate2(n, xkdiff, 1.0, xkp1, -1.0, yk, 0.0);
            op_apply(op, xkp1, Axbkp1);
            mvCount++;
            const double leftTerm1 = vdot(n, xkp1, Axbkp1) * 0.5;
            const double leftTerm2 = vdot(n, xkp1, b);
            const double rightTerm3 = vdot(n, gVec, xkdiff);
            const double rightTerm4 = 0.5 * Lk * vnorm2sq(n, xkdiff);
            if ((leftTerm1 + leftTerm2) <= (rightTerm1 + rightTerm2 + rightTerm3 + rightTerm4)) break;
            Lk *= 2;
            tk = 1 / Lk;
            vupdate2(n, xkp1, 1.0, yk, -tk, gVec, 0.0);
            bound_projection(n, xkp1, lb, ub);
        }
        if (tk < DBL_EPSILON * 10) {
            stag = 1;
            break;
        }
        thetakp1 = (-thetak * thetak + thetak * sqrt(4 + thetak * thetak)) / 2;
        const double betakp1 = thetak * (1 - thetak) / (thetak * thetak + thetakp1);
        vupdate2(n, ykp1, (1 + betakp1), xkp1, -betakp1, xk, 0.0);
        vupdate1(n, Axbkp1, 1.0, b, 1.0);
        resPhi = fabs(projection_residual(n, xkp1, Axbkp1, lb, ub, tempVec, &err));
        if (resPhi < resmin) {
            resmin = resPhi;
            memcpy(xhatk, xkp1, sizeof(double) * (size_t)n);
        }
        push_hist(hist, cap, nh, 1.0 * iteCount, 0, 0, tk, resPhi, 1.0 * mvCount);
        if (resPhi < tol) break;
        vupdate2(n, tempVec, 1.0, xkp1, -1.0, xk, 0.0);
        if (vdot(n, gVec, tempVec) > 0) {
            memcpy(ykp1, xkp1, sizeof(double) * (size_t)n);
            thetakp1 = 1;
        }
        Lk *= 0.9;
        tk = 1 / Lk;
        double *t;
        t = yk; yk = ykp1; ykp1 = t;
        t = xk; xk = xkp1; xkp1 = t;
        thetak = thetakp1;
    }
    memcpy(x, xhatk, sizeof(double) * (size_t)n);
    if (mvOut) *mvOut = mvCount;
    if (iteOut) *iteOut = iteCount;
    if (resOut) *resOut = resPhi;
    free(xk); free(yk); free(xkp1); free(ykp1); free(gVec); free(tempVec); free(xhatk); free(xkdiff);
    free(Axb); free(Axbkp1);
    if (err) return 2;
    return stag ? 1 : 0;
}

int orc_bcqp_csr(const orc_csr *A, const double *b, const double *lb, const double *ub, double *x, double tol,
                 int maxIte, int solverChoice, orc_hist *hist, int histCap, int *nHist) {
    orc_op op;
    memset(&op, 0, sizeof(op));
    op.A = A;
    op.n = A->n;
    int nh = 0, rc;
    if (solverChoice == 1)
        rc = solve_apgd(&op, b, lb, ub, x, tol, maxIte, hist, histCap, &nh, NULL, NULL, NULL);
    else
        rc = solve_bbpgd(&op, b, lb, ub, x, tol, maxIte, hist, histCap, &nh, NULL, NULL, NULL);
    if (nHist) *nHist = nh;
    return rc;
}

int orc_operator_apply(const orc_block *blocks, long long nc, const orc_rod *rods, const int *immovable, int nRods,
                       double mu, double dt, const double *x, double *y, double *force, double *vel) {
    orc_csr DT, D, M;
    double *delta0 = (double *)malloc(8 * (size_t)(nc + 1)), *invK = (double *)malloc(8 * (size_t)(nc + 1));
    double *bi = (double *)malloc(8 * (size_t)(nc + 1)), *g0 = (double *)malloc(8 * (size_t)(nc + 1));
    orc_build_dtrans(nc, blocks, nRods, &DT, delta0, invK, bi, g0);
    for (long long k = 0; k < nc; k++) invK[k] *= 1.0 / dt;
    orc_transpose(&DT, 6 * nRods, &D);
    orc_build_mobility(nRods, rods, immovable, mu, &M);
    orc_op op;
    memset(&op, 0, sizeof(op));
    op.DT = &DT; op.D = &D; op.M = &M; op.invKappa = invK; op.n = nc;
    op.force = force; op.vel = vel;
    op_apply(&op, x, y);
    orc_csr_free(&DT); orc_csr_free(&D); orc_csr_free(&M);
    free(delta0); free(invK); free(bi); free(g0);
    return 0;
}

/* ConstraintSolver::setup (:4-34) + solveConstraints (:60-107) */
int orc_solve_constraints(const orc_block *blocks, const orc_rod *rods, const int *immovable, double mu,
                          const double *velNonCon, orc_solve_info *info, double *gamma, double *forceU, double *velU,
                          double *forceB, double *velB, orc_hist *hist, int histCap, int *nHist) {
    if (info->nthreads > 0) omp_set_num_threads(info->nthreads);
    const long long nc = info->nc;
    const int nR = info->nRods;
    const double dt = info->dt;
    const double t0 = omp_get_wtime();
    orc_csr DT, D, M;
    const size_t vs = 8 * (size_t)(nc + 1);
    double *delta0 = (double *)malloc(vs), *invK = (double *)malloc(vs), *bi = (double *)malloc(vs);
    double *deltanc = (double *)calloc((size_t)nc + 1, 8), *q = (double *)malloc(vs);
    double *lb = (double *)malloc(vs), *ub = (double *)malloc(vs), *gammaBi = (double *)malloc(vs);
    double *force = (double *)calloc((size_t)6 * nR + 1, 8), *vel = (double *)calloc((size_t)6 * nR + 1, 8);
    orc_build_mobility(nR, rods, immovable, mu, &M); /* prepareStep: calcMobOperator */
    orc_build_dtrans(nc, blocks, nR, &DT, delta0, invK, bi, gamma);
    for (long long k = 0; k < nc; k++) {
        delta0[k] *= 1.0 / dt;
        invK[k] *= 1.0 / dt;
    }
    orc_spmv(&DT, velNonCon, deltanc, 1.0, 0.0, 0);
    orc_transpose(&DT, 6 * nR, &D);
    vupdate2(nc, q, 1.0, delta0, 1.0, deltanc, 0.0);
    /* bounds: BCQPSolver.cpp:499-510 then ConstraintSolver.cpp:69 */
    for (long long k = 0; k < nc; k++) {
        ub[k] = DBL_MAX / 10;
        lb[k] = (-DBL_MAX * .1) * bi[k];
    }
    info->tAssemble = omp_get_wtime() - t0;
    const double t1 = omp_get_wtime();
    orc_op op;
    memset(&op, 0, sizeof(op));
    op.DT = &DT; op.D = &D; op.M = &M; op.invKappa = invK; op.n = nc;
    op.force = force; op.vel = vel;
    int nh = 0, rc;
    const double tol = info->res * (1.0 / dt);
    if (info->solverChoice == 1)
        rc = solve_apgd(&op, q, lb, ub, gamma, tol, info->maxIte, hist, histCap, &nh, &info->mvCount, &info->nIte,
                        &info->resFinal);
    else
        rc = solve_bbpgd(&op, q, lb, ub, gamma, tol, info->maxIte, hist, histCap, &nh, &info->mvCount, &info->nIte,
                         &info->resFinal);
    if (nHist) *nHist = nh;
    info->status = rc;
    /* split (:95-106): op.force/op.vel hold the LAST apply */
    for (long long k = 0; k < nc; k++) gammaBi[k] = 1.0 * gamma[k] * bi[k];
    orc_spmv(&D, gammaBi, forceB, 1.0, 0.0, 0);
    orc_spmv(&M, forceB, velB, 1.0, 0.0, 0);
    vupdate2(6LL * nR, forceU, 1.0, force, -1.0, forceB, 0.0);
    vupdate2(6LL * nR, velU, 1.0, vel, -1.0, velB, 0.0);
    info->tSolve = omp_get_wtime() - t1;
    orc_csr_free(&DT); orc_csr_free(&D); orc_csr_free(&M);
    free(delta0); free(invK); free(bi); free(deltanc); free(q); free(lb); free(ub); free(gammaBi);
    free(force); free(vel);
    return rc;
}

void orc_writeback_gamma(long long nc, orc_block *blocks, const double *gamma) {
#pragma omp parallel for
    for (long long k = 0; k < nc; k++) {
        blocks[k].gamma = gamma[k];
        for (int c = 0; c < 9; c++) blocks[k].stress[c] *= blocks[k].gamma;
    }
}

/* ---------------------------------------------------------------------------------------------------------
 * boundaries: Boundary::project of the three shipped shapes, arithmetic in the reference's expression order
 * (Eigen: a.norm() = sqrt(x^2+y^2+z^2), a.normalized() = a / norm) */
void orc_boundary_project(const orc_boundary *b, const double query[3], double project[3], double delta[3]) {
    if (b->type == 0) { /* SphereShell::project, Boundary.cpp:25-41 */
        double Q[3], Proj[3], PQ[3];
        sub3(query, b->center, Q);
        const double QueryR = norm3(Q);
        const double f = b->radius * (1 / QueryR);
        for (int k = 0; k < 3; k++) Proj[k] = f * Q[k];
        for (int k = 0; k < 3; k++) PQ[k] = Q[k] - Proj[k];
        const int out = QueryR > b->radius;
        if ((b->inside && out) || (!b->inside && !out))
            for (int k = 0; k < 3; k++) PQ[k] *= -1;
        for (int k = 0; k < 3; k++) {
            project[k] = Proj[k] + b->center[k];
            delta[k] = PQ[k];
        }
    } else if (b->type == 1) { /* Wall::project, Boundary.cpp:108-124 */
        double CQ[3], Proj[3], PQ[3];
        sub3(query, b->center, CQ);
        const double t = dot3(CQ, b->axis);
        for (int k = 0; k < 3; k++) Proj[k] = query[k] - t * b->axis[k];
        for (int k = 0; k < 3; k++) PQ[k] = query[k] - Proj[k];
        if (t < 0)
            for (int k = 0; k < 3; k++) PQ[k] *= -1;
        for (int k = 0; k < 3; k++) {
            project[k] = Proj[k];
            delta[k] = PQ[k];
        }
    } else { /* Tube::project, Boundary.cpp:185-209 */
        double CQ[3], ProjAxis[3], PAQ[3], Proj[3], d[3];
        sub3(query, b->center, CQ);
        const double t = dot3(CQ, b->axis);
        for (int k = 0; k < 3; k++) ProjAxis[k] = b->center[k] + t * b->axis[k];
        sub3(query, ProjAxis, PAQ);
        const double r = norm3(PAQ);
        for (int k = 0; k < 3; k++) Proj[k] = ProjAxis[k] + b->radius * (PAQ[k] / r);
        sub3(query, Proj, d);
        if (r > b->radius) {
            if (b->inside)
                for (int k = 0; k < 3; k++) d[k] *= -1;
        } else {
            if (!b->inside)
                for (int k = 0; k < 3; k++) d[k] *= -1;
        }
        for (int k = 0; k < 3; k++) {
            project[k] = Proj[k];
            delta[k] = d[k];
        }
    }
}

/* checkEnd of SylinderSystem.cpp:1111-1133 for one query point; returns 1 if a block was written */
static int boundary_check_end(const orc_boundary *b, const orc_rod *sy, const double Query[3], double radius, double colBuf,
                              orc_block *blk) {
    double Proj[3], delta[3], norm[3], posI[3], QP[3];
    orc_boundary_project(b, Query, Proj, delta);
    const double deltanorm = norm3(delta);
    const double inv = 1 / deltanorm;
    for (int k = 0; k < 3; k++) norm[k] = delta[k] * inv;
    sub3(Query, sy->pos, posI);
    sub3(Query, Proj, QP);
    double d0;
    if (dot3(QP, delta) < 0) d0 = -deltanorm - radius;                              /* outside the boundary */
    else if (deltanorm < (1 + colBuf * 2) * sy->radiusCollision) d0 = deltanorm - radius; /* inside but close */
    else return 0;
    memset(blk, 0, sizeof(*blk));
    blk->delta0 = d0;
    blk->gamma = 0;
    blk->gidI = blk->gidJ = sy->gid;
    blk->globalIndexI = blk->globalIndexJ = sy->globalIndex;
    blk->oneSide = 1;
    blk->bilateral = 0;
    blk->kappa = 0;
    for (int k = 0; k < 3; k++) {
        blk->normI[k] = blk->normJ[k] = norm[k];
        blk->posI[k] = blk->posJ[k] = posI[k];
        blk->labI[k] = Query[k];
        blk->labJ[k] = Proj[k];
    }
    return 1;
}

long long orc_collect_boundary(int n, const orc_rod *rods, int nb, const orc_boundary *bnd, double colBuf, orc_block *out,
                               long long cap) {
    long long cnt = 0;
    orc_block tmp;
    for (int ib = 0; ib < nb; ib++) {
        for (int i = 0; i < n; i++) {
            const orc_rod *sy = &rods[i];
            if (is_sphere_col(sy)) { /* SylinderSystem.cpp:1135-1137 */
                const double radius = sy->lengthCollision * 0.5 + sy->radiusCollision;
                if (boundary_check_end(&bnd[ib], sy, sy->pos, radius, colBuf, &tmp)) {
                    if (cnt < cap) out[cnt] = tmp;
                    cnt++;
                }
            } else { /* :1138-1146 */
                double Qm[3], Qp[3];
                const double h = sy->lengthCollision * 0.5;
                for (int k = 0; k < 3; k++) {
                    Qm[k] = sy->pos[k] - sy->direction[k] * h;
                    Qp[k] = sy->pos[k] + sy->direction[k] * h;
                }
                if (boundary_check_end(&bnd[ib], sy, Qm, sy->radiusCollision, colBuf, &tmp)) {
                    if (cnt < cap) out[cnt] = tmp;
                    cnt++;
                }
                if (boundary_check_end(&bnd[ib], sy, Qp, sy->radiusCollision, colBuf, &tmp)) {
                    if (cnt < cap) out[cnt] = tmp;
                    cnt++;
                }
            }
        }
    }
    return cnt;
}

/* ---------------------------------------------------------------------------------------------------------
 * bilateral links (SylinderSystem.cpp:1386-1482) */
static void pbc_image1(double lb, double ub, double *x) { /* Util/GeoUtil.hpp:27-36 */
    const double L = ub - lb;
    while (*x >= ub) *x -= L;
    while (*x < lb) *x += L;
}
static void pbc_image2(double lb, double ub, double *x, double *trg) { /* Util/GeoUtil.hpp:51-60 */
    pbc_image1(lb, ub, trg);
    double dist = *x - *trg;
    pbc_image1(0.0, ub - lb, &dist);
    if (dist > (ub - lb) * 0.5) *x = *trg + dist - (ub - lb);
    else *x = *trg + dist;
}
typedef struct {
    int gid, idx;
} gid_idx;
static int cmp_gid(const void *a, const void *b) {
    const int x = ((const gid_idx *)a)->gid, y = ((const gid_idx *)b)->gid;
    return (x > y) - (x < y);
}

long long orc_collect_links(int n, const orc_rod *rods, long long nLinks, const int *prevGid, const int *nextGid,
                            const double boxLow[3], const double boxHigh[3], const int pbc[3], double linkKappa,
                            double linkGap, orc_block *out) {
    gid_idx *map = (gid_idx *)malloc(sizeof(gid_idx) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        map[i].gid = rods[i].gid;
        map[i].idx = i;
    }
    qsort(map, (size_t)n, sizeof(gid_idx), cmp_gid);
    long long cnt = 0;
    for (long long l = 0; l < nLinks; l++) {
        gid_idx key = {prevGid[l], 0};
        const gid_idx *fi = (const gid_idx *)bsearch(&key, map, (size_t)n, sizeof(gid_idx), cmp_gid);
        key.gid = nextGid[l];
        const gid_idx *fj = (const gid_idx *)bsearch(&key, map, (size_t)n, sizeof(gid_idx), cmp_gid);
        if (!fi || !fj) {
            free(map);
            return -1;
        }
        const orc_rod *syI = &rods[fi->idx], *syJ = &rods[fj->idx];
        double centerJ[3] = {syJ->pos[0], syJ->pos[1], syJ->pos[2]};
        for (int k = 0; k < 3; k++) { /* :1436-1448 */
            if (!pbc[k]) continue;
            double trg = syI->pos[k], xk = centerJ[k];
            pbc_image2(boxLow[k], boxHigh[k], &xk, &trg);
            centerJ[k] = xk;
        }
        double Pp[3], Qm[3], rvec[3], PQ[3];
        const double hI = 0.5 * syI->length, hJ = 0.5 * syJ->length;
        for (int k = 0; k < 3; k++) {
            Pp[k] = syI->pos[k] + syI->direction[k] * hI; /* plus end of I */
            Qm[k] = centerJ[k] - syJ->direction[k] * hJ;   /* minus end of J */
        }
        sub3(Qm, Pp, rvec);
        const double rnorm = norm3(rvec);
        const double delta0 = rnorm - syI->radius - syJ->radius - linkGap;
        sub3(Pp, Qm, PQ);
        const double pqn = norm3(PQ);
        orc_block *b = &out[cnt++];
        memset(b, 0, sizeof(*b));
        b->delta0 = delta0;
        b->gamma = delta0 < 0 ? -delta0 : 0;
        b->gidI = syI->gid;
        b->gidJ = syJ->gid;
        b->globalIndexI = syI->globalIndex;
        b->globalIndexJ = syJ->globalIndex;
        b->oneSide = 0;
        b->bilateral = 1;
        b->kappa = linkKappa;
        for (int k = 0; k < 3; k++) {
            b->normI[k] = pqn > 0 ? PQ[k] / pqn : PQ[k];
            b->normJ[k] = -b->normI[k];
            b->posI[k] = Pp[k] - syI->pos[k];
            b->posJ[k] = Qm[k] - centerJ[k];
            b->labI[k] = Pp[k];
            b->labJ[k] = Qm[k];
        }
        orc_collide_stress(syI->direction, syJ->direction, syI->pos, centerJ, syI->length, syJ->length, syI->radius,
                           syJ->radius, 1.0, Pp, Qm, b->stress);
    }
    free(map);
    return cnt;
}
