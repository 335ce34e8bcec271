// geometry.cuh -- device-side narrow phase: segment/segment and point/segment closest points and
// the spherocylinder pair functor.
//
// Follows the algorithm of SimToolbox/Collision/DCPQuery.hpp:91-128,199-472 and the pair functor
// SimToolbox/Sylinder/SylinderNear.hpp:197-414 (citations relative to the aLENS tree).
//
// Every translation unit that includes this file MUST be compiled with -fmad=false: the decision
// `sep < buffer` and the exact-comparison clamps of the closest-point search have to round exactly
// like the CPU path (built with -ffp-contract=off in the oracle) so that the integer pair list is
// reproducible bit for bit.  fp64 div/sqrt are IEEE-correct on the device by default.
#pragma once
#include <cfloat>

namespace alens {

struct Vec3 {
    double x, y, z;
};
__device__ __forceinline__ Vec3 v3(double x, double y, double z) { return Vec3{x, y, z}; }
__device__ __forceinline__ Vec3 operator+(Vec3 a, Vec3 b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 operator*(Vec3 a, double s) { return Vec3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ Vec3 neg(Vec3 a) { return Vec3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double norm(Vec3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) {
    return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// root of h(z) = h0 + slope*z clamped to [0,1]  (DCPQuery.hpp:310-340)
__device__ __forceinline__ double clampedRoot(double slope, double h0, double h1) {
    const double eps = DBL_EPSILON;
    double r;
    if (fabs(h0) < eps && fabs(h1) < eps) {
        r = 0.5;
    } else if (h0 < 0) {
        if (h1 > 0) {
            r = -h0 / slope;
            r = r > 0.0 ? r : 0.0;
            r = r < 1.0 ? r : 1.0;
        } else {
            r = 1;
        }
    } else {
        r = 0;
    }
    return r;
}

struct SegSegCoef {
    double A, B, C, D, E;
    double F00, F10, F01, F11;
    double G00, G10, G01, G11;
};

// end point of the dR/ds = 0 line on edge s=0 (e=0) or s=1 (e=1)   (DCPQuery.hpp:357-423)
__device__ __forceinline__ double edgeT(const SegSegCoef &q, int e) {
    double v = (e == 0 ? q.F00 : q.F10) / q.B;
    if (v < 0 || v > 1) v = 0.5;
    return v;
}

// Closest points of segments P0P1 and Q0Q1; returns the distance   (DCPQuery.hpp:199-308)
__device__ inline double segSegClosest(Vec3 P0, Vec3 P1, Vec3 Q0, Vec3 Q1, Vec3 &Ploc, Vec3 &Qloc) {
    SegSegCoef q;
    const Vec3 P1mP0 = P1 - P0, Q1mQ0 = Q1 - Q0, P0mQ0 = P0 - Q0;
    q.A = dot(P1mP0, P1mP0);
    q.B = dot(P1mP0, Q1mQ0);
    q.C = dot(Q1mQ0, Q1mQ0);
    q.D = dot(P1mP0, P0mQ0);
    q.E = dot(Q1mQ0, P0mQ0);
    q.F00 = q.D;
    q.F10 = q.F00 + q.A;
    q.F01 = q.F00 - q.B;
    q.F11 = q.F10 - q.B;
    q.G00 = -q.E;
    q.G10 = q.G00 - q.B;
    q.G01 = q.G00 + q.C;
    q.G11 = q.G10 + q.C;

    double s, t;
    if (q.A > 0 && q.C > 0) {
        const double s0 = clampedRoot(q.A, q.F00, q.F10);
        const double s1 = clampedRoot(q.A, q.F01, q.F11);
        const int c0 = s0 <= 0 ? -1 : (s0 >= 1 ? 1 : 0);
        const int c1 = s1 <= 0 ? -1 : (s1 >= 1 ? 1 : 0);
        if (c0 == -1 && c1 == -1) {
            s = 0;
            t = clampedRoot(q.C, q.G00, q.G01);
        } else if (c0 == 1 && c1 == 1) {
            s = 1;
            t = clampedRoot(q.C, q.G10, q.G11);
        } else {
            // intersection of dR/ds = 0 with [0,1]^2: (edge, end) pairs   (DCPQuery.hpp:343-424)
            int e0, e1;
            double a00, a01, a10, a11; // end[0][0], end[0][1], end[1][0], end[1][1]
            if (c0 < 0) {
                e0 = 0; a00 = 0; a01 = edgeT(q, 0);
                if (c1 == 0) { e1 = 3; a10 = s1; a11 = 1; }
                else { e1 = 1; a10 = 1; a11 = edgeT(q, 1); }
            } else if (c0 == 0) {
                e0 = 2; a00 = s0; a01 = 0;
                if (c1 < 0) { e1 = 0; a10 = 0; a11 = edgeT(q, 0); }
                else if (c1 == 0) { e1 = 3; a10 = s1; a11 = 1; }
                else { e1 = 1; a10 = 1; a11 = edgeT(q, 1); }
            } else {
                e0 = 1; a00 = 1; a01 = edgeT(q, 1);
                if (c1 == 0) { e1 = 3; a10 = s1; a11 = 1; }
                else { e1 = 0; a10 = 0; a11 = edgeT(q, 0); }
            }
            // minimum of R along that segment   (DCPQuery.hpp:427-472)
            const double eps = DBL_EPSILON;
            const double delta = a11 - a01;
            const double h0 = delta * ((-q.B * a00 - q.E) + q.C * a01);
            const double h1 = delta * ((-q.B * a10 - q.E) + q.C * a11);
            if (fabs(h0) < fabs(q.C) * eps && fabs(h1) < fabs(q.C) * eps) {
                const double z = 0.5, omz = 1.0 - z;
                s = omz * a00 + z * a10;
                t = omz * a01 + z * a11;
            } else if (h0 >= 0) {
                if (e0 == 0) { s = 0; t = clampedRoot(q.C, q.G00, q.G01); }
                else if (e0 == 1) { s = 1; t = clampedRoot(q.C, q.G10, q.G11); }
                else { s = a00; t = a01; }
            } else if (h1 <= 0) {
                if (e1 == 0) { s = 0; t = clampedRoot(q.C, q.G00, q.G01); }
                else if (e1 == 1) { s = 1; t = clampedRoot(q.C, q.G10, q.G11); }
                else { s = a10; t = a11; }
            } else {
                const double z = clampedRoot(h1 - h0, h0, h1);
                const double omz = 1.0 - z;
                s = omz * a00 + z * a10;
                t = omz * a01 + z * a11;
            }
        }
    } else {
        if (q.A > 0) { s = clampedRoot(q.A, q.F00, q.F10); t = 0; }
        else if (q.C > 0) { s = 0; t = clampedRoot(q.C, q.G00, q.G01); }
        else { s = 0; t = 0; }
    }
    Ploc = v3((1.0 - s) * P0.x + s * P1.x, (1.0 - s) * P0.y + s * P1.y, (1.0 - s) * P0.z + s * P1.z);
    Qloc = v3((1.0 - t) * Q0.x + t * Q1.x, (1.0 - t) * Q0.y + t * Q1.y, (1.0 - t) * Q0.z + t * Q1.z);
    const Vec3 diff = Ploc - Qloc;
    return sqrt(dot(diff, diff));
}

// point / segment   (DCPQuery.hpp:91-128)
__device__ inline double pointSegClosest(Vec3 pt, Vec3 minus, Vec3 plus, Vec3 &perp) {
    const Vec3 direction = plus - minus;
    Vec3 diff = pt - plus;
    double t = dot(direction, diff);
    Vec3 closest;
    if (t >= 0) {
        closest = plus;
    } else {
        diff = pt - minus;
        t = dot(direction, diff);
        if (t <= 0) {
            closest = minus;
        } else {
            const double sqrLength = dot(direction, direction);
            if (sqrLength > 0) {
                t /= sqrLength;
                closest = v3(minus.x + t * direction.x, minus.y + t * direction.y, minus.z + t * direction.z);
            } else {
                closest = minus;
            }
        }
    }
    diff = pt - closest;
    perp = closest;
    return sqrt(dot(diff, diff));
}

// rod as the narrow phase sees it (the fields of SylinderNearEP that enter the functor)
struct RodGeom {
    Vec3 c;    // centre (image shift already applied for the source rod)
    Vec3 d;    // unit direction
    double lc; // lengthCollision
    double rc; // radiusCollision
};

struct Contact {
    double sep;   // delta0
    Vec3 normI;   // (Ploc-Qloc)/|Ploc-Qloc| (sign already adjusted for the reversed sphere case)
    Vec3 posI, posJ, labI, labJ;
};

// a = target (lower gid), b = source.  Same dispatch as CalcSylinderNearForce::operator()
// (SylinderNear.hpp:207-236): sphere iff lengthCollision < 2*radiusCollision.
__device__ inline bool pairContact(const RodGeom &a, const RodGeom &b, double buffer, Contact &out) {
    const bool sa = a.lc < 2 * a.rc, sb = b.lc < 2 * b.rc;
    Vec3 Ploc, Qloc;
    double sep;
    if (sa && sb) { // sp_sp, SylinderNear.hpp:253-294
        const double radI = a.lc * 0.5 + a.rc;
        const double radJ = b.lc * 0.5 + b.rc;
        const Vec3 rIJ = b.c - a.c;
        sep = norm(rIJ) - (radI + radJ);
        Ploc = a.c;
        Qloc = b.c;
    } else if (sa || sb) { // sp_sy, SylinderNear.hpp:307-355 (reverseIJ when the target is the sylinder)
        const RodGeom &sp = sa ? a : b;
        const RodGeom &sy = sa ? b : a;
        const double radI = sp.lc * 0.5 + sp.rc;
        const Vec3 Qm = sy.c - sy.d * (0.5 * sy.lc);
        const Vec3 Qp = sy.c + sy.d * (0.5 * sy.lc);
        Vec3 q;
        const double distMin = pointSegClosest(sp.c, Qm, Qp, q);
        sep = distMin - (radI + sy.rc);
        if (!(sep < buffer)) return false;
        const Vec3 dd = sp.c - q;
        const double n = norm(dd);
        const Vec3 nI = n > 0 ? v3(dd.x / n, dd.y / n, dd.z / n) : dd; // Eigen normalized(): unchanged when |dd| = 0
        const Vec3 pSp = sp.c - sp.c, pSy = q - sy.c;
        out.sep = sep;
        if (sa) {
            out.normI = nI; out.posI = pSp; out.posJ = pSy; out.labI = sp.c; out.labJ = q;
        } else {
            out.normI = neg(nI); out.posI = pSy; out.posJ = pSp; out.labI = q; out.labJ = sp.c;
        }
        return true;
    } else { // sy_sy, SylinderNear.hpp:367-414
        const Vec3 Pm = a.c - a.d * (0.5 * a.lc);
        const Vec3 Pp = a.c + a.d * (0.5 * a.lc);
        const Vec3 Qm = b.c - b.d * (0.5 * b.lc);
        const Vec3 Qp = b.c + b.d * (0.5 * b.lc);
        const double distMin = segSegClosest(Pm, Pp, Qm, Qp, Ploc, Qloc);
        sep = distMin - (a.rc + b.rc);
    }
    if (!(sep < buffer)) return false;
    const Vec3 dd = Ploc - Qloc;
    const double n = norm(dd);
    out.sep = sep;
    out.normI = n > 0 ? v3(dd.x / n, dd.y / n, dd.z / n) : dd; // Eigen normalized(): unchanged when |dd| = 0
    out.posI = Ploc - a.c;
    out.posJ = Qloc - b.c;
    out.labI = Ploc;
    out.labJ = Qloc;
    return true;
}

// collideStress for unit gamma   (SylinderNear.hpp:432-519), closed form of the 3^5-term epsilon loop
// kept in the reference's accumulation order (non-zero epsilon terms only).
__device__ inline void syN(double r, double h, double rho, double &a, double &b) {
    const double beta = h / 2.0 / r;
    const double s = 1.0; (void)s;
    a = 1.0 / 30.0 * (15.0 * beta + 8);
    b = 1.0 / 15.0 * (10.0 * beta * beta * beta + 20.0 * beta * beta + 15.0 * beta + 4.0);
    a = a * rho * r * r * r * r * r * 3.14159265358979323846;
    b = b * rho * r * r * r * r * r * 3.14159265358979323846;
}
__device__ inline void syGA(double r, double h, double rho, double &a, double &b) {
    const double beta = h / 2.0 / r;
    a = 1.0 / 30.0 * (20.0 * beta * beta * beta + 40.0 * beta * beta + 45.0 * beta + 16.0);
    b = 1.0 / 15.0 * (15 * beta + 8);
    a = a * 3.14159265358979323846 * r * r * r * r * r * rho;
    b = b * 3.14159265358979323846 * r * r * r * r * r * rho;
}
__device__ inline void isoPlusDyad(double a, double b, const double d[3], double out[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i][j] = a * (i == j ? 1.0 : 0.0) + (b - a) * (d[i] * d[j]);
}

__device__ inline void collideStress(Vec3 dirI_, Vec3 dirJ_, Vec3 cI_, Vec3 cJ_, double hI, double hJ, double rI,
                                     double rJ, double rho, Vec3 Ploc, Vec3 Qloc, double stress[9]) {
    const double dirI[3] = {dirI_.x, dirI_.y, dirI_.z}, dirJ[3] = {dirJ_.x, dirJ_.y, dirJ_.z};
    const double cI[3] = {cI_.x, cI_.y, cI_.z}, cJ[3] = {cJ_.x, cJ_.y, cJ_.z};
    double NI[3][3], NJ[3][3], iGI[3][3], iGJ[3][3];
    double aI, bI, aJ, bJ;
    syN(rI, hI, rho, aI, bI);
    syN(rJ, hJ, rho, aJ, bJ);
    isoPlusDyad(aI, bI, dirI, NI);
    isoPlusDyad(aJ, bJ, dirJ, NJ);
    syGA(rI, hI, rho, aI, bI);
    syGA(rJ, hJ, rho, aJ, bJ);
    aI = 1.0 / aI; bI = 1.0 / bI; aJ = 1.0 / aJ; bJ = 1.0 / bJ;
    isoPlusDyad(aI, bI, dirI, iGI);
    isoPlusDyad(aJ, bJ, dirJ, iGJ);
    Vec3 F1 = Qloc - Ploc;
    {
        const double n2 = dot(F1, F1);
        if (n2 > 0) {
            const double n = sqrt(n2);
            F1 = v3(F1.x / n, F1.y / n, F1.z / n);
        }
    }
    const Vec3 mF1 = neg(F1);
    const Vec3 xI = cross(Ploc - cI_, mF1), xJ = cross(Qloc - cJ_, F1);
    const double xICf[3] = {xI.x, xI.y, xI.z}, xJCf[3] = {xJ.x, xJ.y, xJ.z};
    const double f1[3] = {F1.x, F1.y, F1.z}, mf1[3] = {mF1.x, mF1.y, mF1.z};
    // epsilon[j][k][l] != 0 only for the 6 permutations; loop order i, j, k, l, r as the reference
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double sI = 0, sJ = 0;
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) {
                    if (j == k || k == l || j == l) continue;
                    const double e = ((l - k + 3) % 3 == 1 && (k - j + 3) % 3 == 1) ? 1.0 : -1.0;
                    for (int r = 0; r < 3; r++) {
                        sI = sI + NI[i][l] * e * iGI[k][r] * xICf[r];
                        sJ = sJ + NJ[i][l] * e * iGJ[k][r] * xJCf[r];
                    }
                }
            const double rIf = cI[i] * mf1[j];
            const double rJf = cJ[i] * f1[j];
            stress[3 * i + j] = ((rIf + rJf) + sI) + sJ;
        }
}

} // namespace alens
