/*
 * oracle/alens_oracle.h -- CPU restatement of the aLENS/SimToolbox collision-constraint hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or called by the product
 * (alens_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: the pair functor + closest-point code is PINNED against (a) the reference's own
 * known-answer vectors (SimToolbox/Sylinder/SylinderNear_test.cpp:43-111,172-241) and (b) the
 * reference's own DCPQuery.hpp and vendored FDPS tree compiled in place (oracle/ref_driver.cpp ->
 * oracle/_ref/libalens_ref.so).  The D/M assembly and the BBPGD/APGD loops restate Trilinos-typed
 * code that cannot be built here (Trilinos 12.18.1, Eigen >= 3.3 are not in /root/reference); for
 * those the reference holds no numeric golden vectors ("parity unpinned" by the reference's own
 * tests) -- they are cross-checked against scipy (tests/test_oracle_bcqp.py).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef ALENS_ORACLE_H_
#define ALENS_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* same fields as SimToolbox/Sylinder/SylinderNear.hpp:40-53 (SylinderNearEP), 104 bytes */
typedef struct {
    int gid, globalIndex, rank, pad_;
    double radius, length, radiusCollision, lengthCollision, colBuf;
    double pos[3];
    double direction[3];
} orc_rod;

/* binary layout of SimToolbox/Constraint/ConstraintBlock.hpp:30-48, 272 bytes */
typedef struct {
    double delta0, gamma, gammaLB;
    int gidI, gidJ, globalIndexI, globalIndexJ;
    unsigned char oneSide, bilateral, pad_[6];
    double kappa;
    double normI[3], normJ[3], posI[3], posJ[3], labI[3], labJ[3];
    double stress[9];
} orc_block;

typedef struct {
    int n;        /* rows */
    long long nnz;
    long long *rowptr;
    int *col;
    double *val;
} orc_csr;

/* history record of BCQPSolver (BCQPSolver.hpp:23): {ite, 0, 0, alpha|tk, resPhi, mvCount} */
typedef struct {
    double v[6];
} orc_hist;

/* SimToolbox/Boundary/Boundary.hpp:24-94: type 0 SphereShell(center, radius, inside), 1 Wall(center, norm),
 * 2 Tube(center, axis, radius, inside); `axis` holds the wall normal / tube axis (normalised by the constructor) */
typedef struct {
    int type, inside;
    double center[3], axis[3], radius;
} orc_boundary;

int orc_sizeof_rod(void);
/* Boundary::project (Boundary.cpp:25-41, :108-124, :185-209) */
void orc_boundary_project(const orc_boundary *b, const double query[3], double project[3], double delta[3]);
/* SylinderSystem::collectBoundaryCollision (SylinderSystem.cpp:1093-1150): one-sided blocks of the rods' end points
 * (centre for spheres) against every boundary, order = (boundary, rod, minus end, plus end); returns the count */
/* SylinderSystem::collectLinkBilateral (SylinderSystem.cpp:1386-1482): one bilateral block per link prev -> next between
 * the PLUS end of prev and the MINUS end of next (true length / radius, nearest periodic image of next, findPBCImage of
 * Util/GeoUtil.hpp:27-60), delta0 = |Q - P| - rI - rJ - linkGap, kappa = linkKappa, stress by collideStress.  Blocks in
 * link order; returns the count, -1 if a gid is unknown. */
long long orc_collect_links(int n, const orc_rod *rods, long long nLinks, const int *prevGid, const int *nextGid,
                            const double boxLow[3], const double boxHigh[3], const int pbc[3], double linkKappa,
                            double linkGap, orc_block *out);
long long orc_collect_boundary(int n, const orc_rod *rods, int nb, const orc_boundary *bnd, double colBuf, orc_block *out,
                               long long cap);
int orc_sizeof_block(void);

/* ---- geometry (Collision/DCPQuery.hpp) */
double orc_dcp_segseg(const double P0[3], const double P1[3], const double Q0[3], const double Q1[3], double Ploc[3],
                      double Qloc[3], double *s, double *t);
double orc_dist_point_seg(const double pt[3], const double minus[3], const double plus[3], double perp[3]);

/* ---- rod preparation (SylinderNear.hpp:74-90, SylinderSystem.cpp:897-905) */
void orc_quat_to_dir(const double quat_xyzw[4], double dir[3]);
void orc_make_rods(int n, const int *gid, const double *radius, const double *length, const double *pos,
                   const double *quat_xyzw, double diameterColRatio, double lengthColRatio, double colBuf,
                   int globalIndexBase, orc_rod *out);
void orc_wrap_positions(int n, double *pos, const double boxLow[3], const double boxHigh[3], const int *pbc);

/* ---- pair functor (SylinderNear.hpp:197-519) */
int orc_pair_functor(const orc_rod *a, const orc_rod *b, int withStress, orc_block *out);
void orc_collide_stress(const double dirI[3], const double dirJ[3], const double centerI[3], const double centerJ[3],
                        double hI, double hJ, double rI, double rJ, double rho, const double Ploc[3],
                        const double Qloc[3], double stress[9]);

/* ---- geometric pair list P_geo (SURVEY.md 8c contract): every (gidI<gidJ, image) with sep < max(colBuf)
 * returns the count; blocks sorted by (gidI,gidJ).  cap = capacity of out[] (count is returned even if > cap) */
long long orc_collect_pairs_brute(int n, const orc_rod *rods, const double boxLow[3], const double boxHigh[3],
                                  const int pbc[3], int withStress, orc_block *out, long long cap);
long long orc_collect_pairs_cells(int n, const orc_rod *rods, const double boxLow[3], const double boxHigh[3],
                                  const int pbc[3], int withStress, orc_block *out, long long cap, int nthreads);

/* ---- assembly (ConstraintCollector.cpp:237-423, SylinderSystem.cpp:622-717, Sylinder.cpp:69-82) */
void orc_drag_coeff(double radius, double length, double viscosity, double *dragPara, double *dragPerp,
                    double *dragRot);
/* DT: nc rows, 12 (6 if oneSide) nnz/row, columns 6*globalIndex+c; caller frees with orc_csr_free */
int orc_build_dtrans(long long nc, const orc_block *blocks, int nRodsGlobal, orc_csr *DT, double *delta0,
                     double *invKappa, double *biFlag, double *gammaGuess);
int orc_transpose(const orc_csr *A, int ncols, orc_csr *AT);
int orc_build_mobility(int n, const orc_rod *rods, const int *immovable, double viscosity, orc_csr *M);
void orc_csr_free(orc_csr *A);
void orc_spmv(const orc_csr *A, const double *x, double *y, double alpha, double beta, int nthreads);

/* ---- the whole solve (ConstraintSolver.cpp:4-107, ConstraintOperator.cpp:30-71, BCQPSolver.cpp:134-497) */
typedef struct {
    /* inputs */
    int nRods;
    long long nc;
    double dt, res;
    int maxIte, solverChoice, nthreads;
    /* outputs */
    int nIte, mvCount, status;
    double resFinal;
    double tAssemble, tSolve; /* seconds */
} orc_solve_info;

/* gamma[nc] (out), forceU/velU/forceB/velB [6*nRods] (out), hist[histCap] (out, *nHist records written) */
int orc_solve_constraints(const orc_block *blocks, const orc_rod *rods, const int *immovable, double viscosity,
                          const double *velNonCon, orc_solve_info *info, double *gamma, double *forceU, double *velU,
                          double *forceB, double *velB, orc_hist *hist, int histCap, int *nHist);

/* raw operator y = (D^T M D + K^-1/dt) x with the same explicit-matrix structure, for operator parity tests */
int orc_operator_apply(const orc_block *blocks, long long nc, const orc_rod *rods, const int *immovable, int nRods,
                       double viscosity, double dt, const double *x, double *y, double *force, double *vel);

/* generic BCQP on an explicit CSR A (for the scipy cross-check, BCQPSolver_verify.py style) */
int orc_bcqp_csr(const orc_csr *A, const double *b, const double *lb, const double *ub, double *x, double tol,
                 int maxIte, int solverChoice, orc_hist *hist, int histCap, int *nHist);

/* write-back (ConstraintCollector.cpp:439-461): gamma -> blocks, stress *= gamma */
void orc_writeback_gamma(long long nc, orc_block *blocks, const double *gamma);

#ifdef __cplusplus
}
#endif
#endif
