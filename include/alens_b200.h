/*
 * alens_b200.h -- C ABI of libalens_b200.so: the B200-native (sm_100a) implementation of the
 * aLENS / SimToolbox per-timestep collision-constraint path.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain pointers and sizes
 * (no torch / Tpetra / Eigen types) and replaces one piece of the reference's CPU path; the
 * reference-side citation (relative to the aLENS tree) is given at each declaration.  The C++ classes
 * in include/alens_b200/ (SylinderSystem.hpp, / ConstraintSolver / ConstraintCollector / BCQPSolver
 * with the reference's method names) are thin forwarding wrappers over these functions.
 *
 * Conventions
 *   - every function returns ALENS_OK (0) or a negative error code; alens_last_error() gives the text.
 *   - host pointers are caller-owned; the library owns all device memory.
 *   - calls are synchronous at the boundary (they return after the work on the context's stream
 *     finished) unless the name ends in _async.
 *   - rods are identified by their LOCAL index (position in the arrays given to alens_set_rods);
 *     ConstraintBlock::globalIndexI/J = rank offset + local index as in
 *     SimToolbox/Sylinder/SylinderSystem.cpp:868-880 (updateSylinderMap).
 *   - there is NO CPU fallback: without a CUDA device alens_create fails.
 */
#ifndef ALENS_B200_H_
#define ALENS_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALENS_OK 0
#define ALENS_ERR_CUDA (-1)      /* a CUDA runtime call failed                      */
#define ALENS_ERR_ARG (-2)       /* bad argument / call order                       */
#define ALENS_ERR_STATE (-3)     /* required earlier stage has not been run         */
#define ALENS_ERR_UNSUPPORTED (-4)
#define ALENS_ERR_PROJECTION (-5) /* BCQPSolver.cpp:484-494 "projection error"      */
#define ALENS_ERR_COMM (-6)

#define ALENS_SOLVER_BBPGD 0 /* SylinderConfig conSolverChoice, ConstraintSolver.cpp:74-84 */
#define ALENS_SOLVER_APGD 1

typedef struct alens_ctx alens_ctx;

/* Binary-compatible with SimToolbox/Constraint/ConstraintBlock.hpp:30-48 (272 bytes, x86-64). */
typedef struct alens_constraint_block {
    double delta0, gamma, gammaLB;
    int gidI, gidJ, globalIndexI, globalIndexJ;
    unsigned char oneSide, bilateral, pad_[6];
    double kappa;
    double normI[3], normJ[3], posI[3], posJ[3], labI[3], labJ[3];
    double stress[9];
} alens_constraint_block;

/* Outcome of one constraint solve; history rows follow BCQPSolver.hpp:23 (IteHistory). */
typedef struct alens_solve_report {
    int status;       /* 0 ok, 1 stagnated (BCQPSolver.cpp:229-233 / :332-336)           */
    int iterations;   /* iteCount                                                          */
    int matvecs;      /* mvCount                                                           */
    int history_rows; /* rows available through alens_get_history                          */
    double residual;  /* last resPhi (velocity units, i.e. before the *dt of the log line) */
    double step;      /* last alpha (BBPGD) or tk (APGD)                                   */
    long long n_constraints;
    int n_rods;
} alens_solve_report;

/* per-phase device time of the last step, milliseconds (CUDA events on the context's stream);
 * names follow the reference's Teuchos timers (SylinderSystem.cpp:831-851, ConstraintOperator.cpp:8-11) */
typedef struct alens_timers {
    double upload_ms;         /* alens_set_rods H2D + rod_pack                                  */
    double collect_ms;        /* SylinderSystem::CollectCollision                               */
    double setup_ms;          /* ConstraintSolver::setup equivalent (incidence, q, bounds)      */
    double solve_ms;          /* SylinderSystem::SolveConstraints (BCQP loop only)              */
    double split_ms;          /* uni/bi split + result permutation                              */
    double download_ms;       /* D2H of results                                                 */
    double op_force_vel_ms;   /* ConstraintOperator::ApplyDMat + ApplyMobility (accumulated)    */
    double op_dtrans_ms;      /* ConstraintOperator::ApplyDMatTrans + fused vector work         */
    double op_update_ms;      /* x = P(x - alpha g) passes                                      */
    long long op_force_vel_n, op_dtrans_n, op_update_n; /* launches behind the three sums       */
    long long op_launches;    /* kernel launches inside the last BCQP loop                      */
    long long total_launches; /* kernel launches since alens_reset_timers                       */
    long long op_rows_live;   /* BBPGD: constraint rows that could be non-zero, summed over the
                                 operator applies of the last solve (the force kernel reads only those) */
    long long op_applies;     /* ... and the number of applies behind that sum                   */
} alens_timers;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* device: CUDA ordinal.  rank/nranks: position in the slab decomposition (1 process per GPU).
 * Replaces SylinderSystem::initialize's solver/collector construction (SylinderSystem.cpp:53-55). */
int alens_create(int device, int rank, int nranks, alens_ctx **out);
void alens_destroy(alens_ctx *ctx);
const char *alens_last_error(const alens_ctx *ctx);
const char *alens_version(void);
/* use an externally created cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream */
int alens_set_stream(alens_ctx *ctx, void *cuda_stream);

/* ---- configuration -------------------------------------------------------------------------- */
/* simBoxLow/High/PBC: SylinderSystem::setDomainInfo (SylinderSystem.cpp:569-610) */
int alens_set_domain(alens_ctx *ctx, const double boxLow[3], const double boxHigh[3], const int pbc[3]);
/* sylinderDiameterColRatio / sylinderLengthColRatio / sylinderColBuf: SylinderSystem.cpp:897-905 */
int alens_set_collision_params(alens_ctx *ctx, double diameterColRatio, double lengthColRatio, double colBuf);

/* ---- rods ----------------------------------------------------------------------------------- */
/* Host SoA of the Sylinder fields the path reads (SimToolbox/Sylinder/Sylinder.hpp:38-84):
 * gid[n], pos[3n], orientation[4n] (Eigen coeff order x,y,z,w), length[n], radius[n],
 * immovable[n] (may be NULL = all movable).  wrapIntoBox != 0 applies applyBoxBC
 * (FDPS/particle_system.hpp:798-843: all three axes).  Also performs prepareStep's per-rod work
 * (collision radius/length, direction = q*z; SylinderNear.hpp:74-90). */
int alens_set_rods(alens_ctx *ctx, int n, const int *gid, const double *pos, const double *orientation,
                   const double *length, const double *radius, const unsigned char *immovable, int wrapIntoBox);
/* same from an array of reference `Sylinder` records (568-byte AoS), stride in bytes */
int alens_set_rods_aos(alens_ctx *ctx, int n, const void *sylinders, size_t stride, int wrapIntoBox);
/* The per-step form of alens_set_rods for a rod set whose gids, lengths, radii and immovable flags did not change since the
 * last alens_set_rods: only position and orientation (56 of 77 bytes per rod) cross the bus, the rest of prepareStep
 * (applyBoxBC, ghost exchange, cell list) runs as usual.  ALENS_ERR_STATE without resident rods. */
int alens_set_rod_state(alens_ctx *ctx, const double *pos, const double *orientation, int wrapIntoBox);
/* SylinderSystem::prepareStep on the rods already resident on the device (after alens_step_euler or a
 * previous alens_set_rods): box wrap, cell list, sorted SoA -- no host traffic. */
int alens_prepare_step(alens_ctx *ctx, int wrapIntoBox);
/* wrapped positions back (what applyBoxBC left in the container), pos[3n] */
int alens_get_positions(alens_ctx *ctx, double *pos);

/* ---- constraint collection ------------------------------------------------------------------ */
/* SylinderSystem::collectPairCollision (SylinderSystem.cpp:1152-1160) = FDPS calcForceAll +
 * CalcSylinderNearForce (SylinderNear.hpp:197-414).  Clears the pool first (conCollectorPtr->clear(),
 * SylinderSystem.cpp:925).  The list is the full geometric list (SURVEY.md 8c contract), in a
 * deterministic order. */
int alens_collect_pair_collision(alens_ctx *ctx, long long *nConstraints);
/* host-generated blocks pushed into the pool (boundary / link / protein bilateral constraints:
 * SylinderSystem.cpp:1093-1150, :1386-1482, SRC/TubuleSystem.cpp:694-745).  Two-sided blocks must
 * have normJ == -normI (true for every producer in the reference). */
int alens_append_constraints(alens_ctx *ctx, const alens_constraint_block *blocks, long long n);
/* One boundary of RunConfig::boundaryPtr (SimToolbox/Boundary/Boundary.hpp:24-94).  type: 0 SphereShell(center,
 * radius, inside), 1 Wall(center, norm), 2 Tube(center, axis, radius, inside); `axis` = wall normal / tube axis and is
 * normalised here as the reference's constructors do. */
typedef struct alens_boundary {
    int type, inside;
    double center[3], axis[3], radius;
} alens_boundary;
/* SylinderSystem::collectBoundaryCollision (SylinderSystem.cpp:1093-1150) on the device: every rod end point (the centre
 * of a sphere) is projected onto every boundary; a point outside, or inside within (1 + 2 colBuf) radiusCollision, adds a
 * one-sided block (delta0 = -+|delta| - radius, normI = delta/|delta|, posI = Q - centre, labI = Q, labJ = projection).
 * Order: (boundary, rod in the caller's order, minus end, plus end).  Call after alens_collect_pair_collision. */
int alens_collect_boundary_collision(alens_ctx *ctx, const alens_boundary *boundaries, int nBoundaries, long long *nAdded);
/* SylinderSystem::collectLinkBilateral (SylinderSystem.cpp:1386-1482) on the device: for every link prev -> next (gids,
 * the reference's linkMap) one bilateral block between the plus end of `prev` and the minus end of `next` (true length and
 * radius, nearest periodic image), delta0 = distance - rI - rJ - linkGap, kappa = linkKappa, stress by collideStress.
 * Blocks are appended in link order; the gid -> rod lookup (the reference's ZDD directory) is a device hash table.
 * One rank: both gids must exist.  Slab decomposition: hand EVERY rank the whole link map (as every rank of the reference
 * reads it); a rank builds the blocks of the links it owns at least one rod of -- the partner may be a ghost, both owners of a
 * cross-slab link build the same block -- and skips the others; *nAdded counts the blocks this rank built.  An owned rod whose
 * partner lies outside the ghost layer (cutoff + skin) is ALENS_ERR_ARG. */
int alens_collect_link_bilateral(alens_ctx *ctx, const int *prevGid, const int *nextGid, long long nLinks, double linkKappa,
                                 double linkGap, long long *nAdded);
/* The fields of ProteinData / ProteinBindStatus that TubuleSystem::setProteinConstraints (SRC/TubuleSystem.cpp:694-745)
 * reads for one protein (Protein/ProteinBindStatus.hpp: idBind, indexBind, centerBind, directionBind, posEndBind, lenBind;
 * ProteinData::getProteinForceLength(), property.freeLength, property.kappa).  idBind < 0 = that end is unbound (ID_UB). */
typedef struct alens_protein_bind {
    int idBind[2], indexBind[2];
    double centerBind[2][3], directionBind[2][3], posEndBind[2][3], lenBind[2];
    double forceLength, freeLength, kappa;
} alens_protein_bind;
/* TubuleSystem::setProteinConstraints on the device: one bilateral block per DOUBLY bound protein (delta0 = forceLength -
 * freeLength, gamma0 = -delta0 kappa, normI = (P - Q)/|P - Q|, posI/J relative to the rod centres, stress by collideStress
 * with radius tubuleDiameter / 2), appended in protein order; indexBind = globalIndex of the rods (this rank's). */
int alens_collect_protein_bilateral(alens_ctx *ctx, const alens_protein_bind *proteins, long long n, double tubuleDiameter,
                                    long long *nAdded);
int alens_clear_constraints(alens_ctx *ctx); /* ConstraintCollector::clear */
/* ConstraintCollector::getLocalNumberOfConstraints (ConstraintCollector.cpp:30-36) */
int alens_num_constraints(alens_ctx *ctx, long long *n);
/* refill a host pool: blocks in solver order.  withStress: evaluate collideStress
 * (SylinderNear.hpp:432-484).  writeBack: gamma = solved value and stress *= gamma
 * (ConstraintCollector::writeBackGamma, ConstraintCollector.cpp:439-461). */
int alens_get_constraints(alens_ctx *ctx, alens_constraint_block *out, long long cap, int withStress,
                          int writeBack);

/* ---- two-species neighbour search ------------------------------------------------------------- */
/* MixPairInteraction<FPT, FPS, EPT, EPS, Force>::computeForce (SimToolbox/MPI/MixPairInteraction.hpp:148-311; the protein ->
 * rod search of SRC/TubuleSystem.cpp) on the rods' cell list: for every TARGET point t (host arrays targetPos[3n],
 * targetRSearch[n]) the resident rods j, and their periodic images, with |x_t - x_j| <= max(rs_t, rs_j) -- the distance
 * FDPS' Symmetry search guarantees; the caller's functor makes the fine decision as in the reference.  sourceRSearch:
 * per local rod, or NULL = SylinderNearEP::getRSearch (SylinderNear.hpp:108-113).  Result in CSR form: rowPtr[nTargets + 1]
 * and sourceIndex[rowPtr[nTargets]] (local rod indices).  If capPairs is too small only rowPtr and *nPairs are filled: call
 * again with a larger buffer.  Call after alens_set_rods / alens_prepare_step. */
int alens_mix_pair_search(alens_ctx *ctx, long long nTargets, const double *targetPos, const double *targetRSearch,
                          const double *sourceRSearch, long long *rowPtr, int *sourceIndex, long long capPairs,
                          long long *nPairs);

/* ---- the narrow phase by itself ---------------------------------------------------------------- */
/* DCPQuery<3,double,Evec3>::operator() (SimToolbox/Collision/DCPQuery.hpp:199-308), n independent segment pairs
 * [P0,P1] x [Q0,Q1] (3n host doubles each): minimal distance and the closest points (any output may be NULL).
 * Bit-identical to the reference header, including its behaviour on degenerate input (zero-length and parallel
 * segments, the 0.5 fall-backs of :325-327,361-363,434-438). */
int alens_dcp_query(alens_ctx *ctx, long long n, const double *P0, const double *P1, const double *Q0, const double *Q1,
                    double *dist, double *Ploc, double *Qloc);
/* CalcSylinderNearForce's per-pair body (SylinderNear.hpp:241-414: sp_sp / sp_sy / sy_sy by lengthCollision <
 * 2 radiusCollision) + collideStress (:432-519) for n independent (target I, source J) pairs, no gid filter and no
 * neighbour search.  geomI/geomJ: 9 host doubles per rod {pos[3], direction[3], lengthCollision, radiusCollision,
 * colBuf} (the SylinderNearEP fields the functor reads).  hit[k] = collision found; blocks[k] = the ConstraintBlock
 * the functor would push (gid/globalIndex = 2k, 2k+1; zeroed when there is no hit). */
int alens_pair_functor(alens_ctx *ctx, long long n, const double *geomI, const double *geomJ, int withStress,
                       unsigned char *hit, alens_constraint_block *blocks);

/* ---- mobility ------------------------------------------------------------------------------- */
/* SylinderSystem::calcMobOperator / calcMobMatrix (SylinderSystem.cpp:622-722) with
 * Sylinder::calcDragCoeff (Sylinder.cpp:69-82): stored as 3 inverse drag scalars per rod. */
int alens_calc_mobility(alens_ctx *ctx, double viscosity);
/* y = M x on 6n vectors in local rod order (mobilityOperatorRcp->apply, SylinderSystem.cpp:734) */
int alens_mobility_apply(alens_ctx *ctx, const double *x, double *y);

/* ---- solve ---------------------------------------------------------------------------------- */
/* ConstraintSolver::setup + setControlParams + solveConstraints (ConstraintSolver.cpp:4-107,
 * driven from SylinderSystem::resolveConstraints, SylinderSystem.cpp:829-866) with
 * BCQPSolver::solveBBPGD / solveAPGD (BCQPSolver.cpp:134-389).  velNonCon: 6n host doubles in local
 * rod order (NULL = the resident vector, zero by default).  res is conResTol (the loop stops at resPhi < res/dt). */
int alens_solve_constraints(alens_ctx *ctx, const double *velNonCon, double dt, double res, int maxIte,
                            int solverChoice, alens_solve_report *report);
/* velocityNonCon of SylinderSystem::calcVelocityNonCon (SylinderSystem.cpp:724-800) made resident on the
 * device: 6n host doubles in local rod order, NULL = zero.  alens_solve_constraints / alens_setup_constraints
 * called with velNonCon == NULL use this resident vector. */
int alens_set_velocity_noncon(alens_ctx *ctx, const double *velNonCon);
/* SylinderSystem::calcVelocityNonCon (SylinderSystem.cpp:724-800) in one device pass, result resident:
 *   velNonCon = M forceNonBrown + velocityNonBrown + velocityBrown
 * (each term optional: NULL = absent; host arrays of 6n doubles in local rod order; needs alens_calc_mobility).
 * monolayer != 0 zeroes v_z, omega_x, omega_y of every term as the reference does.  velNonBOut (optional, 6n host doubles)
 * receives M forceNonBrown + velocityNonBrown, what the reference writes to Sylinder::velNonB / omegaNonB. */
/* SylinderSystem::calcVelocityBrown (SylinderSystem.cpp:1020-1091): random finite difference scheme of Delong et al. per
 * rod -- N = (1/zPara - 1/zPerp) q q^T + 1/zPerp I, v = sqrt(2 kBT/dt) chol(N) W_pos + (kBT/delta) (N_rfd - N) W_rfdpos with
 * delta = 0.1 dt and N_rfd at the orientation rotated by W_rfdrot * delta, omega = sqrt(1/zRot) sqrt(2 kBT/dt) W_rot.
 * normals12: 12 standard normal deviates per local rod in the reference's draw order (W_rot, W_pos, W_rfdrot, W_rfdpos;
 * hand in the host application's own stream to reproduce its run), or NULL: drawn on the device from a counter-based
 * Philox4x32-10 generator keyed by (seed, step, rod gid), independent of rod order and of the number of GPUs.
 * velBrownOut: 6n host doubles (velBrown, omegaBrown per rod); pass it to alens_calc_velocity_noncon.
 * Needs alens_calc_mobility (the viscosity given there). */
int alens_calc_velocity_brown(alens_ctx *ctx, double kBT, double dt, const double *normals12, unsigned long long seed,
                              unsigned long long step, double *velBrownOut);
int alens_calc_velocity_noncon(alens_ctx *ctx, const double *forceNonBrown, const double *velocityNonBrown,
                               const double *velocityBrown, int monolayer, double *velNonBOut);
/* Same, without waiting for the copy: it runs on a side stream and overlaps whatever the caller does next
 * (typically alens_collect_pair_collision); the next call that reads the vector waits for it on the device.
 * velNonCon must be page-locked host memory and stay unchanged until alens_solve_constraints /
 * alens_setup_constraints has returned. */
int alens_set_velocity_noncon_async(alens_ctx *ctx, const double *velNonCon);
/* only the setup part (q, bounds, incidence); lets tests call alens_operator_apply */
int alens_setup_constraints(alens_ctx *ctx, const double *velNonCon, double dt);
/* ConstraintOperator::apply (ConstraintOperator.cpp:30-71): y = (D^T M D + K^-1/dt) x, host vectors of
 * length nConstraints; force/vel (6n, may be NULL) = the cached D x and M D x. */
int alens_operator_apply(alens_ctx *ctx, const double *x, double *y, double *force, double *vel);
/* BCQPSolver::solveBBPGD / solveAPGD (BCQPSolver.hpp:75-97) on the operator built by the last setup:
 * min 1/2 x^T A x + b^T x with the bounds ConstraintSolver sets (0 / -inf by the bilateral flag).
 * b: host vector of length nConstraints or NULL (= q of the setup); x: in = initial guess, out = returned
 * iterate; tol is absolute (the caller has already divided by dt).  The force/velocity results of
 * alens_get_force_velocity are refreshed as in ConstraintSolver::solveConstraints. */
int alens_bcqp_solve(alens_ctx *ctx, const double *b, double *x, double tol, int maxIte, int solverChoice,
                     alens_solve_report *report);
/* BCQPSolver as ANY caller sees it (BCQPSolver.hpp:37-111): A = a CSR matrix of the caller (the reference's TCMAT; e.g. its
 * own random self-test problem BCQPSolver(int, double), BCQPSolver.cpp:38-132) or the constraint operator of the last
 * alens_setup_constraints; b and the bounds are the caller's (setLowerBound / setUpperBound; NULL = the default
 * -+DBL_MAX/10 of setDefaultBounds, BCQPSolver.cpp:499-510).  The loops of solveBBPGD / solveAPGD run as device vector
 * kernels, one per Tpetra call of the reference, scalar control on the host; x: in = initial guess, out = result
 * (after an iteMax exit of BBPGD the older iterate, as in the reference).  A projection error (BCQPSolver.cpp:484-494)
 * returns ALENS_ERR_PROJECTION. */
typedef struct alens_bcqp alens_bcqp;
int alens_bcqp_create_csr(alens_ctx *ctx, int n, const long long *rowPtr, const int *colInd, const double *values,
                          const double *b, alens_bcqp **out);
/* b == NULL: q of the setup (delta0/dt + D^T velNonCon) */
int alens_bcqp_create_constraint(alens_ctx *ctx, const double *b, alens_bcqp **out);
int alens_bcqp_set_lower_bound(alens_bcqp *p, const double *lb);
int alens_bcqp_set_upper_bound(alens_bcqp *p, const double *ub);
int alens_bcqp_get_bounds(alens_bcqp *p, double *lb, double *ub);
int alens_bcqp_run(alens_bcqp *p, double *x, double tol, int maxIte, int solverChoice, alens_solve_report *report);
int alens_bcqp_history(alens_bcqp *p, double *rows6, int capRows, int *nRows);
int alens_bcqp_size(alens_bcqp *p);
void alens_bcqp_destroy(alens_bcqp *p);
/* IteHistory rows {ite,0,0,alpha,resPhi,mvCount} (BCQPSolver.hpp:23) */
int alens_get_history(alens_ctx *ctx, double *rows6, int capRows, int *nRows);
/* gamma in solver (= alens_get_constraints) order */
int alens_get_gamma(alens_ctx *ctx, double *gamma, long long cap);
/* ConstraintSolver::getForceUni/getVelocityUni/getForceBi/getVelocityBi (ConstraintSolver.hpp:83-88),
 * 6n each in local rod order; any pointer may be NULL.  This is what
 * SylinderSystem::saveForceVelocityConstraints (SylinderSystem.cpp:971-1018) copies into the rods. */
int alens_get_force_velocity(alens_ctx *ctx, double *forceUni, double *velUni, double *forceBi, double *velBi);

/* ---- next rows: device-resident time stepping --------------------------------- */
/* sumForceVelocity + stepEuler (SylinderSystem.cpp:802-827, Sylinder.cpp:91-99, EquatnHelper.hpp:74-90)
 * on the device copy of the rods: vel = velNonCon + velUni + velBi; pos += vel*dt; q rotated by omega*dt */
int alens_step_euler(alens_ctx *ctx, double dt);
int alens_get_rod_state(alens_ctx *ctx, double *pos, double *orientation);

/* ---- instrumentation ---------------------------------------------------------------------------- */
int alens_get_timers(alens_ctx *ctx, alens_timers *t);
/* on != 0: bracket every kernel of the BCQP loop with CUDA events and accumulate per-kernel device time
 * into alens_timers.op_*_ms (the reference's ConstraintOperator::* Teuchos timers) */
int alens_set_profiling(alens_ctx *ctx, int on);
int alens_reset_timers(alens_ctx *ctx);
/* tuning knobs, none of which changes a result bit (INTEGRATION.md 4b lists them with their defaults): "force_kernel"
 * 1 sparse k_force_vel_act / 0 dense k_force_vel_lm / 2 k_slot_x + k_rod_sum for the operator's D x + M step;
 * "find_split" 1/0 register-split / single-kernel pair search; "pdl", "poll", "lookahead" launch structure of the BBPGD
 * loop; "tail_ring" TMA-staged tail; "tail_ctas_per_sm" persistent grid of k_bb_tail ("bbpgd_batch" = iterations
 * between two host checks when "poll" is 0) */
int alens_set_option(alens_ctx *ctx, const char *name, long long value);
/* average device time (CUDA events) of `reps` back-to-back launches of one BBPGD kernel -- "force_vel", "tail",
 * "force_vel_plain" on the operator of the last alens_setup_constraints (invalidates the setup), or "force_vel_last" on
 * the iterate / mask the last BBPGD solve left behind. */
int alens_time_kernel(alens_ctx *ctx, const char *which, int reps, double *avgMicroseconds);
/* alens_set_option("stamps", 1): %globaltimer stamps (ns) of every BBPGD iteration of the last solve, 8 words each:
 * [0] force kernel: first CTA past its dependency wait, [1] halo flags released to the neighbours, [2] last CTA done;
 * [3] tail kernel: first CTA past its dependency wait, [4] first CTA reached the halo wait, [5] last CTA saw the halo,
 * [6] all rows done (last CTA elected), [7] allreduce over the ranks complete and scalar step taken.  Row 0 = iteration 0.
 * (Multi-GPU: skew between ranks shows up as [5]-[4] and [7]-[6]; the clocks of different GPUs are not aligned.) */
int alens_get_stamps(alens_ctx *ctx, unsigned long long *stamps8, int capIterations, int *nIterations);
/* what the pool holds: pair-collision blocks, one-sided blocks, bilateral blocks.  nBilateral == 0 means forceBi and
 * velocityBi of the next solve are identically zero: a caller need not fetch them (alens_get_force_velocity with NULL) */
int alens_get_pool_stats(alens_ctx *ctx, long long *nCollision, long long *nOneSide, long long *nBilateral);
/* force_kernel 3: incidence slots whose bit is set in the slot bitmap, and rods with at least one such slot, as the last
 * solve left them (what k_force_vel_rec reads per launch: 64 bytes per live slot, mobility data per live rod) */
int alens_get_live_stats(alens_ctx *ctx, long long *liveSlots, long long *liveRods);
/* number of rods / cells / candidate pairs that passed the broad phase in the last collection */
int alens_get_collect_stats(alens_ctx *ctx, long long *nCells, long long *nCandidates, long long *nHits);
/* Polydisperse rods: the cell edge follows shortRadius = min(max, long_rods x mean) of the rods' bounding radii
 * (lengthCollision / 2 + radiusCollision; option "long_rods" in percent, default 200, 0 = off); rods above it are paired
 * with partners beyond the 27-cell stencil by a separate exact pass whose rows follow the stencil rows.  Statistics of the
 * last alens_collect_pair_collision: long rods, rows they added, the two radii. */
int alens_get_long_rod_stats(alens_ctx *ctx, long long *nLongRods, long long *nLongRows, double *shortRadius, double *maxRadius);

/* Order-independent digest of the pool, computed on the device: counts3 = {rows, sum of hash(gidI, gidJ, labJ bits, delta0
 * bits), sum of hash(gidI, gidJ, labJ bits, gamma bits)} (64-bit wrap-around sums), sums3 = {sum gamma, sum gamma^2, sum
 * w gamma} with a weight w in [0,1) derived from the row's identity; gamma = the last solve's result (zeros before).  With a
 * slab decomposition each rank counts the rows whose rod I it owns, so the per-rank digests ADD UP to the digest of the
 * single-rank list (reference canonicalisation: Sylinder/Test2_MixLink/Verify.py:82-83 sorts by the same keys). */
int alens_constraint_digest(alens_ctx *ctx, unsigned long long counts3[3], double sums3[3]);
/* ConstraintCollector::sumLocalConstraintStress (Constraint/ConstraintCollector.cpp:38-74) after writeBackGamma, the sum
 * SylinderSystem::calcConStress takes every step (SylinderSystem.cpp:1226-1263), without bringing the blocks to the host:
 * row-major 3x3 sums of gamma * (unit stress) over the unilateral and the bilateral blocks of the last solve; one-sided
 * blocks are skipped unless withOneSide.  Collision blocks are evaluated (collideStress, SylinderNear.hpp:432-519) and
 * reduced on the device in a fixed order; appended blocks contribute the stress they were appended with.  LOCAL sums:
 * with a slab decomposition a rank counts the blocks whose rod I it owns, and the caller adds the ranks' results as the
 * reference does (Teuchos::reduceAll, :1252-1253).  ALENS_ERR_STATE before a solve. */
int alens_sum_constraint_stress(alens_ctx *ctx, int withOneSide, double uniStress[9], double biStress[9]);

/* ---- multi-GPU (one rank per GPU) ------------------------------------------------------
 * Slab decomposition along one box axis.  Replaces, for the constraint path, the FDPS ghost exchange inside
 * TreeSylinderNear::calcForceAll (FDPS/tree_for_force.hpp:759-842), Tpetra's ghost-column Import
 * (ConstraintOperator.cpp:44-57) and the allreduces of BCQPSolver.cpp:183-233.  Transport: device memory windows
 * mapped between the ranks (NVLink peer access; cudaIpc between processes), written with remote stores and
 * sequence numbers -- see alens_b200/csrc/comm.cu.
 *
 * Every rank owns the rods whose centre lies in [slabLow, slabHigh) along `axis` (a rod may stray up to `skin`
 * outside before the host has to redistribute: alens_prepare_step / alens_set_rods then return ALENS_ERR_STATE).
 * maxBoundingRadius = max over ALL ranks of lengthCollision/2 + radiusCollision.  globalIndexBase = exclusive scan
 * of the local rod counts over the ranks (SylinderSystem::updateSylinderMap, SylinderSystem.cpp:868-880).
 * With a decomposition in place alens_set_rods / alens_prepare_step / alens_solve_constraints are COLLECTIVE
 * (every rank must call them, with the same dt / res / maxIte); results are returned for the owned rods only;
 * constraints that couple rods of two ranks are held (bit-identically) by both. */
int alens_set_decomposition(alens_ctx *ctx, int axis, double slabLow, double slabHigh, double skin,
                            double maxBoundingRadius, int globalIndexBase);
/* allocate this rank's window; maxLocalRods bounds the owned rods per rank (ghost capacity = a third of it) */
int alens_comm_create(alens_ctx *ctx, long long maxLocalRods);
/* multi-process bootstrap: export a blob of alens_comm_blob_size() bytes, all-gather the blobs in rank order with
 * the host program's own means (MPI_Allgather, torch.distributed.all_gather, ...), then connect (collective) */
int alens_comm_blob_size(void);
int alens_comm_export(alens_ctx *ctx, void *blob);
int alens_comm_connect(alens_ctx *ctx, const void *blobsInRankOrder);
/* single-process bootstrap: n contexts (rank order) driven by n host threads; devices may coincide */
int alens_comm_connect_local(alens_ctx **ctxs, int n);
/* ghost rods received / owned rods mirrored on the left and right neighbour in the last exchange */
/* Device-side rod migration, the slab counterpart of decomposeDomain + exchangeSylinder (SylinderSystem.cpp:617-620):
 * COLLECTIVE over the connected ranks, to be called on the resident state between alens_step_euler and
 * alens_prepare_step.  Owned rods whose wrapped centre left this rank's slab move to the neighbouring slab through the
 * peer-memory channels (record: gid, position, orientation, length, radius, immovable flag, velNonCon row); the rods that
 * stay keep their order, arrivals follow (left neighbour's first); every rank's globalIndex base is renumbered
 * (updateSylinderMap, :868-880).  ALENS_ERR_STATE if a rod moved further than the neighbouring slab (the host has to
 * redistribute then).  The caller learns its new local set with alens_get_rod_identity + alens_get_rod_state. */
int alens_migrate_rods(alens_ctx *ctx, long long *nSent, long long *nReceived);
/* One opaque 64-bit tag per owned rod (local order), kept on the device and carried along by alens_migrate_rods -- what
 * FDPS's exchangeParticle does for the rest of the Sylinder record.  alens_set_rods_aos stores Sylinder::group there.
 * alens_set_rods clears the tags; NULL clears them too. */
int alens_set_rod_tags(alens_ctx *ctx, const long long *tags);
int alens_get_rod_tags(alens_ctx *ctx, long long *tags);
/* the resident owned rods in the context's order; any output may be NULL (call once with NULL arrays for the count) */
int alens_get_rod_identity(alens_ctx *ctx, int *nLocal, int *globalIndexBase, int *gid, double *length, double *radius,
                           unsigned char *immovable);
int alens_num_ghosts(alens_ctx *ctx, int *nGhost, int *nSentLeft, int *nSentRight);
/* connected: the windows are mapped; fused: every rank sits on its own GPU, so the BBPGD kernels wait for their
 * neighbours themselves (halo wait + mailbox allreduce inside k_bb_tail, remote stores from k_force_vel_act) instead of
 * through the one-thread helper kernels used when ranks share a device */
int alens_comm_mode(alens_ctx *ctx, int *connected, int *fused);

#ifdef __cplusplus
}
#endif
#endif /* ALENS_B200_H_ */
