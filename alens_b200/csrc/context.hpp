// context.hpp -- the state behind alens_ctx: device buffers, configuration, per-phase timers.
// All device memory is owned here (grow-only buffers sized for the largest step seen so far).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/alens_b200.h"

namespace alens {

struct CudaError {
    cudaError_t code;
    const char *what;
    const char *file;
    int line;
};

#define ALENS_CUDA(expr)                                                                                            \
    do {                                                                                                            \
        cudaError_t e__ = (expr);                                                                                   \
        if (e__ != cudaSuccess) throw ::alens::CudaError{e__, #expr, __FILE__, __LINE__};                           \
    } while (0)

struct ArgError {
    int code;
    std::string msg;
};

// Stream all device allocations of the calling thread are ordered on (set at every C-API entry).  Stream-ordered
// allocation (cudaMallocAsync / cudaFreeAsync) never synchronises the whole device -- cudaFree does, which would
// deadlock two ranks that share one GPU (tests) while one of them waits in a kernel for the other.
extern thread_local cudaStream_t g_allocStream;
extern thread_local bool g_allocAsync;

inline void *devAlloc(size_t bytes) {
    void *p = nullptr;
    if (g_allocAsync) ALENS_CUDA(cudaMallocAsync(&p, bytes, g_allocStream));
    else ALENS_CUDA(cudaMalloc(&p, bytes));
    return p;
}
inline void devFree(void *p) {
    if (!p) return;
    if (g_allocAsync) cudaFreeAsync(p, g_allocStream);
    else cudaFree(p);
}

// grow-only device array
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    bool external = false; // memory owned elsewhere (a region of the communication window): fixed capacity
    ~DevBuf() { release(); }
    void release() {
        if (p && !external) devFree(p);
        p = nullptr;
        cap = 0;
        external = false;
    }
    // contents are NOT preserved on growth unless keep=true; everything is ordered on the context's stream
    void reserve(size_t n, cudaStream_t = 0, bool keep = false, size_t keepN = 0) {
        if (n <= cap) return;
        if (external) throw ArgError{ALENS_ERR_COMM, "buffer inside the communication window is too small"};
        size_t ncap = n + n / 4 + 64;
        T *np = static_cast<T *>(devAlloc(ncap * sizeof(T)));
        if (keep && p && keepN)
            ALENS_CUDA(cudaMemcpyAsync(np, p, keepN * sizeof(T), cudaMemcpyDeviceToDevice, g_allocStream));
        devFree(p);
        p = np;
        cap = ncap;
    }
};

// pinned host staging buffer (grow-only)
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    void *reserve(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            cap = bytes + bytes / 4 + 4096;
            ALENS_CUDA(cudaMallocHost(&p, cap));
        }
        return p;
    }
};

struct Box {
    double lo[3], hi[3], len[3];
    int pbc[3];
};

struct CellGrid {
    int n[3];       // cells per axis
    int ncell;      // product
    double inv[3];  // n[k] / extent[k]
    double cutoff;
    double lo[3];   // origin (box low corner; along the slab axis: slab low face - ghost width)
    int per[3];     // neighbour cells wrap around (periodic axis of a box the grid spans completely)
    int axis;       // slab axis (-1: single rank)
    double axisLen; // box length along the slab axis
};

// scalar block shared between the BCQP kernels (lives in device memory, mirrored to pinned host)
struct SolverScalars {
    double alpha;      // current step (BBPGD alpha / APGD tk)
    double res;        // last resPhi
    double dotA, dotB; // last a, b of the BB step
    int ite;           // iteCount
    int mv;            // mvCount
    int done;          // 1 converged, 2 stagnated, 3 projection error
    int nhist;         // history rows written
    unsigned int ticket;   // last-block election counter of k_bb_tail
    unsigned int ticketFv; // ... of the force kernel (fused halo push: the last CTA releases the neighbours' flags)
    unsigned long long maybeAcc;  // rows marked "may be non-zero" by the running k_bb_tail
    unsigned long long maybeRows; // ... by the last completed one
    unsigned long long maybeSum;  // ... summed over the applies of this solve
};

// ---- multi-GPU (comm.cu) ----------------------------------------------------------------------
static constexpr int kMaxRanks = 16;
static constexpr int kGhostRec = 18; // doubles per ghost record

// start of every rank's window; every word below is written by a PEER (remote store) and polled locally
struct CommHeader {
    unsigned long long chanSeq[2]; // ghost payload arrived from my left / right neighbour
    long long chanCount[2];
    unsigned long long ackSeq[2];  // sorted indices of the rods I mirrored on my left / right neighbour arrived
    unsigned long long haloSeq[2]; // ghost rows of U pushed by my left / right neighbour
    unsigned long long vecSeq[2];  // ghost rows of velNonCon pushed
    unsigned long long mailSeq[2][kMaxRanks]; // [parity][source rank]
    double mail[2][kMaxRanks][4];             // BBPGD partial sums {dx.dx, dx.dg, dg.dg, max |q|}
    int error;                                // set by a waiter that timed out
    // rod migration (commMigrate): every rank publishes its new number of owned rods to every rank
    unsigned long long cntSeq[kMaxRanks];
    long long cnt[kMaxRanks];
    // written by my left / right neighbour with its index acks: the row of ITS rod-velocity vector where the velocities of the
    // rods I mirror over there are expected, contiguous and in my send order (tail_push)
    int ghostBase[2];
};

struct CommBlob { // what a rank publishes to its peers (multi-process bootstrap)
    cudaIpcMemHandle_t handle;
    unsigned long long bytes;
    int device, rank;
};

struct GhostSrc {
    const int *gid;
    const double *pos, *quat, *len, *rad;
    const unsigned char *imm;
    const signed char *img;
    const double *velNC;
    int globalBase;
    const long long *tag; // the host's per-rod tag (alens_set_rod_tags); nullptr: none
};
struct GhostDst {
    int *gid;
    double *pos, *quat, *len, *rad;
    unsigned char *imm;
    signed char *img;
    int *globalIdx;
    double *velNC;
    long long *tag; // nullptr: not stored
};

struct Comm {
    bool active = false;
    int left = -1, right = -1; // neighbour ranks along the slab axis (-1 = none)
    unsigned char *win = nullptr;
    size_t winBytes = 0;
    unsigned char *peerWin[kMaxRanks] = {};
    bool ipcMapped[kMaxRanks] = {};
    size_t offChan[2] = {}, offAck[2] = {}, offU = 0, capGhost = 0, capRods = 0;
    unsigned long long seqGhost = 0, seqAck = 0, seqVec = 0, seqHalo = 0, seqMail = 0, seqCnt = 0; // lockstep counters
    DevBuf<int> sendIdx[2];    // user index of my rods mirrored on the left / right neighbour
    DevBuf<int> sendSorted[2]; // their sorted index (source rows of the U halo)
    DevBuf<int> mirror[2];     // per sorted rod: its row on the left / right neighbour, -1 if not mirrored there
    bool fused = false;        // every rank has its own device: kernels may wait on peers (see solver.cu)
    int devOfRank[kMaxRanks] = {};
    int nSend[2] = {0, 0}, nRecv[2] = {0, 0};
    int pushBase[2] = {0, 0}; // CommHeader::ghostBase as read back after the ack exchange
};

struct Context {
    int device = 0, rank = 0, nranks = 1;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    std::string err;

    Box box{};
    bool haveBox = false;
    double dRatio = 1.0, lRatio = 1.0, colBuf = 0.0;
    double viscosity = 0.0;
    bool haveMob = false;

    // ---- slab decomposition (nranks > 1) ----
    Comm comm;
    int slabAxis = 0;
    double slabLo = 0, slabHi = 0; // my slab along slabAxis
    double skin = 0;               // how far a rod may stray outside its owner's slab
    double ghostWidth = 0;         // cutoff + skin
    double maxRadiusGlobal = 0;    // max over ALL ranks of lengthCollision/2 + radiusCollision (sets the cell size)
    double meanRLocal = 0;         // mean of the same quantity (polydisperse rods: see shortR)
    double gridMaxR = 0;           // the largest bounding radius the pair search has to serve (all ranks)
    double shortR = 0;             // bounding radius the cell grid is sized for; rods above it take the long-rod pass
    double optLongRods = 2.0;      // shortR = this factor x mean bounding radius when the longest rod exceeds it (0: off)
    long long nLongRods = 0, nLongRows = 0; // statistics of the last collect
    double maxRLocal = 0;          // max over this context's rods of lengthCollision/2 + radiusCollision ...
    double maxRLRatio = -1, maxRDRatio = -1; // ... for these collision ratios (recomputed on the device when they change)
    int strays = 0;
    int globalBase = 0;            // global index of my first rod (updateSylinderMap, SylinderSystem.cpp:868-880)

    // ---- rods, user order (what the host uploaded; ghosts appended behind the nLocal owned rods) ----
    int nRods = 0;  // owned + ghost rods: what the kernels see
    int nLocal = 0; // owned rods: what the caller sees
    int nGhost = 0;
    DevBuf<signed char> uImg; // image of a ghost along the slab axis (-1, 0, +1); 0 for owned rods
    DevBuf<int> uGlobalIdx;   // global index (owner's numbering)
    DevBuf<long long> uTag; // the host's per-rod tag (e.g. Sylinder::group): travels with a migrating rod
    bool haveTags = false;
    DevBuf<int> uGid;
    DevBuf<double> uPos, uQuat, uLen, uRad; // 3n, 4n, n, n
    DevBuf<unsigned char> uImm;
    DevBuf<int> uCell;      // cell id per rod
    DevBuf<int> userToSorted;
    DevBuf<double> uVelNC;  // 6n, user order
    bool haveVelNC = false;
    cudaStream_t copyStream = nullptr;      // alens_set_velocity_noncon_async
    cudaEvent_t evVelNC = nullptr, evMain = nullptr;
    bool velNCPending = false;              // a side-stream copy into uVelNC has not been waited for yet

    // ---- cell list + rods, sorted (cell-major) order ----
    CellGrid grid{};
    DevBuf<int> cellCount, cellStart, cellFill; // ncell(+1)
    DevBuf<int> scanTmp;                        // tile sums / offsets of the tiled scan
    DevBuf<int> sUser;                          // sorted -> user index
    DevBuf<int> sGid;
    DevBuf<double> sX, sY, sZ, sDx, sDy, sDz, sLc, sRc; // collision geometry
    DevBuf<double> sLen, sRad;                           // hydrodynamic length/radius
    DevBuf<float> bUx, bUy, bUz, bH, bRho;               // broad phase: unit axis, half length, radius (fp32)
    DevBuf<unsigned char> sImm;
    DevBuf<unsigned char> sGhost; // 1 = ghost rod (owned by a neighbour rank)
    DevBuf<signed char> sImg;     // image along the slab axis
    DevBuf<double> sInvDrag; // 3 per rod: 1/para, 1/perp, 1/rot (0 if immovable)
    DevBuf<double> sMobRec;  // the same with the direction, as one 64-byte record per rod (k_force_vel_rec)
    bool sorted = false;

    // ---- constraints (solver order) ----
    long long nCon = 0, nColl = 0; // total / produced by pair collection
    DevBuf<int> cellHits, cellHitStart;
    DevBuf<int4> hitList;      // staged hits of k_pairs_find: (i, j, cell, seq<<5 | image code)
    DevBuf<int> cIdxI, cIdxJ; // sorted rod index; cIdxJ = -1 for oneSide
    DevBuf<int> cGidI, cGidJ;
    DevBuf<double> cN, cPI, cPJ;     // SoA by component: [3][cap] each (stride = conCap)
    DevBuf<double> cLabI, cLabJ;     // [3][cap]
    DevBuf<double> cDelta0, cGamma0, cInvKappa; // invKappa = 1/kappa (not yet /dt)
    DevBuf<double> cKappa;
    DevBuf<unsigned char> cBi, cOneSide;
    DevBuf<unsigned char> cOwn; // 1 = this rank counts the row in global dot products (owner of rod I)
    DevBuf<signed char> cShift; // image of J relative to I, code = (kx+1)+3(ky+1)+9(kz+1)
    DevBuf<double> cStressHost; // 9 per appended block (host supplied), indexed k - nColl
    size_t conCap = 0;          // component stride of the SoA arrays
    std::vector<alens_constraint_block> hostBlocks; // verbatim copies of appended blocks (for the refill)
    long long statCand = 0;
    DevBuf<unsigned long long> dCounters; // [0] candidates, [1] hits

    // ---- incidence (rod -> constraints), built in setup ----
    DevBuf<int> incDeg, incStart, incFill; // nRods(+1)
    DevBuf<int> incCon;                    // 4*constraint + 2*bilateral + side per slot; rod-major (force_kernel 1/2) or level-major inside a 32-rod group (0)
    DevBuf<int> incRaw;                    // rod-major slot lists before k_inc_emit
    DevBuf<double> incCol;                 // 6 per slot: D column block for that (rod, constraint)
    long long nInc = 0, incStride = 0; // slots; component stride of incCol (multiple of 4, > nInc)
    long long nOneSide = 0, nBilateral = 0; // host-side counts of appended one-sided / bilateral blocks
    int optForceKernel = 3;                 // 3 = k_force_vel_rec (64-byte slot records + slot-ordered live bitmap kept by k_bb_tail), 1 = k_force_vel_act (rod-major slots, zero multipliers skipped), 0 = k_force_vel_lm (level-major, dense)
    int optTailPush = 0;  // fused multi-GPU, measured alternative (slower): the tail kernel instead of the force kernel copies the mirrored rows of U to the neighbours (contiguous staging rows)
    int optHaloDebug = 0; // multi-GPU timing experiments only (results are wrong): 1 = no remote U stores, 2 = no fence
    int optStamps = 0, stampCap = 0, stampIters = 0; // per-iteration nanosecond stamps of the BBPGD kernels (instrumentation)
    DevBuf<unsigned long long> dStamps;
    unsigned long long *stampNow = nullptr;
    bool recDirty = false;                  // rec_mode 2: alens_calc_mobility ran after the setup, the records must be rebuilt
    int optRecMode = 2, recMode = 2;        // force_kernel 3: 0 = k_bb_tail copies {x, g} into the slot records, 1 = the records hold the row id and k_force_vel_rec gathers {x, g} itself, 2 = as 1 with M * column in the record (no mobility read)
    DevBuf<double> incRec;                  // force_kernel 3: 8 doubles per slot {x, g, D column block[6]}, 64-byte aligned records
    DevBuf<int2> cSlot;                     // ... per constraint: slot of its I side / J side (-1: none, ghost or one-sided)
    DevBuf<int> cIdxIU, cIdxJU; // fused multi-GPU: the rod rows the tail kernel gathers U from (ghost rods: the staging rows behind nRods)
    DevBuf<int2> rodHead; // force_kernel 3: per rod {first slot, live bits of its first 32 slots}
    DevBuf<unsigned> slotBi;                // ... bit per slot: the slot's constraint is bilateral (constant during a solve)
    int optFindSplitMinB = 8;               // ... resident CTAs per SM of its stage-1/2 kernel (8: 64 registers)
    int optFindSplit = 1;                   // pair search: stages 1-2 -> candidates, dense narrow phase + ordered emission (0: one kernel)
    int candWords = 16;                     // bitmap words per cell (32 candidates each); follows the fullest cell of the last step
    long long lastCand = 0;                 // staged candidates of the last step (sizes the dense grids and the list)
    DevBuf<unsigned> candBits;              // [ncell][candWords] contact flags of the staged candidates
    DevBuf<int> candPrefix, cellCand;       // exclusive popcounts of the bitmap words; candidates per cell
    int optFindMinB = 4;                    // k_pairs_find: resident CTAs per SM asked of the compiler (4: 128 registers)
    int optForceSplit = 0;                  // force_kernel = 2: k_slot_x + k_rod_sum instead of k_force_vel_act
    DevBuf<double> slotX;                   // multiplier per incidence slot (k_slot_x -> k_rod_sum)
    DevBuf<unsigned> slotLive;              // bit per incidence slot: multiplier non-zero
    int optForceMinB = 4;                   // k_force_vel_act: resident CTAs per SM asked of the compiler (4: 128 registers, no spills: fastest; 5: 96 + spills; 3: 168)
    int optTailRing = 0;                    // k_bb_tail_ring (TMA bulk copies into a shared-memory ring) instead of k_bb_tail
    int optPdl = 1;                         // BBPGD kernels launched with programmatic stream serialization (single rank)
    int optPoll = 1;                        // BBPGD host loop throttled by a progress word in pinned memory instead of stream syncs
    int optLookahead = 3;                   // iterations queued behind the running one
    bool pdlNow = false, profMute = false;  // state of the current solve
    int *hProg = nullptr, *hProgDev = nullptr; // {completed applies, done} in mapped pinned host memory
    int optLateHalo = 1;                    // multi-GPU: k_bb_tail walks the rows that need no halo first and waits for the halo behind them
    DevBuf<int> tailFlag, tailOrder;        // multi-GPU: tiles of 256 constraint rows; order = tiles without a ghost-reading row first, [nTiles] = their number
    DevBuf<int> rodFlag, rodOrder;          // ... tiles of 128 rods; order = tiles with a rod mirrored on a neighbour first, [nTiles] = their number
    int optKeepXG = 1;                      // {x, g} pairs stored / gathered with the L2 evict_last policy
    int optForceMask = 1;                   // k_force_vel_act consults the tail kernel's "may be non-zero" bit mask before gathering {x, g}
    int optForceWaves = 1;                  // k_force_vel_act: grid = resident CTAs x this (1 = persistent)
    int incLayout = 1;                      // layout built by the last setup (= optForceKernel at that time)
    int optForceChunk = 2;                  // k_force_vel_lm: incidence levels per software-pipeline stage (2 or 4)
    int optForceBlock = 64;                 // k_force_vel_lm: threads per CTA (small CTAs: groups differ in depth, a CTA lives as long as its deepest group)
    int optTailCtasPerSM = 2;               // persistent grid of k_bb_tail
    int optBatch = 0;                       // BBPGD iterations enqueued per host check (0 = automatic)
    bool haveSetup = false;
    double dt = 0.0;

    // ---- solver vectors ----
    DevBuf<double> vX0, vX1, vG0, vG1, vB, vLbFlag; // x0 / unpacked iterates, APGD work, q, bilateral flag as double
    DevBuf<unsigned> vMask2;                        // same for a plain vector handed to the operator (k_mask_from_x)
    DevBuf<unsigned> vMask;                         // 1 bit per constraint: 0 = the next BBPGD iterate is certainly 0 there
    DevBuf<double2> vXG0, vXG1;                     // BBPGD iterates as interleaved {x, g} pairs (ping-pong)
    DevBuf<double> vTmp0, vTmp1, vTmp2, vTmp3, vTmp4, vTmp5; // APGD work vectors
    DevBuf<double> rU, rF;                         // 6 per rod: vel, force of the last apply
    DevBuf<double> rUb, rFb;                       // bilateral part
    DevBuf<double> outFU, outVU, outFB, outVB;     // user order results
    DevBuf<double> redPartial;                     // per-block partial reductions
    DevBuf<SolverScalars> dScal;
    DevBuf<double> dHist;                          // 6 per row
    int histCap = 0;
    SolverScalars *hScal = nullptr;                // pinned mirror
    std::vector<double> hist;                      // host copy of the last history
    const double2 *lastXG = nullptr;               // {x, g} pairs the last BBPGD force kernel read
    double *xLastApplied = nullptr;                // device ptr of the vector the operator last saw
    double *xSolution = nullptr;                   // device ptr of the returned iterate
    bool haveSolution = false;
    alens_solve_report lastReport{};

    // ---- instrumentation ----
    alens_timers timers{};
    cudaEvent_t ev[8] = {};
    long long launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> profEv; // pool for per-kernel timing
    std::vector<int> profKind;       // kernel kind per event pair
    int profUsed = 0;

    PinnedBuf pin0, pin1;
    PinnedBuf pinBounce[2]; // 8 MB each: transfers from / to host memory that is not page-locked (capi.cu: downloadAny)
    PinnedBuf pinAos; // alens_set_rods_aos: the gathered hot fields of the caller's Sylinder records

    // ---- multi-GPU ----
    void *nccl = nullptr; // ncclComm_t
};

// kernels' host entry points (collide.cu / solver.cu)
void ctxInit(Context &c);
void ctxFree(Context &c);
void rodsUploaded(Context &c, bool wrap);           // rod_pack + cell list + sorted SoA
void collectPairs(Context &c);                      // broad + narrow phase
void appendBlocks(Context &c, const alens_constraint_block *b, long long n, const int *userIdx = nullptr);
void downloadBlocks(Context &c, alens_constraint_block *out, long long cap, bool withStress, bool writeBack);
void sumConstraintStress(Context &c, bool withOneSide, double uni[9], double bi[9]);
void calcMobility(Context &c, double mu);
void mobilityApply(Context &c, const double *x, double *y);
void setupConstraints(Context &c, const double *velNC, double dt);
void operatorApply(Context &c, const double *x, double *y, double *force, double *vel);
void solveConstraints(Context &c, double res, int maxIte, int choice);
void solveCore(Context &c, double tol, int maxIte, int choice);
void stepEuler(Context &c, double dt);
void profFlush(Context &c, int maxEvents = 1 << 30);
double timeKernel(Context &c, int which, int reps);
// make the main stream wait for a pending alens_set_velocity_noncon_async copy (before anything reads uVelNC)
inline void waitVelNC(Context &c) {
    if (c.velNCPending) {
        cudaStreamWaitEvent(c.stream, c.evVelNC, 0);
        c.velNCPending = false;
    }
}
void calcVelocityBrown(Context &c, double kBT, double dt, const double *normals12, unsigned long long seed, unsigned long long step, double *out);
void calcVelocityNonCon(Context &c, const double *force, const double *velNB, const double *velB, int monolayer, double *velNonBOut);
long long collectBoundary(Context &c, const alens_boundary *bnd, int nb);
long long collectProteins(Context &c, const alens_protein_bind *proteins, long long n, double tubuleDiameter);
long long collectLinks(Context &c, const int *prevGid, const int *nextGid, long long nLinks, double linkKappa, double linkGap);
void dcpBatch(Context &c, long long n, const double *P0, const double *P1, const double *Q0, const double *Q1, double *dist,
              double *Ploc, double *Qloc);
void pairFunctorBatch(Context &c, long long n, const double *geomI, const double *geomJ, int withStress, unsigned char *hit,
                      alens_constraint_block *blocks);
// bcqp.cu: BCQPSolver for any caller (CSR matrix or the constraint operator, caller-set bounds)
struct Bcqp;
Bcqp *bcqpCreate(Context &c, int n, const long long *rowPtr, const int *col, const double *val, const double *b);
void bcqpSetBounds(Bcqp &q, const double *lb, const double *ub, int which);
void bcqpGetBounds(Bcqp &q, double *lb, double *ub);
void bcqpSolve(Bcqp &q, double *x, double tol, int iteMax, int choice, alens_solve_report *rep);
int bcqpHistory(Bcqp &q, double *rows6, int cap);
int bcqpSize(Bcqp &q);
Context *bcqpContext(Bcqp &q);
void bcqpDestroy(Bcqp *q);
void liveStats(Context &c, long long *slots, long long *rods);
void constraintDigest(Context &c, unsigned long long u64[3], double f64[3]);
long long mixPairSearch(Context &c, long long nTrg, const double *trgPos, const double *trgRs, const double *srcRs,
                        long long *rowPtr, int *srcIdx, long long cap);
void preloadCollideKernels();
void preloadSolverKernels();
void preloadBlockKernels();
void preloadCommKernels();
// comm.cu
void commAllocWindow(Context &c, long long maxLocalRods);
void commExport(Context &c, void *blob);
void commImport(Context &c, const void *blobs);
void commConnectLocal(Context **ctxs, int n);
void commFree(Context &c);
void commExchangeGhosts(Context &c);
void commExchangeGhostIndices(Context &c);
void commHaloVelNC(Context &c);
void commPushU(Context &c, unsigned long long seq);
void commSignalHalo(Context &c, unsigned long long seq);
void checkCommError(Context &c); // throws ALENS_ERR_COMM if a device-side wait of this rank has timed out
void reserveConstraints(Context &c, size_t n, bool keep);

inline int gridFor(long long n, int block) { return (int)((n + block - 1) / block); }

// shared small kernels (collide.cu)
void launchScanInt(Context &c, const int *in, int *out, int n);
void wrapRodPositions(Context &c);
void commMigrate(Context &c, long long *nSent, long long *nReceived);
double hostMaxRadius(int n, const double *len, const double *rad, double lRatio, double dRatio, double *meanOut = nullptr);

} // namespace alens
