// comm.cu -- multi-GPU plumbing of the slab decomposition: one rank per GPU, rods partitioned
// into slabs along one box axis, ghost rods within (cutoff + skin) of the slab faces mirrored on the neighbour.
//
// Replaces, for the constraint path, the FDPS ghost machinery (exchangeLocalEssentialTree inside
// TreeSylinderNear::calcForceAll, FDPS/tree_for_force.hpp:759-842), Tpetra's Import of ghost columns
// (ConstraintOperator.cpp:44-57) and the Teuchos/MPI allreduces of BCQPSolver.cpp:183-233.
//
// Transport = peer memory over NVLink: every rank owns one "window" (device allocation) that its peers map
// (same process: direct pointers; other processes: cudaIpcOpenMemHandle).  Payloads are written by kernels
// with plain remote stores, followed by a system-scope release of a 64-bit sequence number in the receiver's
// window; receivers poll their OWN memory with acquire loads.  No kernel of the BCQP loop waits on a peer:
// waiting is done by one-thread kernels between them (k_wait_seq / k_bb_reduce), so two ranks can also share
// one GPU (tests) without starving each other.
#include "context.hpp"
#include "comm_dev.cuh"

#include <algorithm>
#include <cstring>

namespace alens {

// ------------------------------------------------------------------------------------------------
__global__ void k_wait_seq(const unsigned long long *f0, const unsigned long long *f1, unsigned long long want,
                           int *err, const SolverScalars *scal) {
    if (scal && scal->done) return;
    if (f0) waitSeq(f0, want, err);
    if (f1) waitSeq(f1, want, err);
}

// publish `seq` (and optionally a count) in up to two peers' windows, after everything this stream wrote before
__global__ void k_signal(unsigned long long *f0, long long *c0, long long n0, unsigned long long *f1, long long *c1,
                         long long n1, unsigned long long seq, const SolverScalars *scal) {
    if (scal && scal->done) return;
    __threadfence_system();
    if (f0) {
        if (c0) *c0 = n0;
        __threadfence_system();
        stReleaseSys(f0, seq);
    }
    if (f1) {
        if (c1) *c1 = n1;
        __threadfence_system();
        stReleaseSys(f1, seq);
    }
}

// ------------------------------------------------------------------------------------------------
// ghost selection: local rods whose apparent coordinate along the slab axis lies within `gw` of a slab face
__global__ void k_ghost_flags(int n, const double *__restrict__ pos, const signed char *__restrict__ img, double boxLen,
                              int axis, double lo, double hi, double gw, int haveLeft, int haveRight,
                              int *__restrict__ fL, int *__restrict__ fR) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[3 * (size_t)i + axis] + img[i] * boxLen;
    fL[i] = (haveLeft && x < lo + gw) ? 1 : 0;
    fR[i] = (haveRight && x >= hi - gw) ? 1 : 0;
}
__global__ void k_ghost_list(int n, const int *__restrict__ flag, const int *__restrict__ scan, int *__restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[scan[i]] = i;
}

// one ghost record = 18 doubles: pos3, quat4, length, radius, velNonCon6, (gid, globalIndex), (immovable, image), host tag
__global__ void k_ghost_pack(int n, const int *__restrict__ list, GhostSrc s, int image, double *__restrict__ dst) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int i = list[e];
    double *d = dst + (size_t)e * kGhostRec;
    d[0] = s.pos[3 * (size_t)i]; d[1] = s.pos[3 * (size_t)i + 1]; d[2] = s.pos[3 * (size_t)i + 2];
    for (int k = 0; k < 4; k++) d[3 + k] = s.quat[4 * (size_t)i + k];
    d[7] = s.len[i];
    d[8] = s.rad[i];
    for (int k = 0; k < 6; k++) d[9 + k] = s.velNC ? s.velNC[6 * (size_t)i + k] : 0.0;
    d[15] = __hiloint2double(s.gid[i], s.globalBase + i);
    d[16] = __hiloint2double(s.imm ? (int)s.imm[i] : 0, image + s.img[i]); // own image + the channel's
    d[17] = __longlong_as_double(s.tag ? s.tag[i] : 0LL);
}
__global__ void k_ghost_unpack(int n, const double *__restrict__ src, GhostDst o, int base) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double *d = src + (size_t)e * kGhostRec;
    const size_t i = (size_t)base + e;
    o.pos[3 * i] = d[0]; o.pos[3 * i + 1] = d[1]; o.pos[3 * i + 2] = d[2];
    for (int k = 0; k < 4; k++) o.quat[4 * i + k] = d[3 + k];
    o.len[i] = d[7];
    o.rad[i] = d[8];
    if (o.velNC)
        for (int k = 0; k < 6; k++) o.velNC[6 * i + k] = d[9 + k];
    o.gid[i] = __double2hiint(d[15]);
    o.globalIdx[i] = __double2loint(d[15]);
    o.imm[i] = (unsigned char)__double2hiint(d[16]);
    o.img[i] = (signed char)__double2loint(d[16]);
    if (o.tag) o.tag[i] = __double_as_longlong(d[17]);
}

// after the cell sort: tell the sender where each of its ghosts sits in my sorted arrays
__global__ void k_ack_indices(int n, const int *__restrict__ userToSorted, int base, int *__restrict__ dstRemote,
                              int *__restrict__ remoteGhostBase, int ghostBase) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dstRemote[e] = userToSorted[base + e];
    if (e == 0) *remoteGhostBase = ghostBase; // where this neighbour's velocities are expected in my U (staging rows)
}
__global__ void k_map_indices(int n, const int *__restrict__ list, const int *__restrict__ userToSorted,
                              int *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = userToSorted[list[e]];
}

__global__ void k_fill_int(int n, int v, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
__global__ void k_build_mirror(int n, const int *__restrict__ sorted, const int *__restrict__ remote,
                               int *__restrict__ mirror) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) mirror[sorted[e]] = remote[e];
}

// 6-vector halo: dst[6*dstIdx[e] ..] = src[6*srcIdx[e] ..] for the rods mirrored on a neighbour (remote stores)
__global__ void k_halo_push6(int n0, const int *__restrict__ src0, const int *__restrict__ dst0, double *__restrict__ out0,
                             int n1, const int *__restrict__ src1, const int *__restrict__ dst1, double *__restrict__ out1,
                             const double *__restrict__ vec, const SolverScalars *scal) {
    if (scal && scal->done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = t / 3, part = t - 3 * e;
    if (e < n0) {
        const double2 v = reinterpret_cast<const double2 *>(vec + 6 * (size_t)src0[e])[part];
        const size_t d = dst0 ? (size_t)dst0[e] : (size_t)e;
        reinterpret_cast<double2 *>(out0 + 6 * d)[part] = v;
    } else if (e - n0 < n1) {
        const int f = e - n0;
        const double2 v = reinterpret_cast<const double2 *>(vec + 6 * (size_t)src1[f])[part];
        const size_t d = dst1 ? (size_t)dst1[f] : (size_t)f;
        reinterpret_cast<double2 *>(out1 + 6 * d)[part] = v;
    }
}
__global__ void k_copy6_rows(int n, const double *__restrict__ src, double *__restrict__ dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 6 * n) dst[t] = src[t];
}

// ------------------------------------------------------------------------------------------------
static CommHeader *hdrOf(unsigned char *win) { return reinterpret_cast<CommHeader *>(win); }

void commAllocWindow(Context &c, long long maxLocalRods) {
    Comm &m = c.comm;
    if (m.win) throw ArgError{ALENS_ERR_STATE, "comm: window already allocated"};
    if (c.nranks > kMaxRanks) throw ArgError{ALENS_ERR_UNSUPPORTED, "comm: more than 16 ranks"};
    m.capGhost = (size_t)std::max<long long>(maxLocalRods / 3, 4096);
    m.capRods = (size_t)maxLocalRods + 4 * m.capGhost + 64; // owned + ghost rods, + the ghosts' staging rows of U
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = al(sizeof(CommHeader));
    for (int d = 0; d < 2; d++) { m.offChan[d] = off; off = al(off + m.capGhost * kGhostRec * sizeof(double)); }
    for (int d = 0; d < 2; d++) { m.offAck[d] = off; off = al(off + m.capGhost * sizeof(int)); }
    m.offU = off;
    off = al(off + m.capRods * 6 * sizeof(double));
    m.winBytes = off;
    ALENS_CUDA(cudaMalloc((void **)&m.win, m.winBytes));
    ALENS_CUDA(cudaMemset(m.win, 0, m.winBytes));
    ALENS_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < kMaxRanks; r++) m.peerWin[r] = nullptr;
    m.peerWin[c.rank] = m.win;
}

void commFinishConnect(Context &c) {
    Comm &m = c.comm;
    const int R = c.nranks, r = c.rank;
    const bool per = c.box.pbc[c.slabAxis] != 0;
    m.left = r > 0 ? r - 1 : (per && R > 1 ? R - 1 : -1);
    m.right = r < R - 1 ? r + 1 : (per && R > 1 ? 0 : -1);
    for (int q = 0; q < R; q++)
        if (!m.peerWin[q]) throw ArgError{ALENS_ERR_COMM, "comm: a peer window is missing"};
    // the rod velocity vector of the operator lives in the window so that neighbours can write ghost rows
    c.rU.release();
    c.rU.p = reinterpret_cast<double *>(m.win + m.offU);
    c.rU.cap = m.capRods * 6;
    c.rU.external = true;
    m.active = R > 1;
    m.fused = true; // waiting inside compute kernels is only safe when no two ranks share a device
    for (int a = 0; a < R; a++)
        for (int b = a + 1; b < R; b++)
            if (m.devOfRank[a] == m.devOfRank[b]) m.fused = false;
    preloadCollideKernels();
    preloadSolverKernels();
    preloadBlockKernels();
    preloadCommKernels();
    // the runtime's own memset / copy kernels as well
    ALENS_CUDA(cudaMemsetAsync(m.win + m.offU, 0, 256, c.stream));
    ALENS_CUDA(cudaMemcpyAsync(m.win + m.offU + 256, m.win + m.offU, 128, cudaMemcpyDeviceToDevice, c.stream));
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
}

void commExport(Context &c, void *blob) {
    CommBlob b{};
    ALENS_CUDA(cudaIpcGetMemHandle(&b.handle, c.comm.win));
    b.bytes = (unsigned long long)c.comm.winBytes;
    b.device = c.device;
    b.rank = c.rank;
    memcpy(blob, &b, sizeof(b));
}

void commImport(Context &c, const void *blobs) {
    Comm &m = c.comm;
    for (int q = 0; q < c.nranks; q++) {
        if (q == c.rank) continue;
        CommBlob b;
        memcpy(&b, (const char *)blobs + (size_t)q * sizeof(CommBlob), sizeof(b));
        if (b.rank != q || b.bytes != m.winBytes)
            throw ArgError{ALENS_ERR_COMM, "comm: peer blob mismatch (rank order / window size must agree on all ranks)"};
        void *p = nullptr;
        ALENS_CUDA(cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess));
        m.peerWin[q] = (unsigned char *)p;
        m.ipcMapped[q] = true;
        m.devOfRank[q] = b.device;
    }
    m.devOfRank[c.rank] = c.device;
    commFinishConnect(c);
}

void commConnectLocal(Context **ctxs, int n) {
    for (int a = 0; a < n; a++) {
        Context &c = *ctxs[a];
        if (c.nranks != n || c.rank != a) throw ArgError{ALENS_ERR_ARG, "comm: contexts must be given in rank order"};
        ALENS_CUDA(cudaSetDevice(c.device));
        for (int b = 0; b < n; b++) {
            Context &p = *ctxs[b];
            if (p.comm.winBytes != c.comm.winBytes) throw ArgError{ALENS_ERR_ARG, "comm: window sizes differ"};
            if (p.device != c.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(p.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ALENS_CUDA(e);
                cudaGetLastError();
            }
            c.comm.peerWin[b] = p.comm.win;
            c.comm.devOfRank[b] = p.device;
        }
    }
    for (int a = 0; a < n; a++) {
        ALENS_CUDA(cudaSetDevice(ctxs[a]->device));
        commFinishConnect(*ctxs[a]);
    }
}

void commFree(Context &c) {
    Comm &m = c.comm;
    for (int q = 0; q < kMaxRanks; q++)
        if (m.ipcMapped[q] && m.peerWin[q]) cudaIpcCloseMemHandle(m.peerWin[q]);
    if (m.win) {
        c.rU.p = nullptr;
        c.rU.cap = 0;
        cudaFree(m.win);
    }
    m.win = nullptr;
    m.active = false;
}

void checkCommError(Context &c) {
    int err = 0;
    ALENS_CUDA(cudaMemcpyAsync(&err, &hdrOf(c.comm.win)->error, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    ALENS_CUDA(cudaStreamSynchronize(c.stream));
    if (err) throw ArgError{ALENS_ERR_COMM, "comm: timed out waiting for a neighbour rank"};
}

// neighbour `dir` (0 = left, 1 = right): its window, and the channel index my data arrives on over there
static unsigned char *nbWin(Context &c, int dir) {
    const int q = dir == 0 ? c.comm.left : c.comm.right;
    return q < 0 ? nullptr : c.comm.peerWin[q];
}

// ------------------------------------------------------------------------------------------------
// Ghost exchange, first half (before the cell sort): select, send, receive, append after the local rods.
void commExchangeGhosts(Context &c) {
    Comm &m = c.comm;
    cudaStream_t st = c.stream;
    waitVelNC(c); // ghost records carry the owner's velNonCon rows
    const int n = c.nLocal, ax = c.slabAxis;
    const double gw = c.ghostWidth;
    CommHeader *me = hdrOf(m.win);
    DevBuf<int> &fL = c.incDeg, &fR = c.incFill; // scratch (rebuilt by the setup)
    fL.reserve(n + 8); fR.reserve(n + 8);
    c.incStart.reserve(n + 8);
    DevBuf<int> scanL, scanR;
    scanL.reserve(n + 8); scanR.reserve(n + 8);
    if (n > 0)
        k_ghost_flags<<<gridFor(n, 256), 256, 0, st>>>(n, c.uPos.p, c.uImg.p, c.box.len[ax], ax, c.slabLo, c.slabHi, gw,
                                                       m.left >= 0, m.right >= 0, fL.p, fR.p);
    launchScanInt(c, fL.p, scanL.p, n);
    launchScanInt(c, fR.p, scanR.p, n);
    int cnt[2] = {0, 0};
    ALENS_CUDA(cudaMemcpyAsync(&cnt[0], scanL.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(&cnt[1], scanR.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    for (int d = 0; d < 2; d++) {
        if ((size_t)cnt[d] > m.capGhost) throw ArgError{ALENS_ERR_COMM, "comm: ghost channel capacity exceeded (alens_comm_create maxLocalRods too small)"};
        m.nSend[d] = cnt[d];
        m.sendIdx[d].reserve((size_t)cnt[d] + 1);
        m.sendSorted[d].reserve((size_t)cnt[d] + 1);
    }
    if (n > 0) {
        k_ghost_list<<<gridFor(n, 256), 256, 0, st>>>(n, fL.p, scanL.p, m.sendIdx[0].p);
        k_ghost_list<<<gridFor(n, 256), 256, 0, st>>>(n, fR.p, scanR.p, m.sendIdx[1].p);
    }
    // pack straight into the neighbours' windows; image = how the rod appears in the receiver's frame
    const unsigned long long seq = ++m.seqGhost;
    GhostSrc src{c.uGid.p, c.uPos.p, c.uQuat.p, c.uLen.p, c.uRad.p, c.uImm.p, c.uImg.p,
                 c.haveVelNC ? c.uVelNC.p : nullptr, c.globalBase, nullptr};
    unsigned long long *sf[2] = {nullptr, nullptr};
    long long *sc[2] = {nullptr, nullptr};
    for (int d = 0; d < 2; d++) {
        unsigned char *w = nbWin(c, d);
        if (!w) continue;
        const int ch = 1 - d; // I am the right neighbour of my left neighbour
        int image = 0;
        if (d == 0 && c.rank == 0) image = +1;               // wraps around the low face: appears at x + L
        if (d == 1 && c.rank == c.nranks - 1) image = -1;    // wraps around the high face
        if (cnt[d] > 0)
            k_ghost_pack<<<gridFor(cnt[d], 128), 128, 0, st>>>(cnt[d], m.sendIdx[d].p, src, image,
                                                               reinterpret_cast<double *>(w + m.offChan[ch]));
        sf[d] = &hdrOf(w)->chanSeq[ch];
        sc[d] = &hdrOf(w)->chanCount[ch];
    }
    k_signal<<<1, 1, 0, st>>>(sf[0], sc[0], cnt[0], sf[1], sc[1], cnt[1], seq, nullptr);
    // receive
    k_wait_seq<<<1, 1, 0, st>>>(m.left >= 0 ? &me->chanSeq[0] : nullptr, m.right >= 0 ? &me->chanSeq[1] : nullptr, seq,
                                &me->error, nullptr);
    long long rc[2] = {0, 0};
    ALENS_CUDA(cudaMemcpyAsync(rc, me->chanCount, sizeof(rc), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    checkCommError(c);
    m.nRecv[0] = m.left >= 0 ? (int)rc[0] : 0;
    m.nRecv[1] = m.right >= 0 ? (int)rc[1] : 0;
    const int nAll = n + m.nRecv[0] + m.nRecv[1];
    if ((size_t)nAll > m.capRods) throw ArgError{ALENS_ERR_COMM, "comm: rod capacity of the window exceeded"};
    const size_t N = (size_t)nAll;
    c.uGid.reserve(N + 1, st, true, n); c.uPos.reserve(3 * N + 3, st, true, 3 * (size_t)n);
    c.uQuat.reserve(4 * N + 4, st, true, 4 * (size_t)n); c.uLen.reserve(N + 1, st, true, n);
    c.uRad.reserve(N + 1, st, true, n); c.uImm.reserve(N + 1, st, true, n);
    c.uImg.reserve(N + 1, st, true, n); c.uGlobalIdx.reserve(N + 1, st, true, n);
    if (c.haveVelNC) c.uVelNC.reserve(6 * N + 6, st, true, 6 * (size_t)n);
    GhostDst dst{c.uGid.p, c.uPos.p, c.uQuat.p, c.uLen.p, c.uRad.p, c.uImm.p, c.uImg.p, c.uGlobalIdx.p,
                 c.haveVelNC ? c.uVelNC.p : nullptr, nullptr};
    int base = n;
    for (int ch = 0; ch < 2; ch++) {
        if (m.nRecv[ch] > 0)
            k_ghost_unpack<<<gridFor(m.nRecv[ch], 128), 128, 0, st>>>(
                m.nRecv[ch], reinterpret_cast<const double *>(m.win + m.offChan[ch]), dst, base);
        base += m.nRecv[ch];
    }
    c.nGhost = nAll - n;
    c.nRods = nAll;
    c.launches += 8;
    ALENS_CUDA(cudaGetLastError());
}

// second half (after the cell sort): return the sorted index of every received ghost to its owner and learn
// where my own mirrored rods sit on the neighbours.
void commExchangeGhostIndices(Context &c) {
    Comm &m = c.comm;
    cudaStream_t st = c.stream;
    CommHeader *me = hdrOf(m.win);
    const unsigned long long seq = ++m.seqAck;
    unsigned long long *sf[2] = {nullptr, nullptr};
    int base = c.nLocal;
    for (int ch = 0; ch < 2; ch++) { // channel ch came from my left (0) / right (1) neighbour
        unsigned char *w = nbWin(c, ch);
        if (!w) continue;
        const int dirThere = 1 - ch; // over there I am its right / left neighbour
        // (always launched: the staging base goes over with the acks.  Ghost g of this rank, user index nLocal + g, has its
        // velocity expected in row nRods + g of U: contiguous per neighbour, in the neighbour's send order)
        k_ack_indices<<<std::max(1, gridFor(m.nRecv[ch], 256)), 256, 0, st>>>(
            m.nRecv[ch], c.userToSorted.p, base, reinterpret_cast<int *>(w + m.offAck[dirThere]),
            &hdrOf(w)->ghostBase[dirThere], c.nRods + (base - c.nLocal));
        sf[ch] = &hdrOf(w)->ackSeq[dirThere];
        base += m.nRecv[ch];
    }
    k_signal<<<1, 1, 0, st>>>(sf[0], nullptr, 0, sf[1], nullptr, 0, seq, nullptr);
    k_wait_seq<<<1, 1, 0, st>>>(m.left >= 0 ? &me->ackSeq[0] : nullptr, m.right >= 0 ? &me->ackSeq[1] : nullptr, seq,
                                &me->error, nullptr);
    for (int d = 0; d < 2; d++)
        if (m.nSend[d] > 0)
            k_map_indices<<<gridFor(m.nSend[d], 256), 256, 0, st>>>(m.nSend[d], m.sendIdx[d].p, c.userToSorted.p,
                                                                   m.sendSorted[d].p);
    // per-rod mirror rows for the fused halo push of the force kernel
    for (int d = 0; d < 2; d++) {
        m.mirror[d].reserve((size_t)c.nRods + 1);
        if (c.nRods > 0) k_fill_int<<<gridFor(c.nRods, 256), 256, 0, st>>>(c.nRods, -1, m.mirror[d].p);
        if (m.nSend[d] > 0 && nbWin(c, d))
            k_build_mirror<<<gridFor(m.nSend[d], 256), 256, 0, st>>>(
                m.nSend[d], m.sendSorted[d].p, reinterpret_cast<const int *>(m.win + m.offAck[d]), m.mirror[d].p);
    }
    c.launches += 8;
    ALENS_CUDA(cudaGetLastError());
    ALENS_CUDA(cudaMemcpyAsync(m.pushBase, me->ghostBase, sizeof(m.pushBase), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    checkCommError(c);
}

// ------------------------------------------------------------------------------------------------
// Rod migration (the device-side counterpart of decomposeDomain + exchangeSylinder, SylinderSystem.cpp:617-620, for slabs):
// owned rods whose (wrapped) centre left the slab move to the neighbour slab through the ghost channels, the remaining
// rods keep their order, arrivals are appended (left neighbour's first), and every rank learns every rank's new count
// (-> globalIndex base, updateSylinderMap :868-880).  Collective; call between alens_step_euler and alens_prepare_step.
__global__ void k_migrate_flags(int n, const double *__restrict__ pos, int axis, double lo, double hi, double boxLen,
                                int periodic, int haveLeft, int haveRight, int *__restrict__ fL, int *__restrict__ fR,
                                int *__restrict__ keep, int *__restrict__ lost) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pos[3 * (size_t)i + axis];
    double xa = x; // the periodic image of the centre that is nearest to my slab
    if (periodic) {
        const double mid = 0.5 * (lo + hi);
        if (fabs(x - boxLen - mid) < fabs(xa - mid)) xa = x - boxLen;
        if (fabs(x + boxLen - mid) < fabs(xa - mid)) xa = x + boxLen;
    }
    const int L = (haveLeft && xa < lo) ? 1 : 0, R = (haveRight && xa >= hi) ? 1 : 0;
    const double w = hi - lo;
    if ((L && xa < lo - w) || (R && xa >= hi + w)) atomicAdd(lost, 1); // further than the next slab: not a neighbour's rod
    fL[i] = L;
    fR[i] = R;
    keep[i] = (L || R) ? 0 : 1;
}
template <typename T, int W>
__global__ void k_compact(int n, const int *__restrict__ keep, const int *__restrict__ scan, const T *__restrict__ src,
                          T *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const size_t o = (size_t)scan[i];
    for (int w = 0; w < W; w++) dst[W * o + w] = src[W * (size_t)i + w];
}
struct PeerHeaders {
    CommHeader *h[kMaxRanks];
};
__global__ void k_publish_count(PeerHeaders peers, int R, int myRank, long long count, unsigned long long seq) {
    const int q = threadIdx.x;
    if (q >= R) return;
    peers.h[q]->cnt[myRank] = count;
    __threadfence_system();
    stReleaseSys(&peers.h[q]->cntSeq[myRank], seq);
}
__global__ void k_wait_counts(CommHeader *me, int R, unsigned long long seq) {
    const int q = threadIdx.x;
    if (q < R) waitSeq(&me->cntSeq[q], seq, &me->error);
}

template <typename T, int W>
static void compactInto(Context &c, int n, int nKeep, int nNew, const int *keep, const int *scan, DevBuf<T> &buf) {
    DevBuf<T> out;
    out.reserve((size_t)W * ((size_t)std::max(nNew, n) + 1) + 8);
    if (n > 0) k_compact<T, W><<<gridFor(n, 256), 256, 0, c.stream>>>(n, keep, scan, buf.p, out.p);
    std::swap(buf.p, out.p);
    std::swap(buf.cap, out.cap);
    (void)nKeep;
}

void commMigrate(Context &c, long long *nSent, long long *nReceived) {
    Comm &m = c.comm;
    if (!m.active) throw ArgError{ALENS_ERR_STATE, "alens_migrate_rods: no communicator (alens_comm_connect first)"};
    if (!c.uPos.p && c.nLocal > 0) throw ArgError{ALENS_ERR_STATE, "alens_migrate_rods: no resident rods"};
    cudaStream_t st = c.stream;
    waitVelNC(c);
    const int n = c.nLocal, ax = c.slabAxis;
    CommHeader *me = hdrOf(m.win);
    wrapRodPositions(c); // positions into the box first (applyBoxBC), as prepareStep does
    DevBuf<int> fL, fR, keep, scanL, scanR, scanK;
    fL.reserve(n + 8); fR.reserve(n + 8); keep.reserve(n + 8);
    scanL.reserve(n + 8); scanR.reserve(n + 8); scanK.reserve(n + 8);
    ALENS_CUDA(cudaMemsetAsync(c.dCounters.p, 0, sizeof(unsigned long long), st));
    if (n > 0)
        k_migrate_flags<<<gridFor(n, 256), 256, 0, st>>>(n, c.uPos.p, ax, c.slabLo, c.slabHi, c.box.len[ax], c.box.pbc[ax],
                                                         m.left >= 0, m.right >= 0, fL.p, fR.p, keep.p,
                                                         reinterpret_cast<int *>(c.dCounters.p));
    launchScanInt(c, fL.p, scanL.p, n);
    launchScanInt(c, fR.p, scanR.p, n);
    launchScanInt(c, keep.p, scanK.p, n);
    int cnt[3] = {0, 0, 0}, lost = 0;
    ALENS_CUDA(cudaMemcpyAsync(&cnt[0], scanL.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(&cnt[1], scanR.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(&cnt[2], scanK.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaMemcpyAsync(&lost, c.dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    const bool overflow = (size_t)cnt[0] > m.capGhost || (size_t)cnt[1] > m.capGhost;
    if (overflow) cnt[0] = cnt[1] = 0; // stay collective: send nothing, report below
    DevBuf<int> listL, listR;
    listL.reserve((size_t)cnt[0] + 1); listR.reserve((size_t)cnt[1] + 1);
    if (n > 0 && !overflow) {
        k_ghost_list<<<gridFor(n, 256), 256, 0, st>>>(n, fL.p, scanL.p, listL.p);
        k_ghost_list<<<gridFor(n, 256), 256, 0, st>>>(n, fR.p, scanR.p, listR.p);
    }
    const unsigned long long seq = ++m.seqGhost;
    GhostSrc src{c.uGid.p, c.uPos.p, c.uQuat.p, c.uLen.p, c.uRad.p, c.uImm.p, c.uImg.p,
                 c.haveVelNC ? c.uVelNC.p : nullptr, c.globalBase, c.haveTags ? c.uTag.p : nullptr};
    unsigned long long *sf[2] = {nullptr, nullptr};
    long long *sc[2] = {nullptr, nullptr};
    const int *lists[2] = {listL.p, listR.p};
    for (int d = 0; d < 2; d++) {
        unsigned char *w = nbWin(c, d);
        if (!w) continue;
        const int ch = 1 - d;
        if (cnt[d] > 0)
            k_ghost_pack<<<gridFor(cnt[d], 128), 128, 0, st>>>(cnt[d], lists[d], src, 0,
                                                               reinterpret_cast<double *>(w + m.offChan[ch]));
        sf[d] = &hdrOf(w)->chanSeq[ch];
        sc[d] = &hdrOf(w)->chanCount[ch];
    }
    k_signal<<<1, 1, 0, st>>>(sf[0], sc[0], cnt[0], sf[1], sc[1], cnt[1], seq, nullptr);
    k_wait_seq<<<1, 1, 0, st>>>(m.left >= 0 ? &me->chanSeq[0] : nullptr, m.right >= 0 ? &me->chanSeq[1] : nullptr, seq,
                                &me->error, nullptr);
    long long rc[2] = {0, 0};
    ALENS_CUDA(cudaMemcpyAsync(rc, me->chanCount, sizeof(rc), cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    checkCommError(c);
    const int nRecv[2] = {m.left >= 0 ? (int)rc[0] : 0, m.right >= 0 ? (int)rc[1] : 0};
    const int nKeep = overflow ? n : cnt[2];
    const int nNew = nKeep + nRecv[0] + nRecv[1];
    if ((size_t)nNew > m.capRods) throw ArgError{ALENS_ERR_COMM, "comm: rod capacity of the window exceeded"};
    if (!overflow) { // the rods that stay, in their order
        compactInto<int, 1>(c, n, nKeep, nNew, keep.p, scanK.p, c.uGid);
        compactInto<double, 3>(c, n, nKeep, nNew, keep.p, scanK.p, c.uPos);
        compactInto<double, 4>(c, n, nKeep, nNew, keep.p, scanK.p, c.uQuat);
        compactInto<double, 1>(c, n, nKeep, nNew, keep.p, scanK.p, c.uLen);
        compactInto<double, 1>(c, n, nKeep, nNew, keep.p, scanK.p, c.uRad);
        compactInto<unsigned char, 1>(c, n, nKeep, nNew, keep.p, scanK.p, c.uImm);
        if (c.haveVelNC) compactInto<double, 6>(c, n, nKeep, nNew, keep.p, scanK.p, c.uVelNC);
        if (c.haveTags) compactInto<long long, 1>(c, n, nKeep, nNew, keep.p, scanK.p, c.uTag);
    }
    const size_t N = (size_t)nNew;
    c.uGid.reserve(N + 1, st, true, nKeep); c.uPos.reserve(3 * N + 3, st, true, 3 * (size_t)nKeep);
    c.uQuat.reserve(4 * N + 4, st, true, 4 * (size_t)nKeep); c.uLen.reserve(N + 1, st, true, nKeep);
    c.uRad.reserve(N + 1, st, true, nKeep); c.uImm.reserve(N + 1, st, true, nKeep);
    c.uImg.reserve(N + 1); c.uGlobalIdx.reserve(N + 1);
    if (c.haveVelNC) c.uVelNC.reserve(6 * N + 6, st, true, 6 * (size_t)nKeep);
    if (c.haveTags) c.uTag.reserve(N + 1, st, true, nKeep);
    GhostDst dst{c.uGid.p, c.uPos.p, c.uQuat.p, c.uLen.p, c.uRad.p, c.uImm.p, c.uImg.p, c.uGlobalIdx.p,
                 c.haveVelNC ? c.uVelNC.p : nullptr, c.haveTags ? c.uTag.p : nullptr};
    int base = nKeep;
    for (int ch = 0; ch < 2; ch++) {
        if (nRecv[ch] > 0)
            k_ghost_unpack<<<gridFor(nRecv[ch], 128), 128, 0, st>>>(
                nRecv[ch], reinterpret_cast<const double *>(m.win + m.offChan[ch]), dst, base);
        base += nRecv[ch];
    }
    // every rank's new count to every rank; doubles as the barrier behind which the channels may be written again
    const unsigned long long cs = ++m.seqCnt;
    PeerHeaders peers{};
    for (int q = 0; q < c.nranks; q++) peers.h[q] = hdrOf(m.peerWin[q]);
    k_publish_count<<<1, 32, 0, st>>>(peers, c.nranks, c.rank, (long long)nNew, cs);
    k_wait_counts<<<1, 32, 0, st>>>(me, c.nranks, cs);
    long long all[kMaxRanks] = {};
    ALENS_CUDA(cudaMemcpyAsync(all, me->cnt, sizeof(long long) * c.nranks, cudaMemcpyDeviceToHost, st));
    ALENS_CUDA(cudaStreamSynchronize(st));
    checkCommError(c);
    long long before = 0;
    for (int q = 0; q < c.rank; q++) before += all[q];
    c.globalBase = (int)before;
    c.nLocal = c.nRods = nNew;
    c.nGhost = 0;
    c.sorted = false;
    c.haveMob = c.haveSetup = c.haveSolution = false;
    c.launches += 16;
    if (nSent) *nSent = (long long)cnt[0] + cnt[1];
    if (nReceived) *nReceived = (long long)nRecv[0] + nRecv[1];
    ALENS_CUDA(cudaGetLastError());
    if (overflow) throw ArgError{ALENS_ERR_COMM, "alens_migrate_rods: more rods leave than a channel holds (alens_comm_create maxLocalRods too small)"};
    if (lost > 0) throw ArgError{ALENS_ERR_STATE, "alens_migrate_rods: a rod moved further than the neighbouring slab: redistribute on the host"};
}

// velNonCon of the mirrored rods -> the neighbours' user-order vector (ghost rows), staged through the channel
void commHaloVelNC(Context &c) {
    Comm &m = c.comm;
    cudaStream_t st = c.stream;
    CommHeader *me = hdrOf(m.win);
    const unsigned long long seq = ++m.seqVec;
    double *out[2] = {nullptr, nullptr};
    unsigned long long *sf[2] = {nullptr, nullptr};
    for (int d = 0; d < 2; d++) {
        unsigned char *w = nbWin(c, d);
        if (!w) continue;
        out[d] = reinterpret_cast<double *>(w + m.offChan[1 - d]);
        sf[d] = &hdrOf(w)->vecSeq[1 - d];
    }
    const int n0 = out[0] ? m.nSend[0] : 0, n1 = out[1] ? m.nSend[1] : 0;
    if (n0 + n1 > 0)
        k_halo_push6<<<gridFor(3LL * (n0 + n1), 256), 256, 0, st>>>(n0, m.sendIdx[0].p, nullptr, out[0], n1,
                                                                   m.sendIdx[1].p, nullptr, out[1], c.uVelNC.p, nullptr);
    k_signal<<<1, 1, 0, st>>>(sf[0], nullptr, 0, sf[1], nullptr, 0, seq, nullptr);
    k_wait_seq<<<1, 1, 0, st>>>(m.left >= 0 ? &me->vecSeq[0] : nullptr, m.right >= 0 ? &me->vecSeq[1] : nullptr, seq,
                                &me->error, nullptr);
    int base = c.nLocal;
    for (int ch = 0; ch < 2; ch++) {
        if (m.nRecv[ch] > 0)
            k_copy6_rows<<<gridFor(6LL * m.nRecv[ch], 256), 256, 0, st>>>(
                m.nRecv[ch], reinterpret_cast<const double *>(m.win + m.offChan[ch]), c.uVelNC.p + 6 * (size_t)base);
        base += m.nRecv[ch];
    }
    c.launches += 4;
    ALENS_CUDA(cudaGetLastError());
}

// fused path with the dense force kernel (force_kernel = 0): it has stored the mirrored rows; release the sequence
// number on both neighbours (k_force_vel_act does this itself, in its last CTA)
void commSignalHalo(Context &c, unsigned long long seq) {
    unsigned long long *sf[2] = {nullptr, nullptr};
    for (int d = 0; d < 2; d++) {
        unsigned char *w = nbWin(c, d);
        if (w) sf[d] = &hdrOf(w)->haloSeq[1 - d];
    }
    k_signal<<<1, 1, 0, c.stream>>>(sf[0], nullptr, 0, sf[1], nullptr, 0, seq, c.dScal.p);
    c.launches++;
}

// per operator apply: my rows of U -> the ghost rows on the neighbours, then the sequence number
void commPushU(Context &c, unsigned long long seq) {
    Comm &m = c.comm;
    cudaStream_t st = c.stream;
    double *out[2] = {nullptr, nullptr};
    unsigned long long *sf[2] = {nullptr, nullptr};
    const int *ridx[2] = {nullptr, nullptr};
    for (int d = 0; d < 2; d++) {
        unsigned char *w = nbWin(c, d);
        if (!w) continue;
        out[d] = reinterpret_cast<double *>(w + m.offU);
        sf[d] = &hdrOf(w)->haloSeq[1 - d];
        ridx[d] = reinterpret_cast<const int *>(m.win + m.offAck[d]); // written by that neighbour (ack)
    }
    const int n0 = out[0] ? m.nSend[0] : 0, n1 = out[1] ? m.nSend[1] : 0;
    if (n0 + n1 > 0)
        k_halo_push6<<<gridFor(3LL * (n0 + n1), 256), 256, 0, st>>>(n0, m.sendSorted[0].p, ridx[0], out[0], n1,
                                                                   m.sendSorted[1].p, ridx[1], out[1], c.rU.p,
                                                                   c.dScal.p);
    k_signal<<<1, 1, 0, st>>>(sf[0], nullptr, 0, sf[1], nullptr, 0, seq, c.dScal.p);
    CommHeader *me = hdrOf(m.win);
    k_wait_seq<<<1, 1, 0, st>>>(m.left >= 0 ? &me->haloSeq[0] : nullptr, m.right >= 0 ? &me->haloSeq[1] : nullptr, seq,
                                &me->error, c.dScal.p);
    c.launches += 3;
}

// Force the (lazily loaded) kernels of this file into the context now: loading a kernel at its first launch can
// synchronise the context, which deadlocks against a peer rank's waiting kernel when two ranks share one GPU.
void preloadCommKernels() {
    cudaFuncAttributes a;
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_wait_seq));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_signal));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_ghost_flags));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_ghost_list));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_ghost_pack));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_ghost_unpack));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_ack_indices));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_map_indices));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_halo_push6));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_copy6_rows));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_fill_int));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_build_mirror));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_migrate_flags));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_publish_count));
    ALENS_CUDA(cudaFuncGetAttributes(&a, k_wait_counts));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<int, 1>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<double, 1>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<double, 3>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<double, 4>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<double, 6>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<unsigned char, 1>)));
    ALENS_CUDA(cudaFuncGetAttributes(&a, (k_compact<long long, 1>)));
}

} // namespace alens
