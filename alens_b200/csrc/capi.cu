// capi.cu -- the extern "C" boundary (include/alens_b200.h).  Every entry point converts exceptions
// into error codes + a message retrievable with alens_last_error(); nothing here computes on the CPU.
#include "context.hpp"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_set>
#include <vector>

using namespace alens;

struct alens_ctx {
    Context c;
};

static thread_local std::string g_createErr;

template <typename F>
static int guarded(alens_ctx *ctx, F &&f) {
    if (!ctx) return ALENS_ERR_ARG;
    try {
        ALENS_CUDA(cudaSetDevice(ctx->c.device));
        g_allocStream = ctx->c.stream; // device allocations of this call are ordered on the context's stream
        g_allocAsync = true;
        f(ctx->c);
        return ALENS_OK;
    } catch (const CudaError &e) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) in `%s` at %s:%d", (int)e.code, cudaGetErrorString(e.code),
                 e.what, e.file, e.line);
        ctx->c.err = buf;
        cudaGetLastError();
        return ALENS_ERR_CUDA;
    } catch (const ArgError &e) {
        ctx->c.err = e.msg;
        return e.code;
    } catch (const std::bad_alloc &) {
        ctx->c.err = "host allocation failed";
        return ALENS_ERR_ARG;
    }
}

// BCQP handles outlive nothing: alens_destroy retires the handles still open on its context (their device memory goes
// with the context); a retired handle answers ALENS_ERR_STATE and alens_bcqp_destroy only frees the shell.
struct alens_bcqp {
    alens_ctx *ctx;
    alens::Bcqp *q;
};
static std::mutex g_bcqpMutex;
static std::unordered_set<alens_bcqp *> g_bcqpOpen;
static void retireBcqpOf(alens_ctx *ctx) {
    std::lock_guard<std::mutex> lock(g_bcqpMutex);
    for (alens_bcqp *p : g_bcqpOpen)
        if (p->ctx == ctx && p->q) {
            alens::bcqpDestroy(p->q);
            p->q = nullptr;
            p->ctx = nullptr;
        }
}
static alens_bcqp *openBcqp(alens_ctx *ctx, alens::Bcqp *q) {
    alens_bcqp *p = new alens_bcqp{ctx, q};
    std::lock_guard<std::mutex> lock(g_bcqpMutex);
    g_bcqpOpen.insert(p);
    return p;
}

// ---- host buffers that are not page-locked (a std::vector of the host application): the runtime's own staging copies them
// at 6-10 GB/s on one thread.  Large transfers go through the context's page-locked bounce buffers in chunks instead: the DMA
// of chunk k+1 overlaps the host copy of chunk k, which runs on all cores.  Page-locked callers (bench.py) keep the direct path.
static bool isPageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
static void hostCopyParallel(void *dst, const void *src, size_t bytes) {
    const size_t blk = (size_t)1 << 18;
    const long long nb = (long long)((bytes + blk - 1) / blk);
#pragma omp parallel for schedule(static)
    for (long long b = 0; b < nb; b++) {
        const size_t o = (size_t)b * blk;
        memcpy((char *)dst + o, (const char *)src + o, std::min(blk, bytes - o));
    }
}
static constexpr size_t kBounceChunk = (size_t)8 << 20;
static constexpr size_t kBounceMin = (size_t)1 << 20; // smaller transfers: not worth it
// device -> host; returns after the data is in `dst` (synchronises the stream)
static void downloadAny(Context &c, void *dst, const void *src, size_t bytes) {
    if (!bytes) return;
    cudaStream_t st = c.stream;
    if (bytes < kBounceMin || !isPageable(dst)) {
        ALENS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        return;
    }
    char *b0 = (char *)c.pinBounce[0].reserve(kBounceChunk), *b1 = (char *)c.pinBounce[1].reserve(kBounceChunk);
    char *buf[2] = {b0, b1};
    const size_t nChunk = (bytes + kBounceChunk - 1) / kBounceChunk;
    auto len = [&](size_t k) { return std::min(kBounceChunk, bytes - k * kBounceChunk); };
    ALENS_CUDA(cudaMemcpyAsync(buf[0], src, len(0), cudaMemcpyDeviceToHost, st));
    for (size_t k = 0; k < nChunk; k++) {
        ALENS_CUDA(cudaStreamSynchronize(st)); // chunk k is in its bounce buffer
        if (k + 1 < nChunk)
            ALENS_CUDA(cudaMemcpyAsync(buf[(k + 1) & 1], (const char *)src + (k + 1) * kBounceChunk, len(k + 1),
                                       cudaMemcpyDeviceToHost, st));
        hostCopyParallel((char *)dst + k * kBounceChunk, buf[k & 1], len(k));
    }
}
// host -> device; the host buffer may be reused when this returns
static void uploadAny(Context &c, void *dst, const void *src, size_t bytes) {
    if (!bytes) return;
    cudaStream_t st = c.stream;
    if (bytes < kBounceMin || !isPageable(src)) {
        ALENS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        return;
    }
    char *b0 = (char *)c.pinBounce[0].reserve(kBounceChunk), *b1 = (char *)c.pinBounce[1].reserve(kBounceChunk);
    char *buf[2] = {b0, b1};
    cudaEvent_t &e0 = c.ev[6], &e1 = c.ev[7];
    cudaEvent_t ev[2] = {e0, e1};
    const size_t nChunk = (bytes + kBounceChunk - 1) / kBounceChunk;
    for (size_t k = 0; k < nChunk; k++) {
        const size_t n = std::min(kBounceChunk, bytes - k * kBounceChunk);
        if (k >= 2) ALENS_CUDA(cudaEventSynchronize(ev[k & 1])); // the DMA that last read this bounce buffer is done
        hostCopyParallel(buf[k & 1], (const char *)src + k * kBounceChunk, n);
        ALENS_CUDA(cudaMemcpyAsync((char *)dst + k * kBounceChunk, buf[k & 1], n, cudaMemcpyHostToDevice, st));
        ALENS_CUDA(cudaEventRecord(ev[k & 1], st));
    }
    ALENS_CUDA(cudaStreamSynchronize(st));
}

static float evMs(Context &c, int a, int b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, c.ev[a], c.ev[b]);
    return ms;
}

extern "C" {

const char *alens_version(void) { return "alens_b200 0.1 (sm_100a)"; }

int alens_create(int device, int rank, int nranks, alens_ctx **out) {
    if (!out) return ALENS_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_createErr = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
        cudaGetLastError();
        return ALENS_ERR_CUDA;
    }
    if (device < 0 || device >= ndev || nranks < 1 || rank < 0 || rank >= nranks) {
        g_createErr = "alens_create: bad device/rank";
        return ALENS_ERR_ARG;
    }
    alens_ctx *ctx = new (std::nothrow) alens_ctx();
    if (!ctx) return ALENS_ERR_ARG;
    ctx->c.device = device;
    ctx->c.rank = rank;
    ctx->c.nranks = nranks;
    int rc = guarded(ctx, [](Context &c) { ctxInit(c); });
    if (rc != ALENS_OK) {
        g_createErr = ctx->c.err;
        delete ctx;
        return rc;
    }
    *out = ctx;
    return ALENS_OK;
}

void alens_destroy(alens_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaDeviceSynchronize();
    retireBcqpOf(ctx);
    commFree(ctx->c);
    g_allocAsync = false; // the stream goes away: remaining buffers are released with cudaFree
    ctxFree(ctx->c);
    delete ctx;
}

const char *alens_last_error(const alens_ctx *ctx) { return ctx ? ctx->c.err.c_str() : g_createErr.c_str(); }

int alens_set_stream(alens_ctx *ctx, void *s) {
    return guarded(ctx, [&](Context &c) {
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
        if (c.ownStream && c.stream) cudaStreamDestroy(c.stream);
        if (s) {
            c.stream = (cudaStream_t)s;
            c.ownStream = false;
        } else {
            ALENS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
            c.ownStream = true;
        }
        g_allocStream = c.stream;
    });
}

int alens_set_domain(alens_ctx *ctx, const double lo[3], const double hi[3], const int pbc[3]) {
    return guarded(ctx, [&](Context &c) {
        for (int k = 0; k < 3; k++) {
            if (!(hi[k] > lo[k])) throw ArgError{ALENS_ERR_ARG, "alens_set_domain: boxHigh must exceed boxLow"};
            c.box.lo[k] = lo[k];
            c.box.hi[k] = hi[k];
            c.box.len[k] = hi[k] - lo[k]; // PS::F64ort::getFullLength
            c.box.pbc[k] = pbc[k] ? 1 : 0;
        }
        c.haveBox = true;
    });
}

int alens_set_collision_params(alens_ctx *ctx, double dRatio, double lRatio, double colBuf) {
    return guarded(ctx, [&](Context &c) {
        if (!(dRatio > 0) || !(lRatio > 0) || !(colBuf >= 0))
            throw ArgError{ALENS_ERR_ARG, "alens_set_collision_params: ratios must be > 0 and colBuf >= 0"};
        c.dRatio = dRatio;
        c.lRatio = lRatio;
        c.colBuf = colBuf;
    });
}

int alens_set_rods(alens_ctx *ctx, int n, const int *gid, const double *pos, const double *quat, const double *len,
                   const double *rad, const unsigned char *imm, int wrap) {
    return guarded(ctx, [&](Context &c) {
        if (!c.haveBox) throw ArgError{ALENS_ERR_STATE, "alens_set_rods: call alens_set_domain first"};
        if (n < 0 || (n > 0 && (!gid || !pos || !quat || !len || !rad)))
            throw ArgError{ALENS_ERR_ARG, "alens_set_rods: null input"};
        cudaStream_t st = c.stream;
        ALENS_CUDA(cudaEventRecord(c.ev[0], st));
        if (n != c.nLocal) c.haveVelNC = false;
        c.haveTags = false;
        c.nLocal = n;
        c.nRods = n;
        const size_t N = (size_t)n;
        c.uGid.reserve(N + 1); c.uPos.reserve(3 * N + 3); c.uQuat.reserve(4 * N + 4);
        c.uLen.reserve(N + 1); c.uRad.reserve(N + 1); c.uImm.reserve(N + 1);
        if (n > 0) {
            ALENS_CUDA(cudaMemcpyAsync(c.uGid.p, gid, 4 * N, cudaMemcpyHostToDevice, st));
            ALENS_CUDA(cudaMemcpyAsync(c.uPos.p, pos, 24 * N, cudaMemcpyHostToDevice, st));
            ALENS_CUDA(cudaMemcpyAsync(c.uQuat.p, quat, 32 * N, cudaMemcpyHostToDevice, st));
            ALENS_CUDA(cudaMemcpyAsync(c.uLen.p, len, 8 * N, cudaMemcpyHostToDevice, st));
            ALENS_CUDA(cudaMemcpyAsync(c.uRad.p, rad, 8 * N, cudaMemcpyHostToDevice, st));
            if (imm) ALENS_CUDA(cudaMemcpyAsync(c.uImm.p, imm, N, cudaMemcpyHostToDevice, st));
            else ALENS_CUDA(cudaMemsetAsync(c.uImm.p, 0, N, st));
        }
        c.maxRLocal = hostMaxRadius(n, len, rad, c.lRatio, c.dRatio, &c.meanRLocal); // overlaps with the copies
        c.maxRLRatio = c.lRatio;
        c.maxRDRatio = c.dRatio;
        rodsUploaded(c, wrap != 0);
        ALENS_CUDA(cudaEventRecord(c.ev[1], st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        c.timers.upload_ms = evMs(c, 0, 1);
    });
}

int alens_set_rod_state(alens_ctx *ctx, const double *pos, const double *quat, int wrap) {
    return guarded(ctx, [&](Context &c) {
        if (!c.haveBox || !c.uPos.p) throw ArgError{ALENS_ERR_STATE, "alens_set_rod_state: call alens_set_rods first"};
        if (c.nLocal > 0 && (!pos || !quat)) throw ArgError{ALENS_ERR_ARG, "alens_set_rod_state: null input"};
        cudaStream_t st = c.stream;
        ALENS_CUDA(cudaEventRecord(c.ev[0], st));
        const size_t N = (size_t)c.nLocal;
        if (N > 0) {
            ALENS_CUDA(cudaMemcpyAsync(c.uPos.p, pos, 24 * N, cudaMemcpyHostToDevice, st));
            ALENS_CUDA(cudaMemcpyAsync(c.uQuat.p, quat, 32 * N, cudaMemcpyHostToDevice, st));
        }
        rodsUploaded(c, wrap != 0);
        ALENS_CUDA(cudaEventRecord(c.ev[1], st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        c.timers.upload_ms = evMs(c, 0, 1);
    });
}

int alens_set_rods_aos(alens_ctx *ctx, int n, const void *sy, size_t stride, int wrap) {
    if (!ctx) return ALENS_ERR_ARG;
    if (n < 0 || (n > 0 && !sy) || stride < 136) {
        ctx->c.err = "alens_set_rods_aos: bad arguments";
        return ALENS_ERR_ARG;
    }
    // field offsets of the reference `Sylinder` record (Sylinder.hpp:38-57): gid 0, group 12, isImmovable 16,
    // radius 24, length 40, pos 80, orientation 104.  The 85 hot bytes of every 568-byte record are gathered by all host
    // cores into ONE page-locked staging buffer (grow-only, kept by the context), from where the copies run at full
    // PCIe speed; Sylinder::group rides along as the rod's tag (it stays with the rod when the rod migrates).
    const size_t N = (size_t)n;
    auto al = [](size_t x) { return (x + 63) & ~(size_t)63; };
    const size_t oPos = 0, oQ = oPos + al(24 * N), oLen = oQ + al(32 * N), oRad = oLen + al(8 * N), oTag = oRad + al(8 * N),
                 oGid = oTag + al(8 * N), oImm = oGid + al(4 * N), total = oImm + al(N) + 64;
    char *stage = nullptr;
    try {
        stage = (char *)ctx->c.pinAos.reserve(total);
    } catch (const CudaError &) {
        cudaGetLastError();
        ctx->c.err = "alens_set_rods_aos: cannot allocate the page-locked staging buffer";
        return ALENS_ERR_CUDA;
    }
    double *pos = (double *)(stage + oPos), *q = (double *)(stage + oQ), *len = (double *)(stage + oLen),
           *rad = (double *)(stage + oRad);
    long long *tag = (long long *)(stage + oTag);
    int *gid = (int *)(stage + oGid);
    unsigned char *imm = (unsigned char *)(stage + oImm);
    const char *base = (const char *)sy;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const char *p = base + (size_t)i * stride;
        int g;
        memcpy(&gid[i], p + 0, 4);
        memcpy(&g, p + 12, 4);
        tag[i] = g;
        imm[i] = *(const unsigned char *)(p + 16);
        memcpy(&rad[i], p + 24, 8);
        memcpy(&len[i], p + 40, 8);
        memcpy(&pos[3 * (size_t)i], p + 80, 24);
        memcpy(&q[4 * (size_t)i], p + 104, 32);
    }
    const int rc = alens_set_rods(ctx, n, gid, pos, q, len, rad, imm, wrap);
    return rc != ALENS_OK ? rc : alens_set_rod_tags(ctx, tag);
}

int alens_set_rod_tags(alens_ctx *ctx, const long long *tags) {
    return guarded(ctx, [&](Context &c) {
        c.haveTags = tags != nullptr;
        if (!tags) return;
        c.uTag.reserve((size_t)c.nLocal + 1);
        if (c.nLocal > 0) ALENS_CUDA(cudaMemcpyAsync(c.uTag.p, tags, 8 * (size_t)c.nLocal, cudaMemcpyHostToDevice, c.stream));
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int alens_get_rod_tags(alens_ctx *ctx, long long *tags) {
    return guarded(ctx, [&](Context &c) {
        if (!tags) throw ArgError{ALENS_ERR_ARG, "alens_get_rod_tags: NULL output"};
        if (c.nLocal == 0) return;
        if (!c.haveTags) {
            memset(tags, 0, 8 * (size_t)c.nLocal);
            return;
        }
        ALENS_CUDA(cudaMemcpyAsync(tags, c.uTag.p, 8 * (size_t)c.nLocal, cudaMemcpyDeviceToHost, c.stream));
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int alens_prepare_step(alens_ctx *ctx, int wrap) {
    return guarded(ctx, [&](Context &c) {
        if (!c.haveBox || c.uPos.p == nullptr) throw ArgError{ALENS_ERR_STATE, "alens_prepare_step: no resident rods"};
        ALENS_CUDA(cudaEventRecord(c.ev[0], c.stream));
        rodsUploaded(c, wrap != 0);
        ALENS_CUDA(cudaEventRecord(c.ev[1], c.stream));
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
        c.timers.upload_ms = evMs(c, 0, 1);
    });
}

int alens_set_velocity_noncon(alens_ctx *ctx, const double *v) {
    return guarded(ctx, [&](Context &c) {
        if (!v) {
            c.haveVelNC = false;
            return;
        }
        c.uVelNC.reserve(6 * (size_t)c.nRods + 6);
        if (c.nLocal > 0) uploadAny(c, c.uVelNC.p, v, 48 * (size_t)c.nLocal);
        c.haveVelNC = true;
    });
}

int alens_calc_velocity_brown(alens_ctx *ctx, double kBT, double dt, const double *normals12, unsigned long long seed,
                              unsigned long long step, double *velBrownOut) {
    return guarded(ctx, [&](Context &c) { calcVelocityBrown(c, kBT, dt, normals12, seed, step, velBrownOut); });
}

int alens_calc_velocity_noncon(alens_ctx *ctx, const double *forceNonBrown, const double *velocityNonBrown,
                               const double *velocityBrown, int monolayer, double *velNonBOut) {
    return guarded(ctx, [&](Context &c) {
        calcVelocityNonCon(c, forceNonBrown, velocityNonBrown, velocityBrown, monolayer, velNonBOut);
    });
}

int alens_set_velocity_noncon_async(alens_ctx *ctx, const double *v) {
    return guarded(ctx, [&](Context &c) {
        if (!v) {
            c.haveVelNC = false;
            return;
        }
        if (!c.copyStream) {
            ALENS_CUDA(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
            ALENS_CUDA(cudaEventCreateWithFlags(&c.evVelNC, cudaEventDisableTiming));
            ALENS_CUDA(cudaEventCreateWithFlags(&c.evMain, cudaEventDisableTiming));
        }
        c.uVelNC.reserve(6 * (size_t)c.nRods + 6);
        // the side stream starts behind everything the main stream has been given so far (allocation, last readers)
        ALENS_CUDA(cudaEventRecord(c.evMain, c.stream));
        ALENS_CUDA(cudaStreamWaitEvent(c.copyStream, c.evMain, 0));
        if (c.nLocal > 0)
            ALENS_CUDA(cudaMemcpyAsync(c.uVelNC.p, v, 48 * (size_t)c.nLocal, cudaMemcpyHostToDevice, c.copyStream));
        ALENS_CUDA(cudaEventRecord(c.evVelNC, c.copyStream));
        c.velNCPending = true;
        c.haveVelNC = true;
    });
}

int alens_set_profiling(alens_ctx *ctx, int on) {
    return guarded(ctx, [&](Context &c) { c.profiling = on != 0; });
}

int alens_get_positions(alens_ctx *ctx, double *pos) {
    return guarded(ctx, [&](Context &c) {
        if (c.nLocal > 0) downloadAny(c, pos, c.uPos.p, 24 * (size_t)c.nLocal);
    });
}

int alens_get_rod_state(alens_ctx *ctx, double *pos, double *quat) {
    return guarded(ctx, [&](Context &c) {
        if (c.nLocal > 0) {
            if (pos) downloadAny(c, pos, c.uPos.p, 24 * (size_t)c.nLocal);
            if (quat) downloadAny(c, quat, c.uQuat.p, 32 * (size_t)c.nLocal);
        }
    });
}

int alens_migrate_rods(alens_ctx *ctx, long long *nSent, long long *nReceived) {
    return guarded(ctx, [&](Context &c) { commMigrate(c, nSent, nReceived); });
}

int alens_get_rod_identity(alens_ctx *ctx, int *nLocal, int *globalIndexBase, int *gid, double *length, double *radius,
                           unsigned char *immovable) {
    return guarded(ctx, [&](Context &c) {
        if (nLocal) *nLocal = c.nLocal;
        if (globalIndexBase) *globalIndexBase = c.globalBase;
        const size_t n = (size_t)c.nLocal;
        if (n == 0) return;
        if (gid) ALENS_CUDA(cudaMemcpyAsync(gid, c.uGid.p, 4 * n, cudaMemcpyDeviceToHost, c.stream));
        if (length) ALENS_CUDA(cudaMemcpyAsync(length, c.uLen.p, 8 * n, cudaMemcpyDeviceToHost, c.stream));
        if (radius) ALENS_CUDA(cudaMemcpyAsync(radius, c.uRad.p, 8 * n, cudaMemcpyDeviceToHost, c.stream));
        if (immovable) ALENS_CUDA(cudaMemcpyAsync(immovable, c.uImm.p, n, cudaMemcpyDeviceToHost, c.stream));
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
    });
}

int alens_collect_pair_collision(alens_ctx *ctx, long long *nCon) {
    return guarded(ctx, [&](Context &c) {
        ALENS_CUDA(cudaEventRecord(c.ev[0], c.stream));
        collectPairs(c);
        ALENS_CUDA(cudaEventRecord(c.ev[1], c.stream));
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
        c.timers.collect_ms = evMs(c, 0, 1);
        if (nCon) *nCon = c.nCon;
    });
}

int alens_append_constraints(alens_ctx *ctx, const alens_constraint_block *b, long long n) {
    return guarded(ctx, [&](Context &c) {
        if (n < 0 || (n > 0 && !b)) throw ArgError{ALENS_ERR_ARG, "alens_append_constraints: bad arguments"};
        appendBlocks(c, b, n);
    });
}

int alens_collect_boundary_collision(alens_ctx *ctx, const alens_boundary *boundaries, int nBoundaries, long long *nAdded) {
    return guarded(ctx, [&](Context &c) {
        if (nBoundaries > 0 && !boundaries) throw ArgError{ALENS_ERR_ARG, "alens_collect_boundary_collision: NULL boundaries"};
        const long long n = collectBoundary(c, boundaries, nBoundaries);
        if (nAdded) *nAdded = n;
    });
}

int alens_collect_link_bilateral(alens_ctx *ctx, const int *prevGid, const int *nextGid, long long nLinks, double linkKappa,
                                 double linkGap, long long *nAdded) {
    return guarded(ctx, [&](Context &c) {
        const long long n = collectLinks(c, prevGid, nextGid, nLinks, linkKappa, linkGap);
        if (nAdded) *nAdded = n;
    });
}

int alens_collect_protein_bilateral(alens_ctx *ctx, const alens_protein_bind *proteins, long long n, double tubuleDiameter,
                                    long long *nAdded) {
    return guarded(ctx, [&](Context &c) {
        const long long m = collectProteins(c, proteins, n, tubuleDiameter);
        if (nAdded) *nAdded = m;
    });
}

int alens_clear_constraints(alens_ctx *ctx) {
    return guarded(ctx, [&](Context &c) {
        c.nCon = c.nColl = 0;
        c.nOneSide = c.nBilateral = 0;
        c.hostBlocks.clear();
        c.haveSetup = false;
        c.haveSolution = false;
    });
}

int alens_num_constraints(alens_ctx *ctx, long long *n) {
    return guarded(ctx, [&](Context &c) {
        if (n) *n = c.nCon;
    });
}

int alens_get_constraints(alens_ctx *ctx, alens_constraint_block *out, long long cap, int withStress, int writeBack) {
    return guarded(ctx, [&](Context &c) { downloadBlocks(c, out, cap, withStress != 0, writeBack != 0); });
}

int alens_calc_mobility(alens_ctx *ctx, double mu) {
    return guarded(ctx, [&](Context &c) { calcMobility(c, mu); });
}

int alens_mobility_apply(alens_ctx *ctx, const double *x, double *y) {
    return guarded(ctx, [&](Context &c) { mobilityApply(c, x, y); });
}

int alens_setup_constraints(alens_ctx *ctx, const double *velNC, double dt) {
    return guarded(ctx, [&](Context &c) {
        ALENS_CUDA(cudaEventRecord(c.ev[0], c.stream));
        setupConstraints(c, velNC, dt);
        ALENS_CUDA(cudaEventRecord(c.ev[1], c.stream));
        ALENS_CUDA(cudaStreamSynchronize(c.stream));
        c.timers.setup_ms = evMs(c, 0, 1);
    });
}

int alens_solve_constraints(alens_ctx *ctx, const double *velNC, double dt, double res, int maxIte, int choice,
                            alens_solve_report *rep) {
    int rc = alens_setup_constraints(ctx, velNC, dt);
    if (rc != ALENS_OK) return rc;
    rc = guarded(ctx, [&](Context &c) {
        if (maxIte < 0) throw ArgError{ALENS_ERR_ARG, "alens_solve_constraints: maxIte < 0"};
        solveConstraints(c, res, maxIte, choice);
    });
    if (rep && ctx) *rep = ctx->c.lastReport;
    return rc;
}

int alens_bcqp_solve(alens_ctx *ctx, const double *b, double *x, double tol, int maxIte, int choice,
                     alens_solve_report *rep) {
    int rc = guarded(ctx, [&](Context &c) {
        if (!c.haveSetup) throw ArgError{ALENS_ERR_STATE, "alens_bcqp_solve: call alens_setup_constraints first"};
        if (maxIte < 0 || !x) throw ArgError{ALENS_ERR_ARG, "alens_bcqp_solve: bad arguments"};
        const size_t bytes = 8 * (size_t)c.nCon;
        if (bytes) {
            if (b) ALENS_CUDA(cudaMemcpyAsync(c.vB.p, b, bytes, cudaMemcpyHostToDevice, c.stream));
            ALENS_CUDA(cudaMemcpyAsync(c.vX0.p, x, bytes, cudaMemcpyHostToDevice, c.stream));
        }
        solveCore(c, tol, maxIte, choice);
        if (bytes) {
            ALENS_CUDA(cudaMemcpyAsync(x, c.xSolution, bytes, cudaMemcpyDeviceToHost, c.stream));
            ALENS_CUDA(cudaStreamSynchronize(c.stream));
        }
    });
    if (rep && ctx) *rep = ctx->c.lastReport;
    return rc;
}

int alens_operator_apply(alens_ctx *ctx, const double *x, double *y, double *force, double *vel) {
    return guarded(ctx, [&](Context &c) { operatorApply(c, x, y, force, vel); });
}

int alens_get_history(alens_ctx *ctx, double *rows, int cap, int *nRows) {
    return guarded(ctx, [&](Context &c) {
        const int have = (int)(c.hist.size() / 6);
        const int n = have < cap ? have : cap;
        if (rows && n > 0) memcpy(rows, c.hist.data(), 48 * (size_t)n);
        if (nRows) *nRows = have;
    });
}

int alens_get_gamma(alens_ctx *ctx, double *gamma, long long cap) {
    return guarded(ctx, [&](Context &c) {
        if (!c.haveSolution) throw ArgError{ALENS_ERR_STATE, "alens_get_gamma: no solution available"};
        if (cap < c.nCon) throw ArgError{ALENS_ERR_ARG, "alens_get_gamma: capacity too small"};
        if (c.nCon > 0) {
            ALENS_CUDA(cudaMemcpyAsync(gamma, c.xSolution, 8 * (size_t)c.nCon, cudaMemcpyDeviceToHost, c.stream));
            ALENS_CUDA(cudaStreamSynchronize(c.stream));
        }
    });
}

int alens_get_force_velocity(alens_ctx *ctx, double *fU, double *vU, double *fB, double *vB) {
    return guarded(ctx, [&](Context &c) {
        if (!c.haveSolution) throw ArgError{ALENS_ERR_STATE, "alens_get_force_velocity: no solution available"};
        const size_t bytes = 48 * (size_t)c.nLocal;
        cudaStream_t st = c.stream;
        ALENS_CUDA(cudaEventRecord(c.ev[0], st));
        if (bytes) {
            double *dst[4] = {fU, vU, fB, vB};
            const double *src[4] = {c.outFU.p, c.outVU.p, c.outFB.p, c.outVB.p};
            bool anyPageable = false;
            for (int k = 0; k < 4; k++) anyPageable = anyPageable || (dst[k] && bytes >= kBounceMin && isPageable(dst[k]));
            for (int k = 0; k < 4; k++) {
                if (!dst[k]) continue;
                if (anyPageable) downloadAny(c, dst[k], src[k], bytes);
                else ALENS_CUDA(cudaMemcpyAsync(dst[k], src[k], bytes, cudaMemcpyDeviceToHost, st));
            }
        }
        ALENS_CUDA(cudaEventRecord(c.ev[1], st));
        ALENS_CUDA(cudaStreamSynchronize(st));
        c.timers.download_ms = evMs(c, 0, 1);
    });
}

int alens_step_euler(alens_ctx *ctx, double dt) {
    return guarded(ctx, [&](Context &c) { stepEuler(c, dt); });
}

int alens_get_timers(alens_ctx *ctx, alens_timers *t) {
    return guarded(ctx, [&](Context &c) {
        c.timers.total_launches = c.launches;
        if (t) *t = c.timers;
    });
}

int alens_reset_timers(alens_ctx *ctx) {
    return guarded(ctx, [&](Context &c) {
        memset(&c.timers, 0, sizeof(c.timers));
        c.launches = 0;
    });
}

int alens_set_option(alens_ctx *ctx, const char *name, long long value) {
    return guarded(ctx, [&](Context &c) {
        const std::string k = name ? name : "";
        if (k == "force_kernel") { // 3 k_force_vel_rec (slot records), 0 dense level-major, 1 k_force_vel_act, 2 k_slot_x + k_rod_sum (rod-major layout)
            c.optForceKernel = value == 0 ? 0 : (value == 3 ? 3 : 1);
            c.optForceSplit = value == 2;
            if (c.optForceKernel == 3) c.optTailRing = 0; // the TMA-staged tail does not maintain the slot records
        }
        else if (k == "rec_mode") c.optRecMode = (value < 0 || value > 2) ? 1 : (int)value; // 2: records hold M * column
        else if (k == "stamps") c.optStamps = value != 0;
        else if (k == "halo_debug") c.optHaloDebug = (int)value;
        else if (k == "tail_push") c.optTailPush = value != 0;
        else if (k == "long_rods") c.optLongRods = value <= 0 ? 0.0 : value / 100.0; // percent of the mean bounding radius (200 = default)
        else if (k == "l2_fetch") { // cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes): a device-wide hint
            ALENS_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
        }
        else if (k == "find_minb") c.optFindMinB = (value == 5 || value == 3) ? (int)value : 4;
        else if (k == "find_split") c.optFindSplit = value != 0;
        else if (k == "find_split_minb") c.optFindSplitMinB = value == 6 ? 6 : 8;
        else if (k == "force_minb") c.optForceMinB = (value == 3 || value == 5) ? (int)value : 4;
        else if (k == "tail_ring") {
            c.optTailRing = value != 0;
            if (c.optTailRing && c.optForceKernel == 3) c.optForceKernel = 1; // see force_kernel
        }
        else if (k == "pdl") c.optPdl = (int)std::max(0LL, std::min(2LL, value));
        else if (k == "poll") c.optPoll = value != 0;
        else if (k == "lookahead") c.optLookahead = (int)std::max(1LL, std::min(64LL, value));
        else if (k == "late_halo") c.optLateHalo = value != 0;
        else if (k == "keep_xg") c.optKeepXG = value != 0;
        else if (k == "force_mask") c.optForceMask = value != 0;
        else if (k == "force_waves") c.optForceWaves = (int)std::max(1LL, std::min(64LL, value));
        else if (k == "force_chunk") c.optForceChunk = value == 4 ? 4 : 2;
        else if (k == "force_block") c.optForceBlock = value == 32 ? 32 : (value == 64 ? 64 : (value == 128 ? 128 : 256));
        else if (k == "tail_ctas_per_sm") c.optTailCtasPerSM = (int)std::max(1LL, std::min(8LL, value));
        else if (k == "comm_fused") c.comm.fused = value != 0;
        else if (k == "bbpgd_batch") c.optBatch = (int)std::max(0LL, std::min(1024LL, value));
        else throw ArgError{ALENS_ERR_ARG, "alens_set_option: unknown option '" + k + "'"};
    });
}

int alens_time_kernel(alens_ctx *ctx, const char *which, int reps, double *avgMicroseconds) {
    return guarded(ctx, [&](Context &c) {
        const std::string k = which ? which : "";
        const int w = k == "force_vel" ? 0 : (k == "tail" ? 1 : (k == "force_vel_plain" ? 2 : (k == "force_vel_last" ? 3 : -1)));
        if (w < 0 || reps < 1) throw ArgError{ALENS_ERR_ARG, "alens_time_kernel: which = force_vel | tail | force_vel_plain"};
        const double us = timeKernel(c, w, reps);
        if (avgMicroseconds) *avgMicroseconds = us;
    });
}

int alens_get_collect_stats(alens_ctx *ctx, long long *nCells, long long *nCand, long long *nHits) {
    return guarded(ctx, [&](Context &c) {
        if (nCells) *nCells = c.grid.ncell;
        if (nCand) *nCand = c.statCand;
        if (nHits) *nHits = c.nColl;
    });
}

int alens_get_long_rod_stats(alens_ctx *ctx, long long *nLongRods, long long *nLongRows, double *shortRadius, double *maxRadius) {
    return guarded(ctx, [&](Context &c) {
        if (nLongRods) *nLongRods = c.nLongRods;
        if (nLongRows) *nLongRows = c.nLongRows;
        if (shortRadius) *shortRadius = c.shortR;
        if (maxRadius) *maxRadius = c.gridMaxR;
    });
}

int alens_mix_pair_search(alens_ctx *ctx, long long nTargets, const double *targetPos, const double *targetRSearch,
                          const double *sourceRSearch, long long *rowPtr, int *sourceIndex, long long capPairs,
                          long long *nPairs) {
    return guarded(ctx, [&](Context &c) {
        const long long m = mixPairSearch(c, nTargets, targetPos, targetRSearch, sourceRSearch, rowPtr, sourceIndex, capPairs);
        if (nPairs) *nPairs = m;
    });
}

int alens_dcp_query(alens_ctx *ctx, long long n, const double *P0, const double *P1, const double *Q0, const double *Q1,
                    double *dist, double *Ploc, double *Qloc) {
    return guarded(ctx, [&](Context &c) {
        if (n < 0 || (n > 0 && (!P0 || !P1 || !Q0 || !Q1))) throw ArgError{ALENS_ERR_ARG, "alens_dcp_query: null input"};
        dcpBatch(c, n, P0, P1, Q0, Q1, dist, Ploc, Qloc);
    });
}

int alens_pair_functor(alens_ctx *ctx, long long n, const double *geomI, const double *geomJ, int withStress,
                       unsigned char *hit, alens_constraint_block *blocks) {
    return guarded(ctx, [&](Context &c) {
        if (n < 0 || (n > 0 && (!geomI || !geomJ || !hit || !blocks)))
            throw ArgError{ALENS_ERR_ARG, "alens_pair_functor: null argument"};
        pairFunctorBatch(c, n, geomI, geomJ, withStress, hit, blocks);
    });
}

/* ---- multi-GPU ---------------------------------------------------------------------------------- */
int alens_set_decomposition(alens_ctx *ctx, int axis, double slabLow, double slabHigh, double skin,
                            double maxBoundingRadius, int globalIndexBase) {
    return guarded(ctx, [&](Context &c) {
        if (axis < 0 || axis > 2 || !(slabHigh > slabLow) || !(skin >= 0) || !(maxBoundingRadius >= 0))
            throw ArgError{ALENS_ERR_ARG, "alens_set_decomposition: bad arguments"};
        c.slabAxis = axis;
        c.slabLo = slabLow;
        c.slabHi = slabHigh;
        c.skin = skin;
        c.maxRadiusGlobal = maxBoundingRadius;
        c.globalBase = globalIndexBase;
    });
}

int alens_comm_create(alens_ctx *ctx, long long maxLocalRods) {
    return guarded(ctx, [&](Context &c) {
        if (maxLocalRods < 1) throw ArgError{ALENS_ERR_ARG, "alens_comm_create: maxLocalRods < 1"};
        commAllocWindow(c, maxLocalRods);
    });
}

int alens_comm_blob_size(void) { return (int)sizeof(CommBlob); }

int alens_comm_export(alens_ctx *ctx, void *blob) {
    return guarded(ctx, [&](Context &c) {
        if (!c.comm.win || !blob) throw ArgError{ALENS_ERR_STATE, "alens_comm_export: call alens_comm_create first"};
        commExport(c, blob);
    });
}

int alens_comm_connect(alens_ctx *ctx, const void *blobsInRankOrder) {
    return guarded(ctx, [&](Context &c) {
        if (!c.comm.win || !blobsInRankOrder) throw ArgError{ALENS_ERR_STATE, "alens_comm_connect: call alens_comm_create first"};
        if (!c.haveBox) throw ArgError{ALENS_ERR_STATE, "alens_comm_connect: call alens_set_domain first"};
        commImport(c, blobsInRankOrder);
    });
}

int alens_comm_connect_local(alens_ctx **ctxs, int n) {
    if (!ctxs || n < 1) return ALENS_ERR_ARG;
    std::vector<Context *> v;
    for (int i = 0; i < n; i++) {
        if (!ctxs[i] || !ctxs[i]->c.comm.win || !ctxs[i]->c.haveBox) return ALENS_ERR_STATE;
        v.push_back(&ctxs[i]->c);
    }
    try {
        commConnectLocal(v.data(), n);
        return ALENS_OK;
    } catch (const CudaError &e) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) in `%s` at %s:%d", (int)e.code, cudaGetErrorString(e.code),
                 e.what, e.file, e.line);
        ctxs[0]->c.err = buf;
        cudaGetLastError();
        return ALENS_ERR_CUDA;
    } catch (const ArgError &e) {
        ctxs[0]->c.err = e.msg;
        return e.code;
    }
}

int alens_num_ghosts(alens_ctx *ctx, int *nGhost, int *nSentLeft, int *nSentRight) {
    return guarded(ctx, [&](Context &c) {
        if (nGhost) *nGhost = c.nGhost;
        if (nSentLeft) *nSentLeft = c.comm.nSend[0];
        if (nSentRight) *nSentRight = c.comm.nSend[1];
    });
}

/* ---- BCQPSolver for any caller ------------------------------------------------------------------ */
int alens_bcqp_create_csr(alens_ctx *ctx, int n, const long long *rowPtr, const int *colInd, const double *values,
                          const double *b, alens_bcqp **out) {
    if (!out) return ALENS_ERR_ARG;
    *out = nullptr;
    return guarded(ctx, [&](Context &c) {
        if (!rowPtr) throw ArgError{ALENS_ERR_ARG, "alens_bcqp_create_csr: null matrix"};
        Bcqp *q = bcqpCreate(c, n, rowPtr, colInd, values, b);
        *out = openBcqp(ctx, q);
    });
}
int alens_bcqp_create_constraint(alens_ctx *ctx, const double *b, alens_bcqp **out) {
    if (!out) return ALENS_ERR_ARG;
    *out = nullptr;
    return guarded(ctx, [&](Context &c) {
        Bcqp *q = bcqpCreate(c, 0, nullptr, nullptr, nullptr, b);
        *out = openBcqp(ctx, q);
    });
}
int alens_bcqp_set_lower_bound(alens_bcqp *p, const double *lb) {
    if (!p) return ALENS_ERR_ARG;
    if (!p->q) return ALENS_ERR_STATE; // its context is gone
    return guarded(p->ctx, [&](Context &) { bcqpSetBounds(*p->q, lb, nullptr, 1); });
}
int alens_bcqp_set_upper_bound(alens_bcqp *p, const double *ub) {
    if (!p) return ALENS_ERR_ARG;
    if (!p->q) return ALENS_ERR_STATE; // its context is gone
    return guarded(p->ctx, [&](Context &) { bcqpSetBounds(*p->q, nullptr, ub, 2); });
}
int alens_bcqp_get_bounds(alens_bcqp *p, double *lb, double *ub) {
    if (!p) return ALENS_ERR_ARG;
    if (!p->q) return ALENS_ERR_STATE; // its context is gone
    return guarded(p->ctx, [&](Context &) { bcqpGetBounds(*p->q, lb, ub); });
}
int alens_bcqp_run(alens_bcqp *p, double *x, double tol, int maxIte, int solverChoice, alens_solve_report *report) {
    if (!p) return ALENS_ERR_ARG;
    if (!p->q) return ALENS_ERR_STATE; // its context is gone
    return guarded(p->ctx, [&](Context &) { bcqpSolve(*p->q, x, tol, maxIte, solverChoice, report); });
}
int alens_bcqp_history(alens_bcqp *p, double *rows6, int capRows, int *nRows) {
    if (!p) return ALENS_ERR_ARG;
    if (!p->q) return ALENS_ERR_STATE; // its context is gone
    return guarded(p->ctx, [&](Context &) {
        const int n = bcqpHistory(*p->q, rows6, capRows);
        if (nRows) *nRows = n;
    });
}
int alens_bcqp_size(alens_bcqp *p) { return (p && p->q) ? bcqpSize(*p->q) : 0; }
void alens_bcqp_destroy(alens_bcqp *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_bcqpMutex);
        g_bcqpOpen.erase(p);
    }
    if (p->q) guarded(p->ctx, [&](Context &) { bcqpDestroy(p->q); });
    delete p;
}

int alens_get_stamps(alens_ctx *ctx, unsigned long long *stamps8, int capIterations, int *nIterations) {
    return guarded(ctx, [&](Context &c) {
        const int n = std::min(c.stampIters, capIterations);
        if (n > 0 && stamps8)
            ALENS_CUDA(cudaMemcpy(stamps8, c.dStamps.p, 64 * (size_t)n, cudaMemcpyDeviceToHost));
        if (nIterations) *nIterations = c.stampIters;
    });
}

int alens_get_pool_stats(alens_ctx *ctx, long long *nCollision, long long *nOneSide, long long *nBilateral) {
    return guarded(ctx, [&](Context &c) {
        if (nCollision) *nCollision = c.nColl;
        if (nOneSide) *nOneSide = c.nOneSide;
        if (nBilateral) *nBilateral = c.nBilateral;
    });
}

int alens_get_live_stats(alens_ctx *ctx, long long *liveSlots, long long *liveRods) {
    return guarded(ctx, [&](Context &c) {
        long long a = 0, b = 0;
        liveStats(c, &a, &b);
        if (liveSlots) *liveSlots = a;
        if (liveRods) *liveRods = b;
    });
}

int alens_constraint_digest(alens_ctx *ctx, unsigned long long counts3[3], double sums3[3]) {
    return guarded(ctx, [&](Context &c) {
        if (!counts3 || !sums3) throw ArgError{ALENS_ERR_ARG, "alens_constraint_digest: null output"};
        constraintDigest(c, counts3, sums3);
    });
}

int alens_sum_constraint_stress(alens_ctx *ctx, int withOneSide, double uniStress[9], double biStress[9]) {
    return guarded(ctx, [&](Context &c) {
        if (!uniStress || !biStress) throw ArgError{ALENS_ERR_ARG, "alens_sum_constraint_stress: null output"};
        sumConstraintStress(c, withOneSide != 0, uniStress, biStress);
    });
}

int alens_comm_mode(alens_ctx *ctx, int *connected, int *fused) {
    return guarded(ctx, [&](Context &c) {
        if (connected) *connected = c.comm.active ? 1 : 0;
        if (fused) *fused = (c.comm.active && c.comm.fused) ? 1 : 0;
    });
}

} // extern "C"
