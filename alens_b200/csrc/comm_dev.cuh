// comm_dev.cuh -- device-side primitives of the peer-memory transport (see comm.cu)
#pragma once

namespace alens {

__device__ __forceinline__ void stReleaseSys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// spin (one thread) until *flag >= want; gives up after ~4 s and raises the error word
__device__ __forceinline__ bool waitSeq(const unsigned long long *flag, unsigned long long want, int *err) {
    const long long t0 = clock64();
    unsigned ns = 32;
    while (ldAcquireSys(flag) < want) {
        __nanosleep(ns);
        if (ns < 1024) ns *= 2;
        if (clock64() - t0 > 8000000000LL) {
            *err = 1;
            return false;
        }
    }
    return true;
}

} // namespace alens
