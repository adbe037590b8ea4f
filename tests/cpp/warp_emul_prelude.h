// Test harness (not product): a 32-lane warp on 32 host threads, for the warp-cooperative device functions of b2_kernels.cuh (pack_one,
// warp_crc32c_update, warp_snappy_*, k_pack_responses ...).  Every warp collective (__shfl*_sync, __ballot_sync, __syncwarp) is a barrier among
// the lane threads plus an exchange through a per-warp scratch array, so converged code behaves as on the device; full masks only.
#pragma once
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#define __CUDACC__ 1
#define __host__
#define __device__
#define __global__
#define __constant__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
struct we_dim { unsigned x = 0, y = 0, z = 0; };
static thread_local we_dim threadIdx;                 // per lane thread
static we_dim blockIdx, blockDim, gridDim;            // one block
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 v = { a, b, c, d }; return v; }
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 v = { a, b }; return v; }

struct we_warp { pthread_barrier_t bar; uint64_t slot[32]; };
static we_warp g_we_warp;
static inline void we_sync() { pthread_barrier_wait(&g_we_warp.bar); }
static inline unsigned we_lane() { return threadIdx.x & 31u; }
template <typename T> static inline T we_exchange(T v, unsigned src) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    g_we_warp.slot[we_lane()] = bits; we_sync();
    const uint64_t got = g_we_warp.slot[src & 31u]; we_sync();
    T r; memcpy(&r, &got, sizeof(T)); return r;
}
static inline void we_full(unsigned mask) { if (mask != 0xffffffffu) abort(); }
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) { we_full(m); return we_exchange(v, (unsigned)src); }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { we_full(m); const unsigned l = we_lane(); return we_exchange(v, l >= d ? l - d : l); }
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) { we_full(m); const unsigned l = we_lane(); return we_exchange(v, l + d < 32 ? l + d : l); }
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { we_full(m); return we_exchange(v, we_lane() ^ (unsigned)x); }
static inline unsigned __ballot_sync(unsigned m, bool p) {
    we_full(m);
    g_we_warp.slot[we_lane()] = p ? 1 : 0; we_sync();
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= (unsigned)g_we_warp.slot[i] << i;
    we_sync(); return r;
}
static inline void __syncwarp(unsigned m = 0xffffffffu) { we_full(m); we_sync(); }
static inline void __syncthreads() { we_sync(); }     // (a block of one warp)
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicMin(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t, uint32_t sel) { return sel == 0x0123 ? __builtin_bswap32(x) : x; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline unsigned long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}
using std::min; using std::max;

// run fn(lane) on 32 lane threads
template <typename F> static void we_run_warp(F fn) {
    pthread_barrier_init(&g_we_warp.bar, nullptr, 32);
    struct Arg { F* f; unsigned lane; } args[32];
    pthread_t th[32];
    for (unsigned l = 0; l < 32; l++) {
        args[l].f = &fn; args[l].lane = l;
        pthread_create(&th[l], nullptr, [](void* p) -> void* { Arg* a = (Arg*)p; threadIdx.x = a->lane; (*a->f)(a->lane); return nullptr; }, &args[l]);
    }
    for (unsigned l = 0; l < 32; l++) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&g_we_warp.bar);
}
