// Test harness (not product): a thread BLOCK on host threads — warps of 32 lanes (collectives = barrier among the warp's lane threads + exchange)
// and __syncthreads() = a barrier among all threads of the block.  Lets a whole single-CTA kernel of b2_kernels.cuh (k_small: the latency
// path end to end) run on the CPU.  Full warp masks only; one block at a time.
#pragma once
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
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
struct be_dim { unsigned x = 0, y = 0, z = 0; };
static thread_local be_dim threadIdx;
static be_dim blockIdx, blockDim, gridDim;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 v = { a, b, c, d }; return v; }
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 v = { a, b }; return v; }

struct be_warp { pthread_barrier_t bar; uint64_t slot[32]; };
static be_warp g_be_warps[32];
static pthread_barrier_t g_be_block;
static inline be_warp& be_w() { return g_be_warps[threadIdx.x >> 5]; }
static inline unsigned be_lane() { return threadIdx.x & 31u; }
static inline void be_full(unsigned mask) { if (mask != 0xffffffffu) abort(); }
template <typename T> static inline T be_exchange(T v, unsigned src) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
    be_warp& w = be_w();
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    w.slot[be_lane()] = bits; pthread_barrier_wait(&w.bar);
    const uint64_t got = w.slot[src & 31u]; pthread_barrier_wait(&w.bar);
    T r; memcpy(&r, &got, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) { be_full(m); return be_exchange(v, (unsigned)src); }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { be_full(m); const unsigned l = be_lane(); return be_exchange(v, l >= d ? l - d : l); }
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) { be_full(m); const unsigned l = be_lane(); return be_exchange(v, l + d < 32 ? l + d : l); }
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { be_full(m); return be_exchange(v, be_lane() ^ (unsigned)x); }
static inline unsigned __ballot_sync(unsigned m, bool p) {
    be_full(m); be_warp& w = be_w();
    w.slot[be_lane()] = p ? 1 : 0; pthread_barrier_wait(&w.bar);
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= (unsigned)w.slot[i] << i;
    pthread_barrier_wait(&w.bar); return r;
}
static inline void __syncwarp(unsigned m = 0xffffffffu) { be_full(m); pthread_barrier_wait(&be_w().bar); }
static inline void __syncthreads() { pthread_barrier_wait(&g_be_block); }
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

// run fn() on a block of n threads (n a multiple of 32, <= 1024); threadIdx.x is set per thread
template <typename F> static void be_run_block(unsigned n, F fn) {
    blockDim.x = n; blockIdx.x = 0; gridDim.x = 1;
    pthread_barrier_init(&g_be_block, nullptr, n);
    for (unsigned w = 0; w < n / 32; w++) pthread_barrier_init(&g_be_warps[w].bar, nullptr, 32);
    struct Arg { F* f; unsigned tid; };
    std::vector<Arg> args(n); std::vector<pthread_t> th(n);
    pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, 1 << 20);
    for (unsigned t = 0; t < n; t++) {
        args[t].f = &fn; args[t].tid = t;
        pthread_create(&th[t], &at, [](void* p) -> void* { Arg* a = (Arg*)p; threadIdx.x = a->tid; (*a->f)(); return nullptr; }, &args[t]);
    }
    for (unsigned t = 0; t < n; t++) pthread_join(th[t], nullptr);
    pthread_attr_destroy(&at);
    for (unsigned w = 0; w < n / 32; w++) pthread_barrier_destroy(&g_be_warps[w].bar);
    pthread_barrier_destroy(&g_be_block);
}
