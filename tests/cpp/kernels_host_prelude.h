// Test harness (not product): the rest of what a host compiler needs for tests/cpp/kernels_host.cuh (the generated host-compilable copy of
// b2_kernels.cuh).  A "warp" has ONE lane and a block one thread: only per-thread device functions may be called through this.
#pragma once
#include "h2_host_prelude.h"
#define __maxnreg__(...)
#define __restrict__
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 v = { a, b }; return v; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = (T)(o | v); return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline unsigned long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}
