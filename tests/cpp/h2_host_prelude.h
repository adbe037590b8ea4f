// Test harness (not product): what a host compiler needs to build brpc_b200/csrc/b2_h2.cuh — the DEVICE h2 / HPACK code — as plain C++.
// One "thread" at a time: threadIdx / blockIdx are variables the harness sets before calling a kernel function.
#pragma once
#include <stddef.h>
#include <stdint.h>
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
#define __align__(n) __attribute__((aligned(n)))
struct h2h_dim { unsigned x = 0, y = 0, z = 0; };
static h2h_dim threadIdx, blockIdx, blockDim, gridDim;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 v = { a, b, c, d }; return v; }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t, uint32_t sel) { return sel == 0x0123 ? __builtin_bswap32(x) : x; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }        // a "warp" of one
static inline void __syncwarp(unsigned = 0xffffffffu) {}
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::min; using std::max;
