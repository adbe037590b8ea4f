// Test harness (not product): enough of the CUDA execution model and runtime on host threads to build the WHOLE product library
// (brpc_b200/csrc/b2_api.cu with its kernels) with g++ and run it on the CPU — tests/cpp/gen_emul_lib.py writes the translation unit,
// tests/test_emulated_library.py runs the GPU test files against the result.
//
//   kernel<<<grid, block, smem, stream>>>(args)  ->  be_launch(grid, block, smem, stream, [&] { kernel(args); })
//     blocks run one after the other; the threads of a block are host threads; __syncthreads() is a barrier among them;
//     warp collectives (full masks) are a barrier among the 32 lane threads plus an exchange; dynamic shared memory is a heap block of
//     exactly the launch's size (so an overrun is an ASan report); static __shared__ variables are function-local statics
//   cudaMalloc / cudaHostAlloc = malloc (filled with B2_EMUL_FILL, default 0xa5: device memory comes as it is), copies are memcpy,
//     streams are synchronous — except a launch on a stream created for a persistent kernel (be_launch_async), which runs on its own thread
//   cp.async / cp.async.bulk (TMA) copies happen at issue time — or, with B2_EMUL_ASYNC=late, at the latest moment the program's own waits
//     allow (see "asynchronous copies" below): a missing wait shows as wrong bytes.  fence.proxy.async is not modelled
//   A thread that returns from the kernel leaves the barriers (what the hardware does); B2_EMUL_WARN=1 reports a full-mask collective
//     executed after lanes of the warp returned, B2_EMUL_TRACE=1 prints every launch
#pragma once
#include <execinfo.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <set>
#include <thread>
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

// ------------------------------------------------------------------------------------------ the execution model
struct be_dim { unsigned x = 0, y = 0, z = 0; };
// a barrier whose participants can leave: a thread that returns from the kernel no longer counts (what the hardware does with exited threads)
struct be_barrier {
    std::mutex m; std::condition_variable cv; unsigned expected = 0, full = 0, waiting = 0; unsigned long gen = 0;
    void init(unsigned n) { expected = full = n; waiting = 0; gen = 0; }
    void wait() {
        std::unique_lock<std::mutex> l(m);
        const unsigned long g = gen;
        if (++waiting >= expected) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
    void drop() {
        std::unique_lock<std::mutex> l(m);
        expected--;
        if (expected && waiting >= expected) { waiting = 0; gen++; cv.notify_all(); }
    }
};
struct be_warp {
    be_barrier bar; uint64_t slot[32];
    std::mutex gm; std::vector<std::pair<unsigned, be_barrier*>> groups;      // barriers of the lane groups that collectives with a partial mask name
    be_barrier& group(unsigned mask) {
        std::lock_guard<std::mutex> l(gm);
        for (auto& g : groups) if (g.first == mask) return *g.second;
        be_barrier* b = new be_barrier; b->init((unsigned)__builtin_popcount(mask)); groups.push_back({ mask, b }); return *b;
    }
    ~be_warp() { for (auto& g : groups) delete g.second; }
};
struct be_block_state { be_barrier block; be_warp warps[32]; uint8_t* dyn = nullptr; const char* kernel = ""; std::atomic<int> warned{0}; };
static thread_local be_block_state* be_cur = nullptr;
static thread_local be_dim threadIdx, blockIdx, blockDim, gridDim;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 v = { a, b, c, d }; return v; }
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 v = { a, b }; return v; }

static inline be_warp& be_w() { return be_cur->warps[threadIdx.x >> 5]; }
static inline unsigned be_lane() { return threadIdx.x & 31u; }
// the barrier a collective with this mask waits on: the whole warp's (which exited lanes leave) or the named group's
static inline be_barrier& be_bar(unsigned mask) {
    be_warp& w = be_cur->warps[threadIdx.x >> 5];
    if (!(mask >> (threadIdx.x & 31u) & 1u)) { fprintf(stderr, "cuda_emul: %s: lane %u calls a collective whose mask %08x does not name it\n", be_cur->kernel, threadIdx.x & 31u, mask); abort(); }
    if (mask != 0xffffffffu) return w.group(mask);
    if (w.bar.expected != w.bar.full && getenv("B2_EMUL_WARN") && !be_cur->warned.exchange(1)) {
        fprintf(stderr, "cuda_emul: %s: warp collective with a full mask after lanes of the warp returned (thread %u, %u of %u lanes left)\n", be_cur->kernel, threadIdx.x, w.bar.expected, w.bar.full);
        void* bt[24]; const int nb = backtrace(bt, 24); backtrace_symbols_fd(bt, nb, 2);
    }
    return w.bar;
}
template <typename T> static inline T be_exchange(unsigned mask, T v, unsigned src) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 8 bytes");
    be_warp& w = be_w(); be_barrier& bar = be_bar(mask);
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    w.slot[be_lane()] = bits; bar.wait();
    const uint64_t got = w.slot[src & 31u]; bar.wait();
    T r; memcpy(&r, &got, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) { return be_exchange(m, v, (unsigned)src); }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { const unsigned l = be_lane(); return be_exchange(m, v, l >= d ? l - d : l); }
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) { const unsigned l = be_lane(); return be_exchange(m, v, l + d < 32 ? l + d : l); }
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { return be_exchange(m, v, be_lane() ^ (unsigned)x); }
static inline unsigned __ballot_sync(unsigned m, bool p) {
    be_warp& w = be_w(); be_barrier& bar = be_bar(m);
    w.slot[be_lane()] = p ? 1 : 0; bar.wait();
    const unsigned n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
    unsigned r = 0; for (unsigned i = 0; i < n; i++) if (m >> i & 1u) r |= (unsigned)w.slot[i] << i;
    bar.wait(); return r;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0; }
static inline void __syncwarp(unsigned m = 0xffffffffu) { be_bar(m).wait(); }
static inline void __syncthreads() { be_cur->block.wait(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
    const uint64_t v = ((uint64_t)y << 32) | x; uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline unsigned long long be_now_ns() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec; }
static inline unsigned long long clock64() { return be_now_ns(); }
static inline void __nanosleep(unsigned ns) { timespec ts = { 0, (long)ns }; nanosleep(&ts, nullptr); }
using std::min; using std::max;

static inline void be_thread_exit();
static void be_run_block(be_block_state& st, unsigned n, unsigned bidx, unsigned grid, void (*fn)(void*), void* arg) {
    st.block.init(n);
    const unsigned nw = (n + 31) / 32;
    for (unsigned w = 0; w < nw; w++) st.warps[w].bar.init(std::min(32u, n - 32 * w));
    struct Arg { be_block_state* st; unsigned tid, n, bidx, grid; void (*fn)(void*); void* arg; };
    std::vector<Arg> args(n); std::vector<pthread_t> th(n);
    auto body = [](void* p) -> void* {
        Arg* a = (Arg*)p; be_cur = a->st; threadIdx.x = a->tid; blockDim.x = a->n; blockIdx.x = a->bidx; gridDim.x = a->grid;
        a->fn(a->arg);
        be_thread_exit();
        a->st->warps[a->tid >> 5].bar.drop(); a->st->block.drop();      // an exited thread no longer takes part in barriers
        return nullptr;
    };
    for (unsigned t = 0; t < n; t++) { Arg v = { &st, t, n, bidx, grid, fn, arg }; args[t] = v; }
    if (n == 1) body(&args[0]);
    else {
        pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, 2 << 20);
        for (unsigned t = 0; t < n; t++) if (pthread_create(&th[t], &at, body, &args[t]) != 0) { fprintf(stderr, "cuda_emul: pthread_create failed\n"); abort(); }
        for (unsigned t = 0; t < n; t++) pthread_join(th[t], nullptr);
        pthread_attr_destroy(&at);
    }
}

// ------------------------------------------------------------------------------------------ asynchronous copies
// B2_EMUL_ASYNC=issue (default): cp.async / cp.async.bulk copies happen when they are issued.
// B2_EMUL_ASYNC=late: they happen at the LATEST moment the program allows — a bulk load lands when somebody waits on its mbarrier (the
// destination holds 0xEE until then), a bulk store reads its shared-memory source when the issuing thread's wait_group(.read) retires its
// group (or the thread exits), a 16-byte cp.async lands at cp.async.wait_group.  A kernel that reads a staging buffer before the wait, or
// refills it before the store that reads it has drained, produces wrong bytes in this mode.
#include <map>
struct be_copy { void* dst; const void* src; uint32_t n; };
static inline bool be_late() { static const bool v = [] { const char* e = getenv("B2_EMUL_ASYNC"); return e && !strcmp(e, "late"); }(); return v; }
static std::mutex g_be_mbar_mu;
static std::map<const void*, std::vector<be_copy>> g_be_mbar;                 // loads in flight, by the mbarrier that will announce them
static thread_local std::vector<std::vector<be_copy>> be_store_groups;        // this thread's committed bulk-store groups, oldest first
static thread_local std::vector<be_copy> be_store_open, be_cp16_pending;
static inline void be_align16(const void* a, const void* b, uint32_t n, const char* what) {
    if (((uintptr_t)a | (uintptr_t)b | n) & 15u) { fprintf(stderr, "cuda_emul: %s: misaligned %s (%p, %p, %u)\n", be_cur ? be_cur->kernel : "?", what, a, b, n); abort(); }
}
static inline void be_bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, const void* bar) {
    be_align16(sdst, gsrc, bytes, "bulk load");
    if (!be_late()) { memcpy(sdst, gsrc, bytes); return; }
    memset(sdst, 0xEE, bytes);
    std::lock_guard<std::mutex> l(g_be_mbar_mu); g_be_mbar[bar].push_back({ sdst, gsrc, bytes });
}
static inline void be_mbar_wait(const void* bar) {
    if (!be_late()) return;
    std::lock_guard<std::mutex> l(g_be_mbar_mu);
    auto it = g_be_mbar.find(bar);
    if (it == g_be_mbar.end()) return;
    for (const be_copy& c : it->second) memcpy(c.dst, c.src, c.n);
    g_be_mbar.erase(it);
}
static inline void be_bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    be_align16(gdst, ssrc, bytes, "bulk store");
    if (!be_late()) { memcpy(gdst, ssrc, bytes); return; }
    be_store_open.push_back({ gdst, ssrc, bytes });
}
static inline void be_bulk_commit() { if (be_late()) { be_store_groups.push_back(std::move(be_store_open)); be_store_open.clear(); } }
static inline void be_bulk_retire(size_t keep) {
    while (be_store_groups.size() > keep) {
        for (const be_copy& c : be_store_groups.front()) memcpy(c.dst, c.src, c.n);
        be_store_groups.erase(be_store_groups.begin());
    }
}
static inline void be_cp16(void* sdst, const void* gsrc) {
    if (!be_late()) { memcpy(sdst, gsrc, 16); return; }
    memset(sdst, 0xEE, 16); be_cp16_pending.push_back({ sdst, gsrc, 16 });
}
static inline void be_cp16_wait() { for (const be_copy& c : be_cp16_pending) memcpy(c.dst, c.src, c.n); be_cp16_pending.clear(); }
static inline void be_thread_exit() {          // a thread that leaves the kernel has nothing in flight any more
    be_cp16_wait();
    if (!be_store_open.empty()) be_bulk_commit();
    be_bulk_retire(0);
}
static inline void be_prefetch(const void* gsrc, uint32_t bytes) {       // an L2 prefetch reads: a range beyond the buffer is an ASan report
    be_align16(gsrc, gsrc, bytes, "prefetch");
    const volatile uint8_t* p = (const volatile uint8_t*)gsrc; uint8_t a = 0; for (uint32_t i = 0; i < bytes; i++) a ^= p[i]; (void)a;
}

// ------------------------------------------------------------------------------------------ the runtime
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaStreamNonBlocking = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct be_stream { std::thread* worker = nullptr; std::atomic<int> running{0}; };
typedef be_stream* cudaStream_t;
struct be_event { unsigned long long t = 0; };
typedef be_event* cudaEvent_t;

static std::mutex g_be_mu;
static std::set<std::pair<uintptr_t, size_t>> g_be_host;            // cudaHostAlloc'ed ranges (what cudaPointerGetAttributes reports as pinned + mapped)
static inline int be_fill() { static const int v = [] { const char* e = getenv("B2_EMUL_FILL"); return e ? (int)strtol(e, nullptr, 0) & 0xff : 0xa5; }(); return v; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorNotReady ? "not ready" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { const char* e = getenv("B2_EMUL_SMS"); *v = e ? atoi(e) : 2; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetPCIBusId(char* out, int cap, int) { snprintf(out, (size_t)cap, "0000:00:00.0"); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); if (!*p) return cudaErrorMemoryAllocation; memset(*p, be_fill(), n); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) {
    void* q = nullptr; if (posix_memalign(&q, 4096, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    memset(q, be_fill(), n); *p = (T*)q;
    std::lock_guard<std::mutex> g(g_be_mu); g_be_host.insert({ (uintptr_t)q, n ? n : 1 }); return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p) {
    if (!p) return cudaSuccess;
    { std::lock_guard<std::mutex> g(g_be_mu); auto it = g_be_host.lower_bound({ (uintptr_t)p, 0 }); if (it != g_be_host.end() && it->first == (uintptr_t)p) g_be_host.erase(it); }
    free(p); return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { *d = (T*)h; return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    memset(a, 0, sizeof *a);
    std::lock_guard<std::mutex> g(g_be_mu);
    auto it = g_be_host.upper_bound({ (uintptr_t)p, (size_t)-1 });
    if (it != g_be_host.begin()) { --it; if ((uintptr_t)p >= it->first && (uintptr_t)p < it->first + it->second) { a->type = cudaMemoryTypeHost; a->devicePointer = (void*)p; a->hostPointer = (void*)p; } }
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; r++) memmove((uint8_t*)d + r * dp, (const uint8_t*)s + r * sp, w);
    return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* s, size_t n) { memcpy(&sym, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline void be_join(cudaStream_t s) { if (s && s->worker) { s->worker->join(); delete s->worker; s->worker = nullptr; } }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new be_stream; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { be_join(s); delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { be_join(s); return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t s) { return s && s->running.load() ? cudaErrorNotReady : cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new be_event; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = be_now_ns(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((double)(b->t - a->t) * 1e-6); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ------------------------------------------------------------------------------------------ launches
template <typename F> static void be_launch_now(const char* name, unsigned grid, unsigned block, size_t smem, F& fn) {
    static const bool trace = getenv("B2_EMUL_TRACE") != nullptr;
    if (trace) fprintf(stderr, "cuda_emul: %s<<<%u, %u, %zu>>>\n", name, grid, block, smem);
    if (grid == 0 || block == 0 || block > 1024) { fprintf(stderr, "cuda_emul: bad launch configuration %u x %u\n", grid, block); abort(); }
    be_block_state* st = new be_block_state; st->kernel = name;
    void* dyn = nullptr;
    if (smem && posix_memalign(&dyn, 128, smem) != 0) abort();
    st->dyn = (uint8_t*)dyn;
    for (unsigned b = 0; b < grid; b++) {
        if (dyn) memset(dyn, be_fill(), smem);
        be_run_block(*st, block, b, grid, [](void* p) { (*(F*)p)(); }, &fn);
    }
    free(dyn); delete st;
}
template <typename F> static void be_launch(const char* name, unsigned grid, unsigned block, size_t smem, cudaStream_t, F fn) { be_launch_now(name, grid, block, smem, fn); }
// a persistent kernel on its own stream: the launch returns, the kernel runs until it retires (cudaStreamSynchronize / cudaStreamQuery see it)
template <typename F> static void be_launch_async(const char* name, unsigned grid, unsigned block, size_t smem, cudaStream_t s, F fn) {
    be_join(s);
    s->running.store(1);
    s->worker = new std::thread([=]() mutable { be_launch_now(name, grid, block, smem, fn); s->running.store(0); });
}
