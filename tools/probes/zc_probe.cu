// zc_probe — measurement tool (not product): how fast can kernels PULL sparse bytes out of mapped pinned host memory?
// For a stride S and a contiguous read size R per "message", every warp lane group reads R bytes at positions k*S.
// Prints GB/s of useful bytes and requests/s.  Also: latency of a dependent chain of zero-copy loads.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// each group of (R/16) lanes reads one record's first R bytes (16 B per lane), records are S bytes apart
__global__ void k_read(const uint8_t* base, size_t n_rec, uint32_t S, uint32_t R, unsigned long long* sink) {
    const uint32_t lanes_per = R / 16;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t t = gtid; t < n_rec * lanes_per; t += total) {
        const size_t rec = t / lanes_per; const uint32_t sub = (uint32_t)(t % lanes_per);
        const uint4 v = *reinterpret_cast<const uint4*>(base + rec * S + sub * 16);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x1234567) *sink = acc;
}
// pointer chase: each thread follows `hops` dependent loads (offset stored in the data), threads independent
__global__ void k_chase(const uint8_t* base, uint32_t n_chains, uint32_t chain_stride, uint32_t hops, unsigned long long* sink) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_chains) return;
    uint32_t pos = t * chain_stride; unsigned long long acc = 0;
    for (uint32_t h = 0; h < hops; h++) { const uint32_t nxt = *reinterpret_cast<const uint32_t*>(base + pos); acc += nxt; pos = nxt; }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    const size_t N = 256ull << 20;
    uint8_t* h; CK(cudaHostAlloc(&h, N, cudaHostAllocMapped | cudaHostAllocPortable));
    for (size_t i = 0; i < N; i += 4) *(uint32_t*)(h + i) = (uint32_t)(i * 2654435761u);
    uint8_t* d_map; CK(cudaHostGetDevicePointer(&d_map, h, 0));
    uint8_t* d_hbm; CK(cudaMalloc(&d_hbm, N)); CK(cudaMemcpy(d_hbm, h, N, cudaMemcpyHostToDevice));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("mapped ptr %s host ptr\n", d_map == h ? "==" : "!=");
    const uint32_t strides[] = {1088, 1088, 1088, 1088, 8192, 8192, 128, 4096};
    const uint32_t reads[]   = {  32,   64,  128,  256,  512, 1024, 128, 4096};
    for (int src = 0; src < 2; src++) {
        const uint8_t* base = src == 0 ? d_map : d_hbm;
        for (int c = 0; c < 8; c++) {
            const uint32_t S = strides[c], R = reads[c];
            const size_t n_rec = N / S;
            for (int grid_mul = 1; grid_mul <= 8; grid_mul *= 8) {
                k_read<<<148 * grid_mul, 256>>>(base, n_rec, S, R, sink);   // warm
                cudaEventRecord(e0);
                k_read<<<148 * grid_mul, 256>>>(base, n_rec, S, R, sink);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("%s stride %5u read %5u grid %5d: %8.3f ms  %7.2f GB/s useful  %7.1f M records/s\n", src == 0 ? "zero-copy" : "hbm      ", S, R, 148 * grid_mul,
                       ms, (double)n_rec * R / ms / 1e6, (double)n_rec / ms / 1e3);
            }
        }
    }
    // dependent chains: chain c hops through its own 8 KB region, 1088 bytes per hop
    {
        const uint32_t cs = 8192, hops = 7;
        const uint32_t n_chains = (uint32_t)(N / cs);
        for (uint32_t c = 0; c < n_chains; c++) for (uint32_t k = 0; k < hops; k++) *(uint32_t*)(h + (size_t)c * cs + k * 1088) = c * cs + (k + 1) * 1088;
        CK(cudaMemcpy(d_hbm, h, N, cudaMemcpyHostToDevice));
        for (int src = 0; src < 2; src++) {
            const uint8_t* base = src == 0 ? d_map : d_hbm;
            k_chase<<<(n_chains + 127) / 128, 128>>>(base, n_chains, cs, hops, sink);
            cudaEventRecord(e0);
            k_chase<<<(n_chains + 127) / 128, 128>>>(base, n_chains, cs, hops, sink);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%s chase: %u chains x %u hops: %8.3f ms (%.2f us per hop if serial; %.1f M hops/s)\n", src == 0 ? "zero-copy" : "hbm      ", n_chains, hops, ms, ms * 1e3 / hops, (double)n_chains * hops / ms / 1e3);
        }
        // single chain latency
        k_chase<<<1, 1>>>(d_map, 1, cs, hops, sink);
        cudaEventRecord(e0); k_chase<<<1, 1>>>(d_map, 1, cs, hops, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); printf("zero-copy single chain: %.2f us per hop\n", ms * 1e3 / hops);
        cudaEventRecord(e0); k_chase<<<1, 1>>>(d_hbm, 1, cs, hops, sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("hbm single chain: %.2f us per hop (incl. launch)\n", ms * 1e3 / hops);
    }
    // plain copies for comparison
    cudaEventRecord(e0); for (int i = 0; i < 5; i++) CK(cudaMemcpyAsync(d_hbm, h, N, cudaMemcpyHostToDevice)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    { float ms; cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemcpyAsync H2D: %.2f GB/s\n", 5.0 * N / ms / 1e6); }
    return 0;
}
