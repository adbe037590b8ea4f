// Test harness (not product): brpc_b200/csrc/b2_inflate.cuh — the DEVICE inflate — compiled for the host, so that the CPU suite can run it over
// the same thousands of valid / corrupted / truncated / hand-built streams that pin the oracle against the system zlib (tests/test_oracle_gzip.py).
// The GPU tests compare device and oracle on whole messages; this compares the device source itself with zlib, stream by stream.
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define __noinline__
static inline uint32_t __byte_perm(uint32_t x, uint32_t, uint32_t selector) { return selector == 0x0123 ? __builtin_bswap32(x) : x; }
#include "../../include/b2rpc.h"
#include "../../brpc_b200/csrc/b2_inflate.cuh"

extern "C" {
// the bytes the real pass hands over (written to out[0, cap)); *too_big as on the device
uint32_t dev_gz_input_stream(const uint8_t* in, uint32_t n, int format, uint8_t* out, uint32_t cap, int* too_big) {
    bool big = false;
    const uint32_t r = b2::gz_input_stream<true>(in, n, format, out, cap, &big);
    *too_big = big ? 1 : 0;
    return r;
}
// the sizing pass of the decode stage (no output, no data checks): an upper bound, or too_big beyond kGzMaxOut
uint32_t dev_gz_sizing_bound(const uint8_t* in, uint32_t n, int format, int* too_big) {
    bool big = false;
    const uint32_t r = b2::gz_input_stream<false>(in, n, format, nullptr, b2::kGzMaxOut, &big);
    *too_big = big ? 1 : 0;
    return r;
}
uint32_t dev_gz_max_out(void) { return b2::kGzMaxOut; }
}
