// Test harness (not product): the host build of brpc_b200/csrc/b2_core.cuh (the same source the kernels compile) — the CutInputMessage chain of
// one run (what k_resolve's sequential semantics are defined by) exported for ctypes, so that the CPU suite can compare the PRODUCT's cut rules
// with the oracle's on every protocol, preferred index, client / rpc_dump run and corruption, without a GPU.
#include <stddef.h>
#include <stdint.h>
#include "../../brpc_b200/csrc/b2_core.cuh"

extern "C" {
// ProcessNewMessage's loop over one run: returns the number of messages; out[0] consumed, [1] the ParseError that ended the loop, [2] the
// preferred index afterwards; offs[i] = frame start | (not baidu_std) << 31 for the first `cap` messages
uint32_t core_cut_run(const uint8_t* run, uint32_t len, int preferred, uint64_t max_body, int client, uint32_t ctx_mask, uint32_t run_flags,
                      uint32_t* out, uint32_t* offs, uint32_t cap) {
    const uint32_t mask = b2::run_mask(ctx_mask, run_flags);
    uint32_t pos = 0, n = 0; int pf = preferred;
    for (;;) {
        const b2::Step s = b2::cut_input_message(run, len, pos, pf, max_body, client != 0, mask);
        pos = s.new_pos; pf = s.pf;
        if (s.err != B2_PARSE_OK) { out[1] = (uint32_t)s.err; break; }
        if (n < cap) offs[n] = s.frame_pos | ((uint32_t)(s.index != 1) << 31);
        n++;
    }
    out[0] = pos; out[2] = (uint32_t)pf;
    return n;
}
}
