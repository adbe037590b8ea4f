// Test harness (not product): the host build of brpc_b200/csrc/b2_core.cuh (the same source the kernels compile) — the CutInputMessage chain of
// one run (what k_resolve's sequential semantics are defined by) exported for ctypes, so that the CPU suite can compare the PRODUCT's cut rules
// with the oracle's on every protocol, preferred index, client / rpc_dump run and corruption, without a GPU.
#include <stddef.h>
#include <stdint.h>
#include <string.h>
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

// the product's protobuf wire decoders (host build), flattened for ctypes: returns ok; out = {has, cid, log_id, compress, att, cks_type, content, svc off/len,
// mth off/len, cks off/len, error_code}
extern "C" int core_decode_rpc_meta(const uint8_t* p, uint32_t n, int fast, long long* out) {
    b2::RpcMetaOut o; memset(&o, 0, sizeof o);
    const bool ok = fast ? b2::decode_rpc_meta_fast(p, n, o) : b2::decode_rpc_meta(p, n, o);
    out[0] = o.has; out[1] = o.correlation_id; out[2] = o.log_id; out[3] = o.compress_type; out[4] = o.attachment_size; out[5] = o.checksum_type;
    out[6] = o.content_type; out[7] = o.service_name.off; out[8] = o.service_name.len; out[9] = o.method_name.off; out[10] = o.method_name.len;
    out[11] = o.checksum_value.off; out[12] = o.checksum_value.len; out[13] = o.error_code;
    return ok ? 1 : 0;
}
extern "C" int core_decode_stream_meta(const uint8_t* p, uint32_t n, long long* out) {
    b2::StreamMetaOut o; memset(&o, 0, sizeof o);
    const bool ok = b2::decode_stream_meta(p, n, o);
    out[0] = o.has; out[1] = o.stream_id; out[2] = o.source_stream_id; out[3] = o.consumed_size; out[4] = o.frame_type;
    return ok ? 1 : 0;
}
extern "C" int core_decode_echo_request(const uint8_t* p, uint32_t n, uint32_t* off, uint32_t* len) {
    b2::Span s; s.off = 0; s.len = 0;
    const bool ok = b2::decode_echo_request(p, n, s);
    *off = s.off; *len = s.len;
    return ok ? 1 : 0;
}
