// b2_core.cuh — scalar building blocks of the B200 brpc hot path.
//
// Everything here is __host__ __device__ so the same logic that the sm_100a
// kernels run can be unit-tested on the authoring box (tests/emul, CPU,
// diagnostics only — the product never executes these on the host).
//
// Semantics follow the reference (paths relative to apache/brpc):
//   cut_input_message  <- InputMessenger::CutInputMessage  src/brpc/input_messenger.cpp:84-179
//   parse_prefixed     <- ParseRpcMessage                  src/brpc/policy/baidu_rpc_protocol.cpp:105-146
//                         ParseStreamingMessage            src/brpc/policy/streaming_rpc_protocol.cpp:61-96
//   decode_rpc_meta    <- ParsePbFromIOBuf(RpcMeta)        src/brpc/protocol.cpp:202-239 over
//                         policy/baidu_rpc_meta.proto:26-55 (libprotobuf wire semantics)
//   decode_stream_meta <- streaming_rpc_meta.proto:39-53
//   decode_echo_request<- example/echo_c++/echo.proto:23-25
#pragma once
#include <stdint.h>
#include "../../include/b2rpc.h"

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_HD_NOINLINE __host__ __device__ __noinline__
#else
#define B2_HD inline
#define B2_HD_NOINLINE
#endif

namespace b2 {

// little-endian views of the 4-byte magics
constexpr uint32_t kMagicPRPC = 0x43505250u;   // "PRPC"
constexpr uint32_t kMagicSTRM = 0x4d525453u;   // "STRM"
constexpr uint32_t kNone = 0xffffffffu;

B2_HD uint32_t magic_of(int index) { return index == B2_PROTOCOL_BAIDU_STD ? kMagicPRPC : kMagicSTRM; }

B2_HD uint32_t load_le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
B2_HD uint32_t load_be32(const uint8_t* p) {   // butil::RawUnpacker::unpack32, raw_pack.h:76-80
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

// ---------------------------------------------------------------------------
// Protocol::parse for the two 12-byte-header protocols.
struct Cut {
    int err;            // B2_PARSE_*
    uint32_t pop;       // bytes the handler removed from the front of the source
    uint32_t body, meta;
};

B2_HD Cut parse_prefixed(const uint8_t* p, uint32_t n, uint32_t magic, uint64_t max_body) {
    Cut c; c.err = B2_PARSE_OK; c.pop = 0; c.body = 0; c.meta = 0;
    if (n >= 4) {
        if (load_le32(p) != magic) { c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }
    } else {
        for (uint32_t i = 0; i < n; i++)
            if (p[i] != (uint8_t)(magic >> (8 * i))) { c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }
    }
    if (n < 12) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
    const uint32_t body = load_be32(p + 4), meta = load_be32(p + 8);
    if ((uint64_t)body > max_body) { c.err = B2_PARSE_ERROR_TOO_BIG_DATA; return c; }
    if ((uint64_t)n < 12ull + body) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
    if (meta > body) { c.pop = 12 + body; c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }
    c.pop = 12 + body; c.body = body; c.meta = meta;
    return c;
}

// The other length-prefixed protocols that share MostCommonMessage, handler index == ProtocolType
// (src/brpc/options.proto:38-67) so that the probing order of CutInputMessage is the reference's:
//   3  hulu_pbrpc  ParseHuluMessage   policy/hulu_pbrpc_protocol.cpp:178-223  "HULU" body_size meta_size, host (little-endian) order
//   4  sofa_pbrpc  ParseSofaMessage   policy/sofa_pbrpc_protocol.cpp:165-205  "SOFA" meta_size(32) body_size(64) msg_size(64), little-endian
//   12 nshead      ParseNsheadMessage policy/nshead_protocol.cpp:154-182      36-byte nshead_t, magic 0xfb709394 at +24, body_len at +32
constexpr uint32_t kMagicHULU = 0x554c5548u;   // "HULU"
constexpr uint32_t kMagicSOFA = 0x41464f53u;   // "SOFA"
constexpr uint32_t kMagicNshead = 0xfb709394u; // NSHEAD_MAGICNUM, src/brpc/nshead.h:27
constexpr uint32_t kProtoMaskDefault = (1u << 1) | (1u << 2);
constexpr uint32_t kProtoMaskDump = 1u << 31;      // the run is an rpc_dump file, not a socket: records are cut like SampleIterator::Pop
B2_HD uint32_t run_mask(uint32_t ctx_mask, uint32_t run_flags) { return (run_flags & B2_RUN_RPC_DUMP) ? kProtoMaskDump : ctx_mask; }
B2_HD uint32_t frame_header_len(int idx) { return idx == 4 ? 24u : idx == 12 ? 36u : 12u; }
B2_HD uint64_t load_le64(const uint8_t* p) { return (uint64_t)load_le32(p) | ((uint64_t)load_le32(p + 4) << 32); }

// Protocol::parse of handler `idx` on the bytes at p.  Cut.pop = bytes the handler removed (a whole message, or the garbage it
// popped before answering TRY_OTHERS); Cut.body = bytes behind the header that belong to the message, Cut.meta = its meta part.
B2_HD Cut parse_frame(const uint8_t* p, uint32_t n, int idx, uint64_t max_body) {
    if (idx == 1 || idx == 2) return parse_prefixed(p, n, magic_of(idx), max_body);
    Cut c; c.err = B2_PARSE_OK; c.pop = 0; c.body = 0; c.meta = 0;
    if (idx == 3 || idx == 4) {
        const uint32_t magic = idx == 3 ? kMagicHULU : kMagicSOFA;
        if (n >= 4) { if (load_le32(p) != magic) { c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; } }
        else for (uint32_t i = 0; i < n; i++) if (p[i] != (uint8_t)(magic >> (8 * i))) { c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }
        if (idx == 3) {
            if (n < 12) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
            const uint32_t body = load_le32(p + 4), meta = load_le32(p + 8);
            if ((uint64_t)body > max_body) { c.err = B2_PARSE_ERROR_TOO_BIG_DATA; return c; }
            if ((uint64_t)n < 12ull + body) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
            if (meta > body) { c.pop = 12 + body; c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }
            c.pop = 12 + body; c.body = body; c.meta = meta;
            return c;
        }
        if (n < 24) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
        const uint32_t meta = load_le32(p + 4);
        const uint64_t body = load_le64(p + 8), msg = load_le64(p + 16);
        if (msg != (uint64_t)meta + body) { c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }            // (nothing is popped)
        if (body > max_body) { c.err = B2_PARSE_ERROR_TOO_BIG_DATA; return c; }
        if ((uint64_t)n < 24ull + msg) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
        c.pop = 24 + (uint32_t)msg; c.body = (uint32_t)msg; c.meta = meta;
        return c;
    }
    // nshead: the magic sits behind id, version, log_id and provider[16]
    if (n < 28) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
    if (load_le32(p + 24) != kMagicNshead) { c.err = B2_PARSE_ERROR_TRY_OTHERS; return c; }
    if (n < 36) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
    const uint32_t body = load_le32(p + 32);
    if ((uint64_t)body > max_body) { c.err = B2_PARSE_ERROR_TOO_BIG_DATA; return c; }
    if ((uint64_t)n < 36ull + body) { c.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return c; }
    c.pop = 36 + body; c.body = body; c.meta = 0;
    return c;
}
B2_HD int next_handler(uint32_t mask, int after) {             // handlers in index order: 1, 2, 3, 4, 12
    for (int i = after + 1; i <= 12; i++) if ((mask >> i) & 1u) return i;
    return -1;
}

// nshead has no magic in front: it answers NOT_ENOUGH_DATA to ANY 27 bytes and claims whatever carries its magic at +24.  When it is the
// socket's preferred handler it is asked first, so a step at `pos` that another handler would accept comes out differently — the one way
// (besides a handler popping garbage) in which the outcome of CutInputMessage depends on the preferred index.  Speculative walks that
// do not know the preferred index hand such steps to the resolver.
B2_HD bool nshead_claims(const uint8_t* run, uint32_t len, uint32_t pos, uint64_t max_body, uint32_t mask) {
    if (!((mask >> 12) & 1u)) return false;
    return parse_frame(run + pos, len - pos, 12, max_body).err != B2_PARSE_ERROR_TRY_OTHERS;
}

// One CutInputMessage call on a socket whose messenger holds the handlers enabled in `mask`
// (default {1: baidu_std, 2: streaming_rpc}; b2_set_protocols adds 3 hulu_pbrpc, 4 sofa_pbrpc, 12 nshead).
struct Step {
    int err;             // B2_PARSE_OK => one message cut
    int index;           // protocol of the message
    int pf;              // Socket::preferred_index() after the call
    uint32_t frame_pos;  // where the cut message starts (>= pos when a handler popped garbage first)
    uint32_t new_pos;    // read position after the call
    uint32_t body, meta;
    bool popped;         // some handler popped bytes and answered TRY_OTHERS (pf-sensitive path)
};

B2_HD Step cut_input_message(const uint8_t* run, uint32_t len, uint32_t pos, int pf, uint64_t max_body, bool client = false,
                             uint32_t mask = kProtoMaskDefault) {
    Step s; s.err = B2_PARSE_ERROR_TRY_OTHERS; s.index = -1; s.pf = pf; s.frame_pos = pos; s.new_pos = pos;
    s.body = 0; s.meta = 0; s.popped = false;
    if (mask & kProtoMaskDump) {
        // SampleIterator::Pop (src/brpc/rpc_dump.cpp:322-361): records of a dump file carry the baidu_std header
        // (RpcDumpContext::Serialize :237-258); anything malformed is a format error that ends the file — nothing is popped or retried
        const uint32_t n = len - pos; const uint8_t* p = run + pos;
        s.index = B2_PROTOCOL_BAIDU_STD;
        if (n < 12) { s.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return s; }
        if (load_le32(p) != kMagicPRPC) { s.err = B2_PARSE_ERROR_ABSOLUTELY_WRONG; return s; }
        const uint32_t body = load_be32(p + 4), meta = load_be32(p + 8);
        if ((uint64_t)body > max_body) { s.err = B2_PARSE_ERROR_ABSOLUTELY_WRONG; return s; }
        if ((uint64_t)n < 12ull + body) { s.err = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return s; }
        if (meta > body) { s.err = B2_PARSE_ERROR_ABSOLUTELY_WRONG; return s; }
        s.err = B2_PARSE_OK; s.pf = B2_PROTOCOL_BAIDU_STD; s.new_pos = pos + 12 + body; s.body = body; s.meta = meta;
        return s;
    }
    const int pref = pf;
    if (pref >= 1 && pref <= 12 && ((mask >> pref) & 1u)) {
        int cur = pref;
        for (;;) {
            Cut c = parse_frame(run + pos, len - pos, cur, max_body);
            if (c.err == B2_PARSE_OK || c.err == B2_PARSE_ERROR_NOT_ENOUGH_DATA) {
                s.err = c.err; s.index = cur; s.pf = cur;
                if (c.err == B2_PARSE_OK) { s.frame_pos = pos; s.new_pos = pos + c.pop; s.body = c.body; s.meta = c.meta; }
                return s;
            }
            if (c.err != B2_PARSE_ERROR_TRY_OTHERS) { s.err = c.err; return s; }
            if (c.pop) { pos += c.pop; s.popped = true; s.new_pos = pos; }
            if (len - pos >= 4 && load_le32(run + pos) == 0x414d4452u /* "RDMA" */) { s.err = B2_PARSE_ERROR_TRY_OTHERS; return s; }
            if (!client) break;
            // client side (CreatedByConnect): baidu_std may fall to streaming_rpc and back, once;
            // anything else is fixed by the channel's protocol (input_messenger.cpp:122-138)
            if (cur == pref && (cur == 1 || cur == 2)) { cur = 3 - pref; continue; }
            s.err = B2_PARSE_ERROR_ABSOLUTELY_WRONG; return s;
        }
        s.pf = -1;
    }
    for (int i = next_handler(mask, 0); i > 0; i = next_handler(mask, i)) {
        if (i == pref) continue;
        Cut c = parse_frame(run + pos, len - pos, i, max_body);
        if (c.err == B2_PARSE_OK || c.err == B2_PARSE_ERROR_NOT_ENOUGH_DATA) {
            s.err = c.err; s.index = i; s.pf = i;
            if (c.err == B2_PARSE_OK) { s.frame_pos = pos; s.new_pos = pos + c.pop; s.body = c.body; s.meta = c.meta; }
            return s;
        }
        if (c.err != B2_PARSE_ERROR_TRY_OTHERS) { s.err = c.err; return s; }
        if (c.pop) { pos += c.pop; s.popped = true; s.new_pos = pos; }
    }
    s.err = B2_PARSE_ERROR_TRY_OTHERS;
    return s;
}

// ---------------------------------------------------------------------------
// protobuf wire reader (libprotobuf parse_context.h semantics, see oracle notes)
struct Reader { const uint8_t* p; const uint8_t* end; };

B2_HD bool rd_varint(Reader& r, uint64_t& out) {       // VarintParse<uint64_t>: <= 10 bytes
    if (r.p < r.end && *r.p < 0x80) { out = *r.p++; return true; }
    uint64_t v = 0;
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    for (int i = 0; i < 10; i++) {
        if (r.p >= r.end) return false;
        const uint8_t b = *r.p++;
        v |= (uint64_t)(b & 0x7f) << (7 * i);
        if (b < 0x80) { out = v; return true; }
    }
    return false;
}
B2_HD bool rd_tag(Reader& r, uint32_t& tag) {           // ReadTag: <= 5 bytes
    if (r.p < r.end && *r.p < 0x80) { tag = *r.p++; return true; }
    uint32_t v = 0;
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    for (int i = 0; i < 5; i++) {
        if (r.p >= r.end) return false;
        const uint8_t b = *r.p++;
        v |= (uint32_t)(b & 0x7f) << (7 * i);
        if (b < 0x80) { tag = v; return true; }
    }
    return false;
}
B2_HD bool rd_size(Reader& r, uint32_t& n) {            // ReadSize: <= 5 bytes, < 2 GiB - 16, inside parent
    if (r.p < r.end && *r.p < 0x80) {
        const uint32_t v1 = *r.p++;
        if ((uint64_t)v1 > (uint64_t)(r.end - r.p)) return false;
        n = v1; return true;
    }
    uint32_t v = 0;
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    for (int i = 0; i < 5; i++) {
        if (r.p >= r.end) return false;
        const uint8_t b = *r.p++;
        if (i == 4 && b >= 8) return false;
        v |= (uint32_t)(b & 0x7f) << (7 * i);
        if (b < 0x80) {
            if (v > 0x7fffffffu - 16u) return false;
            if ((uint64_t)v > (uint64_t)(r.end - r.p)) return false;
            n = v; return true;
        }
    }
    return false;
}
// skip one non-group field
B2_HD bool skip_scalar(Reader& r, uint32_t wt) {
    uint64_t v; uint32_t n;
    switch (wt) {
    case 0: return rd_varint(r, v);
    case 1: if (r.end - r.p < 8) return false; r.p += 8; return true;
    case 2: if (!rd_size(r, n)) return false; r.p += n; return true;
    case 5: if (r.end - r.p < 4) return false; r.p += 4; return true;
    }
    return false;
}
// unknown group: ParseContext::ParseGroup nesting, depth budget as in the oracle
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
inline bool skip_group(Reader& r, uint32_t fn0, int budget) {
    uint32_t st[100]; int d = 0;
    if (budget <= 0) return false;
    st[d++] = fn0;
    while (d > 0) {
        uint32_t tag;
        if (r.p >= r.end) return false;
        if (!rd_tag(r, tag)) return false;
        if (tag == 0) return false;
        const uint32_t wt = tag & 7, fn = tag >> 3;
        if (wt == 4) { if (fn != st[d - 1]) return false; d--; continue; }
        if (fn == 0) return false;
        if (wt == 3) { if (d >= budget || d >= 100) return false; st[d++] = fn; continue; }
        if (!skip_scalar(r, wt)) return false;
    }
    return true;
}
B2_HD bool skip_field(Reader& r, uint32_t tag, int budget) {
    const uint32_t wt = tag & 7;
    if (wt == 3) return skip_group(r, tag >> 3, budget);
    return skip_scalar(r, wt);
}

struct Span { uint32_t off, len; };

// Decoded RpcMeta: what ProcessRpcRequest / SendRpcResponse consume.
struct RpcMetaOut {
    uint32_t has;               // B2_HAS_*
    int64_t correlation_id, log_id;
    int32_t compress_type, attachment_size, checksum_type, content_type;
    Span service_name, method_name, checksum_value;   // offsets relative to the meta start
    int32_t error_code;         // RpcResponseMeta.error_code (0 when absent)
};

B2_HD bool rd_span(Reader& r, const uint8_t* base, Span& s) {
    uint32_t n;
    if (!rd_size(r, n)) return false;
    s.off = (uint32_t)(r.p - base); s.len = n; r.p += n;
    return true;
}

// Sub-messages of RpcMeta are all flat, so one loop with a two-level cursor
// (top level / one nested message) replaces libprotobuf's recursion.
enum SubKind { kTop = 0, kRequest, kResponse, kChunk, kStreamSettings, kMapEntry };

B2_HD bool decode_rpc_meta(const uint8_t* p, uint32_t n, RpcMetaOut& o) {
    Reader r; r.p = p; r.end = p + n;
    const uint8_t* top_end = r.end;
    int kind = kTop;
    uint32_t req_bits = 0, chunk_bits = 0, ss_bits = 0;
    o.has = 0; o.correlation_id = 0; o.log_id = 0; o.compress_type = 0; o.attachment_size = 0;
    o.checksum_type = 0; o.content_type = 0; o.error_code = 0;
    o.service_name.off = o.service_name.len = 0; o.method_name = o.service_name; o.checksum_value = o.service_name;
    for (;;) {
        if (r.p >= r.end) {
            if (kind == kTop) break;
            kind = kTop; r.end = top_end;       // nested message ended exactly at its limit
            continue;
        }
        uint32_t tag;
        if (!rd_tag(r, tag)) return false;
        if (tag == 0 || (tag & 7) == 4) return false;
        const uint32_t fn = tag >> 3, wt = tag & 7;
        if (fn == 0) return false;
        const int budget = (kind == kTop) ? 100 : 99;
        uint64_t v; Span sp;
        bool handled = false;
        if (kind == kTop) {
            if (wt == 2 && (fn == 1 || fn == 2 || fn == 6 || fn == 8 || fn == 9)) {
                uint32_t len;
                if (!rd_size(r, len)) return false;
                r.end = r.p + len;
                if (fn == 1) { kind = kRequest; o.has |= B2_HAS_REQUEST; }
                else if (fn == 2) { kind = kResponse; o.has |= B2_HAS_RESPONSE; }
                else if (fn == 6) { kind = kChunk; o.has |= B2_HAS_CHUNK_INFO; }
                else if (fn == 8) { kind = kStreamSettings; o.has |= B2_HAS_STREAM_SETTINGS; }
                else { kind = kMapEntry; o.has |= B2_HAS_USER_FIELDS; }
                continue;
            }
            if (wt == 0 && (fn == 3 || fn == 4 || fn == 5 || fn == 10 || fn == 11)) {
                if (!rd_varint(r, v)) return false;
                handled = true;
                if (fn == 3) { o.compress_type = (int32_t)(uint32_t)v; o.has |= B2_HAS_COMPRESS_TYPE; }
                else if (fn == 4) { o.correlation_id = (int64_t)v; o.has |= B2_HAS_CORRELATION_ID; }
                else if (fn == 5) { o.attachment_size = (int32_t)(uint32_t)v; o.has |= B2_HAS_ATTACHMENT_SIZE; }
                else if (fn == 10) {     // closed enum ContentType: unknown numbers leave the field unset
                    const int32_t e = (int32_t)(uint32_t)v;
                    if (e >= 0 && e <= 3) { o.content_type = e; o.has |= B2_HAS_CONTENT_TYPE; }
                } else { o.checksum_type = (int32_t)(uint32_t)v; o.has |= B2_HAS_CHECKSUM_TYPE; }
            } else if (wt == 2 && (fn == 7 || fn == 12)) {
                if (!rd_span(r, p, sp)) return false;
                handled = true;
                if (fn == 7) o.has |= B2_HAS_AUTH_DATA;
                else { o.checksum_value = sp; o.has |= B2_HAS_CHECKSUM_VALUE; }
            }
        } else if (kind == kRequest) {
            if (wt == 2 && (fn == 1 || fn == 2 || fn == 7)) {
                if (!rd_span(r, p, sp)) return false;
                handled = true;
                if (fn == 1) { o.service_name = sp; req_bits |= 1; }
                else if (fn == 2) { o.method_name = sp; req_bits |= 2; }
                else o.has |= B2_HAS_REQUEST_ID;
            } else if (wt == 0 && fn >= 3 && fn <= 8 && fn != 7) {
                if (!rd_varint(r, v)) return false;
                handled = true;
                if (fn == 3) { o.log_id = (int64_t)v; o.has |= B2_HAS_LOG_ID; }
                else if (fn == 4) o.has |= B2_HAS_TRACE_ID;
                else if (fn == 8) o.has |= B2_HAS_TIMEOUT_MS;
            }
        } else if (kind == kResponse) {
            if (wt == 0 && fn == 1) { if (!rd_varint(r, v)) return false; handled = true; o.error_code = (int32_t)(uint32_t)v; }
            else if (wt == 2 && fn == 2) { if (!rd_span(r, p, sp)) return false; handled = true; }
        } else if (kind == kChunk) {
            if (wt == 0 && (fn == 1 || fn == 2)) { if (!rd_varint(r, v)) return false; handled = true; chunk_bits |= fn; }
        } else if (kind == kStreamSettings) {
            if (wt == 0 && fn >= 1 && fn <= 4) { if (!rd_varint(r, v)) return false; handled = true; if (fn == 1) ss_bits = 1; }
            else if (wt == 2 && fn == 4) {      // packed repeated int64
                uint32_t len;
                if (!rd_size(r, len)) return false;
                Reader s; s.p = r.p; s.end = r.p + len; r.p += len;
                while (s.p < s.end) if (!rd_varint(s, v)) return false;
                handled = true;
            }
        } else {   // kMapEntry {1: key, 2: value}
            if (wt == 2 && (fn == 1 || fn == 2)) { if (!rd_span(r, p, sp)) return false; handled = true; }
        }
        if (!handled && !skip_field(r, tag, budget)) return false;
    }
    // IsInitialized(): required fields of the sub-messages that are present
    if ((o.has & B2_HAS_REQUEST) && req_bits != 3) return false;
    if ((o.has & B2_HAS_CHUNK_INFO) && chunk_bits != 3) return false;
    if ((o.has & B2_HAS_STREAM_SETTINGS) && !ss_bits) return false;
    return true;
}

// Fast path for the exact shape PackRpcRequest emits (baidu_rpc_protocol.cpp:1045-1133): known fields in
// ascending order, each at most once, one-byte tags and lengths:
//   0a L { 0a sl <service> 12 ml <method> [18 log_id] }  [18 compress] [20 cid] [28 att] [50 content] [58 cktype] [62 len <bytes>]
// Returns false on ANY deviation — the caller then runs decode_rpc_meta, which alone defines the semantics;
// for inputs it accepts, the result is identical field by field (same varint / int32 truncation rules).
B2_HD bool decode_rpc_meta_fast(const uint8_t* p, uint32_t n, RpcMetaOut& o) {
    o.has = 0; o.correlation_id = 0; o.log_id = 0; o.compress_type = 0; o.attachment_size = 0;
    o.checksum_type = 0; o.content_type = 0; o.error_code = 0;
    o.service_name.off = o.service_name.len = 0; o.method_name = o.service_name; o.checksum_value = o.service_name;
    if (n < 6 || p[0] != 0x0a) return false;
    const uint32_t L = p[1];
    if (L >= 128 || 2 + L > n) return false;
    const uint8_t* q = p + 2; const uint8_t* e = q + L;
    if (e - q < 2 || q[0] != 0x0a || q[1] >= 128 || (uint32_t)(e - q - 2) < q[1]) return false;
    o.service_name.off = (uint32_t)(q + 2 - p); o.service_name.len = q[1]; q += 2 + q[1];
    if (e - q < 2 || q[0] != 0x12 || q[1] >= 128 || (uint32_t)(e - q - 2) < q[1]) return false;
    o.method_name.off = (uint32_t)(q + 2 - p); o.method_name.len = q[1]; q += 2 + q[1];
    o.has = B2_HAS_REQUEST;
    Reader r; uint64_t v;
    if (q < e && *q == 0x18) { r.p = q + 1; r.end = e; if (!rd_varint(r, v)) return false; o.log_id = (int64_t)v; o.has |= B2_HAS_LOG_ID; q = r.p; }
    if (q != e) return false;
    r.p = e; r.end = p + n;
    if (r.p < r.end && *r.p == 0x18) { r.p++; if (!rd_varint(r, v)) return false; o.compress_type = (int32_t)(uint32_t)v; o.has |= B2_HAS_COMPRESS_TYPE; }
    if (r.p < r.end && *r.p == 0x20) { r.p++; if (!rd_varint(r, v)) return false; o.correlation_id = (int64_t)v; o.has |= B2_HAS_CORRELATION_ID; }
    if (r.p < r.end && *r.p == 0x28) { r.p++; if (!rd_varint(r, v)) return false; o.attachment_size = (int32_t)(uint32_t)v; o.has |= B2_HAS_ATTACHMENT_SIZE; }
    if (r.p < r.end && *r.p == 0x50) {
        r.p++; if (!rd_varint(r, v)) return false;
        const int32_t ct = (int32_t)(uint32_t)v;
        if (ct < 0 || ct > 3) return false;                  // closed-enum corner: leave it to the generic decoder
        o.content_type = ct; o.has |= B2_HAS_CONTENT_TYPE;
    }
    if (r.p < r.end && *r.p == 0x58) { r.p++; if (!rd_varint(r, v)) return false; o.checksum_type = (int32_t)(uint32_t)v; o.has |= B2_HAS_CHECKSUM_TYPE; }
    if (r.p < r.end && *r.p == 0x62) {
        if (r.end - r.p < 2 || r.p[1] >= 128 || (uint32_t)(r.end - r.p - 2) < r.p[1]) return false;
        o.checksum_value.off = (uint32_t)(r.p + 2 - p); o.checksum_value.len = r.p[1]; o.has |= B2_HAS_CHECKSUM_VALUE; r.p += 2 + r.p[1];
    }
    return r.p == r.end;
}

struct StreamMetaOut { uint32_t has; int64_t stream_id, source_stream_id, consumed_size; int32_t frame_type; };

B2_HD bool decode_stream_meta(const uint8_t* p, uint32_t n, StreamMetaOut& o) {
    Reader r; r.p = p; r.end = p + n;
    const uint8_t* top_end = r.end;
    bool in_feedback = false;
    o.has = 0; o.stream_id = 0; o.source_stream_id = 0; o.consumed_size = 0; o.frame_type = 0;
    for (;;) {
        if (r.p >= r.end) {
            if (!in_feedback) break;
            in_feedback = false; r.end = top_end; continue;
        }
        uint32_t tag;
        if (!rd_tag(r, tag)) return false;
        if (tag == 0 || (tag & 7) == 4) return false;
        const uint32_t fn = tag >> 3, wt = tag & 7;
        if (fn == 0) return false;
        uint64_t v; bool handled = false;
        if (!in_feedback) {
            if (wt == 0 && fn >= 1 && fn <= 4) {
                if (!rd_varint(r, v)) return false;
                handled = true;
                if (fn == 1) { o.stream_id = (int64_t)v; o.has |= B2_SHAS_STREAM_ID; }
                else if (fn == 2) { o.source_stream_id = (int64_t)v; o.has |= B2_SHAS_SOURCE_STREAM_ID; }
                else if (fn == 3) { const int32_t e = (int32_t)(uint32_t)v; if (e >= 0 && e <= 4) { o.frame_type = e; o.has |= B2_SHAS_FRAME_TYPE; } }
                else { o.has |= B2_SHAS_HAS_CONTINUATION; if (v) o.has |= B2_SVAL_HAS_CONTINUATION; else o.has &= ~B2_SVAL_HAS_CONTINUATION; }
            } else if (wt == 2 && fn == 5) {
                uint32_t len;
                if (!rd_size(r, len)) return false;
                r.end = r.p + len; in_feedback = true; o.has |= B2_SHAS_FEEDBACK;
                continue;
            }
        } else if (wt == 0 && fn == 1) {
            if (!rd_varint(r, v)) return false;
            handled = true; o.consumed_size = (int64_t)v;
        }
        if (!handled && !skip_field(r, tag, in_feedback ? 99 : 100)) return false;
    }
    return (o.has & B2_SHAS_STREAM_ID) != 0;
}

// EchoRequest{required string message = 1}: span of the LAST occurrence, relative to p.
B2_HD bool decode_echo_request(const uint8_t* p, uint32_t n, Span& msg) {
    Reader r; r.p = p; r.end = p + n;
    bool has = false;
    while (r.p < r.end) {
        uint32_t tag;
        if (!rd_tag(r, tag)) return false;
        if (tag == 0 || (tag & 7) == 4) return false;
        if ((tag >> 3) == 0) return false;
        if (tag == ((1u << 3) | 2u)) { if (!rd_span(r, p, msg)) return false; has = true; }
        else if (!skip_field(r, tag, 100)) return false;
    }
    return has;
}

// RpcDumpMeta (src/brpc/rpc_dump.proto:23-48), proto2, every field optional: what rpc_replay needs to re-issue a sampled request.
struct DumpMetaOut { uint32_t has; Span service_name, method_name; int32_t compress_type, protocol_type, attachment_size; };
enum { kDumpHasService = 1, kDumpHasMethod = 2, kDumpHasCompress = 4, kDumpHasProtocol = 8, kDumpHasAttachment = 16 };
B2_HD bool decode_dump_meta(const uint8_t* p, uint32_t n, DumpMetaOut& o) {
    Reader r; r.p = p; r.end = p + n;
    o.has = 0; o.service_name.off = o.service_name.len = 0; o.method_name = o.service_name; o.compress_type = 0; o.protocol_type = 0; o.attachment_size = 0;
    while (r.p < r.end) {
        uint32_t tag;
        if (!rd_tag(r, tag)) return false;
        if (tag == 0 || (tag & 7) == 4 || (tag >> 3) == 0) return false;
        const uint32_t fn = tag >> 3, wt = tag & 7;
        uint64_t v; Span sp; bool handled = false;
        if (wt == 2 && (fn == 1 || fn == 2 || (fn >= 7 && fn <= 9))) {
            if (!rd_span(r, p, sp)) return false;
            handled = true;
            if (fn == 1) { o.service_name = sp; o.has |= kDumpHasService; } else if (fn == 2) { o.method_name = sp; o.has |= kDumpHasMethod; }
        } else if (wt == 0 && fn >= 3 && fn <= 6) {
            if (!rd_varint(r, v)) return false;
            handled = true;
            const int32_t e = (int32_t)(uint32_t)v;
            if (fn == 4) { if (e >= 0 && e <= 4) { o.compress_type = e; o.has |= kDumpHasCompress; } }        // closed enums: unknown numbers stay unset
            else if (fn == 5) { if (e >= 0 && e <= 27) { o.protocol_type = e; o.has |= kDumpHasProtocol; } }
            else if (fn == 6) { o.attachment_size = e; o.has |= kDumpHasAttachment; }
        }
        if (!handled && !skip_field(r, tag, 100)) return false;
    }
    return true;
}
// ---------------------------------------------------------------------------
// encoders
B2_HD uint32_t varint_len(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)(64 - __clzll((long long)(v | 1)) + 6) / 7u;         // ceil(significant bits / 7)
#else
    uint32_t n = 1; while (v >= 0x80) { v >>= 7; n++; } return n;
#endif
}
B2_HD uint8_t* put_varint(uint8_t* p, uint64_t v) { while (v >= 0x80) { *p++ = (uint8_t)(v | 0x80); v >>= 7; } *p++ = (uint8_t)v; return p; }
B2_HD uint8_t* put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; return p + 4; }
B2_HD uint32_t dec_len(uint32_t v) { uint32_t n = 1; while (v >= 10) { v /= 10; n++; } return n; }
B2_HD uint8_t* put_dec(uint8_t* p, uint32_t v) {
    const uint32_t n = dec_len(v);
    for (uint32_t i = 0; i < n; i++) { p[n - 1 - i] = (uint8_t)('0' + v % 10); v /= 10; }
    return p + n;
}
B2_HD uint8_t* put_dec_i32(uint8_t* p, int32_t v) {    // printf("%d")
    if (v < 0) { *p++ = '-'; return put_dec(p, (uint32_t)(-(int64_t)v)); }
    return put_dec(p, (uint32_t)v);
}
B2_HD uint32_t dec_len_i32(int32_t v) { return v < 0 ? 1 + dec_len((uint32_t)(-(int64_t)v)) : dec_len((uint32_t)v); }

// bytes of the RpcMeta PackRpcRequest builds when it REPLAYS a sampled request (baidu_rpc_protocol.cpp:1067-1075,1080,1106-1120): request{service_name,
// method_name}, compress_type, correlation_id, [attachment_size], content_type — no checksum fields on this branch
B2_HD uint32_t replay_meta_len(uint32_t svc_len, uint32_t mth_len, int32_t compress_type, int64_t correlation_id, uint32_t attached) {
    const uint32_t rl = 1 + varint_len(svc_len) + svc_len + 1 + varint_len(mth_len) + mth_len;
    uint32_t n = 1 + varint_len(rl) + rl + 1 + varint_len((uint64_t)(int64_t)compress_type) + 1 + varint_len((uint64_t)correlation_id);
    if (attached) n += 1 + varint_len(attached);
    return n + 2;
}

// CRC-32C Mask, src/butil/crc32c.h:38-47
B2_HD uint32_t crc32c_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }
B2_HD uint32_t crc32c_unmask(uint32_t m) { const uint32_t rot = m - 0xa282ead8u; return (rot >> 17) | (rot << 15); }

// Length of the RpcMeta SendRpcResponse builds (baidu_rpc_protocol.cpp:339-349)
B2_HD uint32_t response_meta_len(int32_t error_code, uint32_t error_text_len, int32_t compress_type,
                                 int64_t correlation_id, uint32_t attached_size, int32_t checksum_type,
                                 uint32_t cks_len) {
    uint32_t rm = 1 + varint_len((uint64_t)(int64_t)error_code);
    if (error_text_len) rm += 1 + varint_len(error_text_len) + error_text_len;
    uint32_t n = 1 + varint_len(rm) + rm;
    n += 1 + varint_len((uint64_t)(int64_t)compress_type);
    n += 1 + varint_len((uint64_t)correlation_id);
    if (attached_size) n += 1 + varint_len(attached_size);
    n += 2;                                                  // content_type = PB: 50 00
    n += 1 + varint_len((uint64_t)(int64_t)checksum_type);
    n += 1 + varint_len(cks_len) + cks_len;
    return n;
}

}  // namespace b2
