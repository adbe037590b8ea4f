// b2_h2.cuh — h2 frame-head scan and HPACK decode on the device.
//   h2 frame heads  <- H2Context::ConsumeFrameHead   src/brpc/policy/http2_rpc_protocol.cpp:438-465
//   HPACK decode    <- HPacker::Decode               src/brpc/details/hpack.cpp:765-843 (+ :531-635, :403-473, :72-229)
// Both are per-connection serial state machines (frame chain; dynamic table), so the unit of
// parallelism is the connection: one thread per connection, thousands of connections per batch.
// The h2 stream state machine, SETTINGS/WINDOW_UPDATE/GOAWAY side effects and response framing
// stay on the host (SURVEY §8a "parity-critical details").
#pragma once
#include <cuda_runtime.h>
#include "b2_core.cuh"
#include "b2_hpack_tables.cuh"

namespace b2 {

struct HpackState {                     // one connection's decoder table (IndexTable, hpack.cpp:72-229)
    uint32_t max_size, size, count, head, byte_head, pad[3];
    struct { uint16_t off, nl, vl, pad; } meta[128];   // ring, newest at `head`
    uint8_t bytes[4096];                // ring of name||value bytes, FIFO like the entries
};

#if defined(__CUDACC__)
__device__ __forceinline__ void hp_pop(HpackState& h) {
    const uint32_t i = (h.head + h.count - 1) & 127u;
    h.size -= (uint32_t)h.meta[i].nl + h.meta[i].vl + 32u;
    h.count--;
}
// AddHeader (hpack.cpp:150-177); entry bytes are read from `src` (name then value, contiguous)
__device__ __forceinline__ int hp_add(HpackState& h, const uint8_t* src, uint32_t nl, uint32_t vl) {
    const uint32_t es = nl + vl + 32u;
    if (nl == 0) return -1;                             // reference CHECK-aborts on an empty name
    while (h.count && h.size + es > h.max_size) hp_pop(h);
    if (es > h.max_size) return 0;
    if (h.count >= 128) return -1;
    h.head = (h.head + 127u) & 127u;
    h.meta[h.head].off = (uint16_t)h.byte_head; h.meta[h.head].nl = (uint16_t)nl; h.meta[h.head].vl = (uint16_t)vl;
    for (uint32_t i = 0; i < nl + vl; i++) h.bytes[(h.byte_head + i) & 4095u] = src[i];
    h.byte_head = (h.byte_head + nl + vl) & 4095u;
    h.count++; h.size += es;
    return 0;
}
// DecodeInteger (hpack.cpp:531-565): >0 bytes used, 0 not enough data, -1 malformed
__device__ __forceinline__ int hp_int(const uint8_t* p, uint32_t n, uint32_t prefix, uint32_t& value) {
    if (n == 0) return 0;
    unsigned long long tmp = p[0] & ((1u << prefix) - 1);
    if (tmp < ((1u << prefix) - 1)) { value = (uint32_t)tmp; return 1; }
    uint32_t i = 1; int m = 0; uint8_t cur;
    do {
        if (i >= n) return 0;
        cur = p[i++];
        tmp += (unsigned long long)(cur & 0x7f) << m;
        m += 7;
    } while ((cur & 0x80) && tmp < 10ull * 1024 * 1024);
    if (tmp >= 10ull * 1024 * 1024) return -1;
    value = (uint32_t)tmp;
    return (int)i;
}
// DecodeString (:606-635) with the Huffman walk of HuffmanDecoder (:414-468) over the pre-built tree
__device__ __forceinline__ int hp_str(const uint8_t* p, uint32_t n, uint8_t* out, uint32_t cap, uint32_t& olen) {
    if (n == 0) return 0;
    const bool huffman = p[0] & 0x80;
    uint32_t length = 0;
    const int ib = hp_int(p, n, 7, length);
    if (ib <= 0) return -1;
    if (length > n - (uint32_t)ib) return 0;
    const uint8_t* s = p + ib;
    if (!huffman) {
        if (length > cap) return -2;
        for (uint32_t i = 0; i < length; i++) out[i] = s[i];
        olen = length; return ib + (int)length;
    }
    int node = 0; uint32_t depth = 0, o = 0; bool padding = true;
    for (uint32_t i = 0; i < length; i++) {
        const uint32_t byte = s[i];
        for (int b = 7; b >= 0; b--) {
            const uint32_t bit = (byte >> b) & 1u;
            const int nx = kHuffTree[node][bit];
            if (nx == 0) return -1;                          // NULL_NODE
            if (nx < 0) {
                const int sym = -nx - 1;
                if (sym == 256) return -1;                   // EOS inside the string
                if (o >= cap) return -2;
                out[o++] = (uint8_t)sym; node = 0; depth = 0; padding = true;
                continue;
            }
            node = nx; depth++; padding = padding && bit;
        }
    }
    if (!(depth == 0 || (depth <= 7 && padding))) return -1;
    olen = o; return ib + (int)length;
}
// HeaderAt: 1..61 static, 62.. dynamic newest first; copies name (and value) into out
__device__ __forceinline__ bool hp_copy_indexed(const HpackState& h, uint32_t index, bool with_value, uint8_t* out, uint32_t cap,
                                                uint32_t& nl, uint32_t& vl, bool& overflow) {
    overflow = false;
    if (index >= 1 && index <= 61) {
        nl = kHpackStaticName[index - 1][1]; vl = with_value ? kHpackStaticValue[index - 1][1] : 0;
        if (nl + vl > cap) { overflow = true; return false; }
        for (uint32_t i = 0; i < nl; i++) out[i] = kHpackStaticBlob[kHpackStaticName[index - 1][0] + i];
        for (uint32_t i = 0; i < vl; i++) out[nl + i] = kHpackStaticBlob[kHpackStaticValue[index - 1][0] + i];
        return true;
    }
    if (index >= 62 && index - 62 < h.count) {
        const uint32_t e = (h.head + (index - 62)) & 127u;
        nl = h.meta[e].nl; vl = with_value ? h.meta[e].vl : 0;
        if (nl + vl > cap) { overflow = true; return false; }
        for (uint32_t i = 0; i < nl + vl; i++) out[i] = h.bytes[(h.meta[e].off + i) & 4095u];
        return true;
    }
    return false;
}
// One header block, the way ConsumeHeaders loops HPacker::Decode.  Records: u16 name_len, u16 value_len, name, value.
// status: 0 consumed, 1 ran out of bytes inside a field, -1 malformed, -2 output capacity exceeded
__device__ __noinline__ int hpack_decode_block(HpackState& h, const uint8_t* in, uint32_t n, uint8_t* out, uint32_t out_cap,
                                               uint32_t& out_len, uint32_t& n_headers) {
    uint32_t pos = 0, o = 0, cnt = 0;
    int status = 0;
    while (pos < n) {
        if (o + 4 > out_cap) { status = -2; break; }
        uint8_t* rec = out + o + 4; const uint32_t cap = out_cap - o - 4;
        uint32_t nl = 0, vl = 0, index = 0;
        const uint8_t* p = in + pos; uint32_t left = n - pos;
        // (001x) dynamic table size updates precede the field they travel with
        int rc = 1; bool size_update_only = false;
        while (left && (p[0] >> 5) == 1) {
            uint32_t max_size = 0;
            const int ib = hp_int(p, left, 5, max_size);
            if (ib <= 0) { rc = ib; break; }
            if (max_size > 4096) { rc = -1; break; }
            if (max_size > h.max_size) h.max_size = max_size;
            else if (max_size < h.max_size) { h.max_size = max_size; while (h.size > h.max_size) hp_pop(h); }
            p += ib; left -= (uint32_t)ib;
            if (!left) { rc = 0; size_update_only = true; }
        }
        (void)size_update_only;
        if (rc <= 0) { status = rc < 0 ? -1 : 1; break; }
        const uint8_t fb = p[0];
        bool ovf = false;
        if (fb & 0x80) {                                     // indexed field
            const int ib = hp_int(p, left, 7, index);
            if (ib <= 0) { status = ib < 0 ? -1 : 1; break; }
            if (!hp_copy_indexed(h, index, true, rec, cap, nl, vl, ovf)) { status = ovf ? -2 : -1; break; }
            p += ib;
        } else {
            const bool incremental = (fb >> 6) == 1;
            const int ib = hp_int(p, left, incremental ? 6 : 4, index);
            if (ib <= 0) { status = -1; break; }
            uint32_t used = (uint32_t)ib;
            if (index != 0) {
                if (!hp_copy_indexed(h, index, false, rec, cap, nl, vl, ovf)) { status = ovf ? -2 : -1; break; }
            } else {
                const int nb = hp_str(p + used, left - used, rec, cap, nl);
                if (nb <= 0) { status = nb == -2 ? -2 : -1; break; }
                used += (uint32_t)nb;
                for (uint32_t i = 0; i < nl; i++) if (rec[i] >= 'A' && rec[i] <= 'Z') rec[i] = (uint8_t)(rec[i] + 32);
            }
            const int vb = hp_str(p + used, left - used, rec + nl, cap - nl, vl);
            if (vb <= 0) { status = vb == -2 ? -2 : -1; break; }
            used += (uint32_t)vb;
            if (incremental && hp_add(h, rec, nl, vl) != 0) { status = -1; break; }
            p += used;
        }
        out[o] = (uint8_t)nl; out[o + 1] = (uint8_t)(nl >> 8); out[o + 2] = (uint8_t)vl; out[o + 3] = (uint8_t)(vl >> 8);
        o += 4 + nl + vl; cnt++;
        pos = (uint32_t)(p - in);
    }
    out_len = o; n_headers = cnt;
    return status;
}

struct H2Frame { uint8_t type, flags; uint16_t pad; uint32_t stream_id, payload_off, payload_len; };   // == b2_h2_frame

// blocks [first[g], first[g+1]) belong to one connection and are decoded in order by one thread
__global__ void k_hpack_decode(const uint8_t* bytes, const uint32_t* blk_conn, const uint32_t* blk_off, const uint32_t* blk_len,
                               const uint32_t* group_first, uint32_t n_groups, HpackState* states, uint8_t* out, uint32_t per_block_cap,
                               uint32_t* out_lens, int32_t* status, uint32_t* n_headers) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    for (uint32_t b = group_first[g]; b < group_first[g + 1]; b++) {
        uint32_t ol = 0, nh = 0;
        const int st = hpack_decode_block(states[blk_conn[b]], bytes + blk_off[b], blk_len[b], out + (size_t)b * per_block_cap, per_block_cap, ol, nh);
        out_lens[b] = ol; status[b] = st; n_headers[b] = nh;
    }
}
__global__ void k_hpack_reset(HpackState* states, uint32_t conn, uint32_t max_size) {
    HpackState& h = states[conn];
    h.max_size = max_size; h.size = 0; h.count = 0; h.head = 0; h.byte_head = 0;
}

// one thread per connection run: the chain of 9-byte frame heads
__global__ void k_h2_scan(const uint8_t* bytes, const b2_run* runs, uint32_t n_runs, uint32_t max_frame_size, H2Frame* frames,
                          uint32_t cap_per_run, uint32_t* n_frames, uint32_t* consumed, uint32_t* err) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint8_t* in = bytes + runs[r].offset; const uint32_t n = runs[r].length;
    uint32_t pos = 0, cnt = 0, e = B2_PARSE_ERROR_NOT_ENOUGH_DATA;
    if (runs[r].flags & 2u) {                               // server side, connection start: the 24-byte client preface
        const char* pre = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";
        const uint32_t k = n < 24 ? n : 24;
        bool match = true;
        for (uint32_t i = 0; i < k; i++) if (in[i] != (uint8_t)pre[i]) { match = false; break; }
        if (!match) { n_frames[r] = 0; consumed[r] = 0; err[r] = B2_PARSE_ERROR_TRY_OTHERS; return; }
        if (n < 24) { n_frames[r] = 0; consumed[r] = 0; err[r] = B2_PARSE_ERROR_NOT_ENOUGH_DATA; return; }
        pos = 24;
    }
    H2Frame* out = frames + (size_t)r * cap_per_run;
    for (;;) {
        if (n - pos < 3) break;
        const uint32_t length = ((uint32_t)in[pos] << 16) | ((uint32_t)in[pos + 1] << 8) | in[pos + 2];
        if (length > max_frame_size) { e = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        if ((unsigned long long)(n - pos - 3) < 6ull + length) break;
        const uint32_t sid = load_be32(in + pos + 5);
        if (sid & 0x80000000u) { e = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        if (cnt < cap_per_run) {
            H2Frame f; f.type = in[pos + 3]; f.flags = in[pos + 4]; f.pad = 0; f.stream_id = sid;
            f.payload_off = runs[r].offset + pos + 9; f.payload_len = length;
            out[cnt] = f;
        }
        cnt++; pos += 9 + length;
    }
    n_frames[r] = cnt; consumed[r] = pos; err[r] = e;
}
#endif
}  // namespace b2
