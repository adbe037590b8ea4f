// rpc_press.cpp — synthetic baidu_std traffic source (the load generator, NOT the hot path).
//
// Plays the role of tools/rpc_press (tools/rpc_press/rpc_press_impl.cpp:218-262) and of
// example/multi_threaded_echo_c++/client.cpp:61-116 for the benchmarks: it produces the
// byte stream a brpc client would write on one connection — frames built the way
// PackRpcRequest does (src/brpc/policy/baidu_rpc_protocol.cpp:1045-1133):
//   "PRPC" be32(meta+body) be32(meta) | RpcMeta{request{service,method,log_id}, compress_type,
//   correlation_id, [attachment_size], content_type, checksum_type, checksum_value} |
//   EchoRequest{message} | attachment
// log_id = i & 0x3fff, correlation_id = ((i & 0xfffff) << 32) | (i % 7 + 1)  (SURVEY §8d config 1).
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace {
uint32_t g_tab[256]; bool g_init = false;
void init_tab() {
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : (c >> 1); g_tab[i] = c; }
    g_init = true;
}
uint32_t crc32c(const uint8_t* p, size_t n) { if (!g_init) init_tab(); uint32_t l = 0xffffffffu; while (n--) l = g_tab[(l ^ *p++) & 0xff] ^ (l >> 8); return l ^ 0xffffffffu; }
uint8_t* put_varint(uint8_t* p, uint64_t v) { while (v >= 0x80) { *p++ = (uint8_t)(v | 0x80); v >>= 7; } *p++ = (uint8_t)v; return p; }
uint8_t* put_be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; return p + 4; }
uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
const char k62[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789";
}

extern "C" {

typedef struct b2press_spec {
    const char* service; const char* method;
    uint32_t payload_bytes, attachment_bytes;
    int32_t payload_kind;      /* 0 = 'r' * n (multi_threaded_echo_c++/client.cpp:116), 1 = random over 62 symbols */
    int32_t checksum_type;     /* 1 = CRC32C (-enable_checksum, echo_c++/client.cpp:76-78) */
    uint64_t seed;
} b2press_spec;

/* Bytes of frame #index (whole frame) into out; returns its length. */
size_t b2press_frame(const b2press_spec* s, uint64_t index, uint8_t* out, size_t cap) {
    const uint32_t n = s->payload_bytes, an = s->attachment_bytes;
    std::vector<uint8_t> body(8 + (size_t)n);
    uint8_t* b = body.data(); *b++ = 0x0a; b = put_varint(b, n);
    if (s->payload_kind == 0) memset(b, 'r', n);
    else { uint64_t st = s->seed ^ (index * 0x9e3779b97f4a7c15ull); for (uint32_t i = 0; i < n; i++) b[i] = (uint8_t)k62[splitmix(st) % 62]; }
    b += n;
    const size_t body_len = (size_t)(b - body.data());
    uint8_t req[512]; uint8_t* r = req;
    const size_t sl = strlen(s->service), ml = strlen(s->method);
    if (sl + ml > 400) return 0;
    *r++ = 0x0a; r = put_varint(r, sl); memcpy(r, s->service, sl); r += sl;
    *r++ = 0x12; r = put_varint(r, ml); memcpy(r, s->method, ml); r += ml;
    *r++ = 0x18; r = put_varint(r, index & 0x3fff);
    uint8_t meta[640]; uint8_t* m = meta;
    *m++ = 0x0a; m = put_varint(m, (uint64_t)(r - req)); memcpy(m, req, (size_t)(r - req)); m += r - req;
    *m++ = 0x18; *m++ = 0x00;
    *m++ = 0x20; m = put_varint(m, ((index & 0xfffff) << 32) | (index % 7 + 1));
    if (an) { *m++ = 0x28; m = put_varint(m, an); }
    *m++ = 0x50; *m++ = 0x00;
    *m++ = 0x58; *m++ = (uint8_t)s->checksum_type;
    *m++ = 0x62;
    if (s->checksum_type == 1) {
        uint32_t c = crc32c(body.data(), body_len); c = ((c >> 15) | (c << 17)) + 0xa282ead8u;
        *m++ = 4; m = put_be32(m, c);
    } else *m++ = 0;
    const size_t meta_len = (size_t)(m - meta);
    const size_t total = 12 + meta_len + body_len + an;
    if (total > cap) return 0;
    uint8_t* o = out; memcpy(o, "PRPC", 4); o += 4; o = put_be32(o, (uint32_t)(meta_len + body_len + an)); o = put_be32(o, (uint32_t)meta_len);
    memcpy(o, meta, meta_len); o += meta_len; memcpy(o, body.data(), body_len); o += body_len;
    for (uint32_t i = 0; i < an; i++) o[i] = (uint8_t)('A' + (index + i) % 26);
    return total;
}

/* Fill one socket run of exactly run_bytes: whole frames from *index on, the last one
 * truncated at run_bytes (what a read() that stops mid-frame leaves in the buffer).
 * Returns the number of COMPLETE frames; *index advances past them. */
uint64_t b2press_fill_run(const b2press_spec* s, uint64_t* index, uint8_t* out, size_t run_bytes) {
    std::vector<uint8_t> f(64 + 1024 + (size_t)s->payload_bytes + s->attachment_bytes);
    size_t pos = 0; uint64_t full = 0;
    while (pos < run_bytes) {
        const size_t n = b2press_frame(s, *index, f.data(), f.size());
        if (n == 0) break;
        const size_t take = n <= run_bytes - pos ? n : run_bytes - pos;
        memcpy(out + pos, f.data(), take);
        pos += take;
        if (take == n) { full++; (*index)++; }
    }
    return full;
}

}  // extern "C"
