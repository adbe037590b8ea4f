/* b2_oracle_h2.c — CPU ORACLE (test infrastructure): the server side of brpc's h2 parser.
 * Restates ParseH2Message (src/brpc/policy/http2_rpc_protocol.cpp:1103-1138), H2Context::Consume (:467-543),
 * ConsumeFrameHead (:438-465), the frame handlers (:545-1041), DeferWindowUpdate (:1078-1094),
 * RemoveStreamAndDeferWU (:378-392), H2StreamContext::ConsumeHeaders (:1221-1306), and the first things
 * ProcessHttpRequest looks at: ParseContentType / RemoveGrpcPrefix (policy/http_rpc_protocol.cpp:176-230, :264-277)
 * and the "/service/method" lookup of FindMethodPropertyByURIImpl (:1088-1138).
 * brpc itself cannot be built here and its tests hold no byte-level h2 server transcripts (test/brpc_h2_unittest*
 * drives client and server together), so this part is "parity unpinned" beyond the RFC 7541 HPACK vectors. */
#include "b2_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int orc_find_method_idx(const orc_config* cfg, const uint8_t* svc, uint32_t svc_len, const uint8_t* mth, uint32_t mth_len);

#define MAX_WINDOW 2147483647ll
typedef struct {
    int32_t id; int stream_ended;
    uint8_t* hdr; uint32_t hdr_len, hdr_cap, n_headers;
    uint8_t* body; uint32_t body_len, body_cap;
    int64_t remote_window_left, deferred_wu;
} h2_stream;
struct orc_h2_conn {
    int conn_state; int32_t last_received_stream_id; int remote_settings_received;
    uint32_t r_header_table_size, r_enable_push, r_max_concurrent_streams, r_stream_window_size, r_max_frame_size, r_max_header_list_size;
    uint32_t l_stream_window_size, l_max_frame_size;
    int64_t remote_window_left, deferred_window_update;
    int64_t last_sent_stream_id; int preface_sent;     /* client side: H2Context::_last_sent_stream_id (:331), ctx == NULL in AppendAndDestroySelf */
    h2_stream* streams; uint32_t n_pending, cap;      /* _pending_streams */
    orc_hpack* hp;
    /* HPacker::_encode_table: newest first */
    struct enc_entry { uint8_t* name; uint32_t nl; uint8_t* value; uint32_t vl; }* enc; uint32_t enc_count, enc_cap, enc_size, enc_max;
};
orc_h2_conn* orc_h2_conn_new(void) {                 /* H2Context::H2Context (:323-353) + Init (:363-371), server side */
    orc_h2_conn* c = (orc_h2_conn*)calloc(1, sizeof *c);
    c->last_received_stream_id = -1; c->last_sent_stream_id = 1;
    c->r_header_table_size = 4096; c->r_enable_push = 0; c->r_max_concurrent_streams = 0xffffffffu;
    c->r_stream_window_size = (uint32_t)MAX_WINDOW; c->r_max_frame_size = 16384; c->r_max_header_list_size = 0xffffffffu;
    c->l_stream_window_size = 256 * 1024; c->l_max_frame_size = 16384;
    c->remote_window_left = MAX_WINDOW;
    c->hp = orc_hpack_new(4096);
    c->enc_max = 4096;
    return c;
}
static void stream_free(h2_stream* s) { free(s->hdr); free(s->body); }
void orc_h2_conn_free(orc_h2_conn* c) {
    if (!c) return;
    for (uint32_t i = 0; i < c->n_pending; i++) stream_free(&c->streams[i]);
    for (uint32_t i = 0; i < c->enc_count; i++) { free(c->enc[i].name); free(c->enc[i].value); }
    free(c->enc);
    free(c->streams); orc_hpack_free(c->hp); free(c);
}
typedef struct { uint8_t* p; uint32_t cap, len; int ovf; } wbuf;
static uint8_t* room(wbuf* w, uint32_t n) { if (w->len + n > w->cap) { w->ovf = 1; return NULL; } uint8_t* p = w->p + w->len; w->len += n; return p; }
static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static uint32_t get32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void put_head(uint8_t* p, uint32_t payload, uint8_t type, uint8_t flags, uint32_t sid) {   /* SerializeFrameHead :123-136 */
    p[0] = (uint8_t)(payload >> 16); p[1] = (uint8_t)(payload >> 8); p[2] = (uint8_t)payload; p[3] = type; p[4] = flags; put32(p + 5, sid);
}
static void write_wu(wbuf* w, uint32_t sid, int64_t inc) { uint8_t* p = room(w, 13); if (p) { put_head(p, 4, 8, 0, sid); put32(p + 9, (uint32_t)inc); } }
static int add_window(int64_t* w, int64_t diff) {       /* AddWindowSize :261-281 */
    const int64_t before = *w; *w = before + diff;
    const int64_t mask = (int64_t)(int32_t)0x80000000;
    if ((((before | diff) >> 31) & 1) == 0) { if ((before + diff) & mask) return 0; }
    if ((((before & diff) >> 31) & 1) == 1) { if (((before + diff) & mask) == 0) return 0; }
    return 1;
}
static void defer_wu(orc_h2_conn* c, wbuf* w, int64_t size) {     /* DeferWindowUpdate :1078-1094 */
    if (size <= 0) return;
    c->deferred_window_update += size;
    if (c->deferred_window_update >= (int64_t)(c->l_stream_window_size / 2)) {
        const int64_t cw = c->deferred_window_update; c->deferred_window_update = 0;
        if (cw > 0) write_wu(w, 0, cw);
    }
}
static int find_stream(orc_h2_conn* c, int32_t id) { for (uint32_t i = 0; i < c->n_pending; i++) if (c->streams[i].id == id) return (int)i; return -1; }
/* RemoveStreamAndDeferWU :378-392; the removed stream is moved to *out (caller frees) */
static int remove_stream(orc_h2_conn* c, wbuf* w, int32_t id, h2_stream* out) {
    const int k = find_stream(c, id);
    if (k < 0) return 0;
    *out = c->streams[k];
    memmove(&c->streams[k], &c->streams[k + 1], sizeof(h2_stream) * (c->n_pending - (uint32_t)k - 1)); c->n_pending--;
    const int64_t d = out->deferred_wu; out->deferred_wu = 0;
    defer_wu(c, w, d);
    return 1;
}
static uint32_t cstrn(const uint8_t* p, uint32_t n) { uint32_t i = 0; while (i < n && p[i]) i++; return i; }
static int lit_eq(const uint8_t* a, uint32_t n, const char* lit) { const size_t l = strlen(lit); return l == n && memcmp(a, lit, l) == 0; }
static int ci_eq(const uint8_t* a, uint32_t n, const char* lit) {
    const size_t l = strlen(lit); if (l != n) return 0;
    for (size_t i = 0; i < l; i++) { uint8_t x = a[i]; if (x >= 'a' && x <= 'z') x = (uint8_t)(x - 32); if (x != (uint8_t)lit[i]) return 0; }
    return 1;
}
static int http_method(const uint8_t* v, uint32_t vl) {           /* Str2HttpMethod, http_method.cpp:104-140 */
    static const char* const names[27] = { "DELETE", "GET", "HEAD", "POST", "PUT", "CONNECT", "OPTIONS", "TRACE", "COPY", "LOCK", "MKCOL", "MOVE",
        "PROPFIND", "PROPPATCH", "SEARCH", "UNLOCK", "REPORT", "MKACTIVITY", "CHECKOUT", "MERGE", "M-SEARCH", "NOTIFY", "SUBSCRIBE",
        "UNSUBSCRIBE", "PATCH", "PURGE", "MKCALENDAR" };
    const uint32_t n = cstrn(v, vl);
    for (int m = 0; m < 27; m++) if (ci_eq(v, n, names[m])) return m;
    return -1;
}
static int starts(const uint8_t** p, uint32_t* n, const char* lit, int eat) {
    const size_t l = strlen(lit);
    if (*n < l || memcmp(*p, lit, l) != 0) return 0;
    if (eat) { *p += l; *n -= (uint32_t)l; }
    return 1;
}
static uint32_t content_type(const uint8_t* ct, uint32_t n, int* is_grpc) {     /* ParseContentType :176-230 */
    *is_grpc = 0;
    if (!starts(&ct, &n, "application/", 1)) return 0;
    if (starts(&ct, &n, "grpc", 0)) {
        if (n == 4 || ct[4] == ';') { *is_grpc = 1; return 2; }
        else if (ct[4] == '+') { ct += 5; n -= 5; *is_grpc = 1; }
    }
    uint32_t type;
    if (starts(&ct, &n, "json", 1)) type = 1;
    else if (starts(&ct, &n, "proto-json", 1)) type = 4;
    else if (starts(&ct, &n, "proto-text", 1)) type = 3;
    else if (starts(&ct, &n, "proto", 1)) type = 2;
    else if (starts(&ct, &n, "x-protobuf", 1)) type = 2;
    else return 0;
    return (n == 0 || ct[0] == ';') ? type : 0;
}
/* one header of ConsumeHeaders (:1232-1287): 0 = the reference returns -1 */
static int check_header(const uint8_t* name, uint32_t nl, const uint8_t* value, uint32_t vl) {
    const uint32_t n = cstrn(name, nl);
    if (n == 0 || name[0] != ':') return 1;
    const uint8_t c1 = n > 1 ? name[1] : 0;
    const uint8_t* rest = name + 2; const uint32_t rn = n > 2 ? n - 2 : 0;
    switch (c1) {
    case 'a': return lit_eq(rest, rn, "uthority");
    case 'm': return lit_eq(rest, rn, "ethod") && http_method(value, vl) >= 0;
    case 'p': return lit_eq(rest, rn, "ath");
    case 's':
        if (lit_eq(rest, rn, "cheme")) return 1;
        if (lit_eq(rest, rn, "tatus")) {
            char* tmp = (char*)malloc(vl + 1); memcpy(tmp, value, vl); tmp[vl] = 0;
            char* end = NULL; (void)strtol(tmp, &end, 10);
            const int ok = *end == '\0';
            free(tmp);
            return ok;
        }
        return 0;
    default: return 0;
    }
}
static int consume_headers(orc_h2_conn* c, h2_stream* st, const uint8_t* frag, uint32_t n) {
    uint8_t* name = (uint8_t*)malloc(1u << 16); uint8_t* value = (uint8_t*)malloc(1u << 16);
    uint32_t pos = 0; int rcode = 0;
    while (pos < n) {
        uint32_t nl = 0, vl = 0, adv = 0;
        const int rc = orc_hpack_field(c->hp, frag + pos, n - pos, name, &nl, value, &vl, &adv);
        if (rc < 0) { rcode = -1; break; }
        if (rc == 0) break;
        if (!check_header(name, nl, value, vl)) { rcode = -1; break; }
        if (st->hdr_len + 4 + nl + vl > st->hdr_cap) { st->hdr_cap = (st->hdr_len + 4 + nl + vl) * 2; st->hdr = (uint8_t*)realloc(st->hdr, st->hdr_cap); }
        uint8_t* r = st->hdr + st->hdr_len;
        r[0] = (uint8_t)nl; r[1] = (uint8_t)(nl >> 8); r[2] = (uint8_t)vl; r[3] = (uint8_t)(vl >> 8);
        memcpy(r + 4, name, nl); memcpy(r + 4 + nl, value, vl);
        st->hdr_len += 4 + nl + vl; st->n_headers++;
        pos += adv;
    }
    free(name); free(value);
    return rcode;
}
typedef struct { int kind; uint32_t err; int32_t err_stream; h2_stream st; } h2_res;   /* kind 0 ok, 1 message (st), 2 error */
static h2_res res_ok(void) { h2_res r; memset(&r, 0, sizeof r); return r; }
static h2_res res_err(uint32_t e, int32_t sid) { h2_res r; memset(&r, 0, sizeof r); r.kind = 2; r.err = e; r.err_stream = sid; return r; }
static h2_res end_stream(orc_h2_conn* c, wbuf* w, int32_t id) {      /* OnEndStream :823-846 */
    h2_res r = res_ok();
    if (remove_stream(c, w, id, &r.st)) r.kind = 1;
    return r;
}
static void emit_message(const orc_config* cfg, const h2_stream* st, b2_h2_msg* m, wbuf* blob) {
    memset(m, 0, sizeof *m);
    m->stream_id = (uint32_t)st->id;
    uint8_t* hp = room(blob, st->hdr_len); uint8_t* bp = room(blob, st->body_len);
    if (!hp || !bp) return;
    memcpy(hp, st->hdr, st->hdr_len); memcpy(bp, st->body, st->body_len);
    m->headers_off = (uint32_t)(hp - blob->p); m->headers_len = st->hdr_len; m->n_headers = st->n_headers;
    m->body_off = (uint32_t)(bp - blob->p); m->body_len = st->body_len;
    m->http_method = B2_H2_NO_METHOD; m->method_idx = -1;
    int is_grpc = 0; const uint8_t* path = NULL; uint32_t path_len = 0;
    for (uint32_t q = 0; q < st->hdr_len;) {
        const uint32_t nl = st->hdr[q] | ((uint32_t)st->hdr[q + 1] << 8), vl = st->hdr[q + 2] | ((uint32_t)st->hdr[q + 3] << 8);
        const uint8_t* nm = st->hdr + q + 4; const uint8_t* v = nm + nl;
        const uint32_t cn = cstrn(nm, nl);
        if (lit_eq(nm, cn, ":method")) m->http_method = (uint32_t)http_method(v, vl);
        else if (lit_eq(nm, cn, ":path")) {                          /* URI::SetH2Path, uri.cpp:403-425 */
            uint32_t e = 0; while (e < vl && v[e] && v[e] != '?' && v[e] != '#') e++;
            path = v; path_len = e; m->path_off = m->headers_off + (uint32_t)(v - st->hdr); m->path_len = e; m->flags |= B2_H2_FLAG_HAS_PATH;
        } else if (lit_eq(nm, cn, "content-type")) m->content_type = content_type(v, vl, &is_grpc);
        q += 4 + nl + vl;
    }
    if (is_grpc) {
        m->flags |= B2_H2_FLAG_GRPC;
        if (st->body_len == 0) { m->flags |= B2_H2_FLAG_GRPC_PREFIX_OK; m->msg_off = m->body_off; }        /* RemoveGrpcPrefix :264-277 */
        else if (st->body_len >= 5) {
            if (st->body[0]) m->flags |= B2_H2_FLAG_GRPC_COMPRESSED;
            if ((uint64_t)get32(st->body + 1) + 5u == st->body_len) { m->flags |= B2_H2_FLAG_GRPC_PREFIX_OK; m->msg_off = m->body_off + 5; m->msg_len = st->body_len - 5; }
        }
    }
    if (path) {                                                      /* butil::StringSplitter(path, '/') skips empty fields */
        uint32_t f0 = 0; while (f0 < path_len && path[f0] == '/') f0++;
        uint32_t e0 = f0; while (e0 < path_len && path[e0] != '/') e0++;
        uint32_t f1 = e0; while (f1 < path_len && path[f1] == '/') f1++;
        uint32_t e1 = f1; while (e1 < path_len && path[e1] != '/') e1++;
        if (e0 > f0 && e1 > f1) m->method_idx = orc_find_method_idx(cfg, path + f0, e0 - f0, path + f1, e1 - f1);
    }
}

uint32_t orc_h2_consume(orc_h2_conn* c, const orc_config* cfg, const uint8_t* in, uint32_t n, uint32_t* consumed,
                        b2_h2_msg* msgs, uint32_t msg_cap, uint32_t* n_msgs,
                        uint8_t* ctrl, uint32_t ctrl_cap, uint32_t* ctrl_len,
                        uint8_t* blob_p, uint32_t blob_cap, uint32_t* blob_len,
                        uint32_t* remote_max_frame_size, uint32_t* remote_stream_window_size) {
    wbuf w = { ctrl, ctrl_cap, 0, 0 }, blob = { blob_p, blob_cap, 0, 0 };
    uint32_t pos = 0, last_ok = 0, nm = 0, perr = B2_PARSE_ERROR_NOT_ENOUGH_DATA;
    for (;;) {
        if (w.ovf || blob.ovf) { perr = B2_PARSE_ERROR_NO_RESOURCE; break; }
        if (c->conn_state == 0) {                                    /* :469-489 */
            static const char pre[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";
            const uint32_t k = (n - pos) < 24 ? (n - pos) : 24;
            if (memcmp(pre, in + pos, k) != 0) { perr = B2_PARSE_ERROR_TRY_OTHERS; break; }
            if (k < 24) break;
            c->conn_state = 1; pos += 24;
            uint8_t* p = room(&w, 34);                               /* SerializeH2SettingsFrameAndWU(default server settings) :230-259 */
            if (p) {
                put_head(p, 12, 4, 0, 0);
                p[9] = 0; p[10] = 2; put32(p + 11, 0);
                p[15] = 0; p[16] = 4; put32(p + 17, c->l_stream_window_size);
                put_head(p + 21, 4, 8, 0, 0); put32(p + 30, 1024 * 1024 - 65535);
            }
            last_ok = pos;
            continue;
        }
        const uint32_t left = n - pos;                               /* ConsumeFrameHead :438-465 */
        if (left < 3) break;
        const uint32_t length = ((uint32_t)in[pos] << 16) | ((uint32_t)in[pos + 1] << 8) | in[pos + 2];
        if (length > c->l_max_frame_size) { perr = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        if ((uint64_t)(left - 3) < 6ull + length) break;
        const uint32_t type = in[pos + 3], flags = in[pos + 4], sid_raw = get32(in + pos + 5);
        if (sid_raw & 0x80000000u) { perr = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        const int32_t sid = (int32_t)sid_raw;
        pos += 9;
        if (type > 9) { perr = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        const uint8_t* pl = in + pos;
        uint32_t used = 0;
        h2_res res = res_ok();
        switch (type) {
        case 0: {                                                    /* OnData :700-779 */
            uint32_t frag = length, padl = 0;
            if ((flags & 0x8) && length == 0) { res = res_err(6, 0); break; }   /* deviation: the reference reads the pad length past the frame */
            if (flags & 0x8) { frag--; padl = pl[used++]; }
            if (frag < padl) { res = res_err(6, 0); break; }
            frag -= padl;
            const int k = find_stream(c, sid);
            if (k < 0) {
                used += frag + padl;
                const int64_t acc = (int64_t)frag, quota = (int64_t)(c->l_stream_window_size / (c->n_pending + 1));
                int64_t tmp_deferred = (int64_t)frag;
                if (acc >= quota && acc <= (int64_t)c->l_stream_window_size) {
                    const int64_t swu = tmp_deferred; tmp_deferred = 0;
                    if (swu > 0) { write_wu(&w, (uint32_t)sid, swu); const int64_t cw = swu + c->deferred_window_update; c->deferred_window_update = 0; write_wu(&w, 0, cw); }
                }
                defer_wu(c, &w, tmp_deferred);
                res = res_err(5, sid);
                break;
            }
            h2_stream* st = &c->streams[k];
            if (st->body_len + frag > st->body_cap) { st->body_cap = (st->body_len + frag) * 2 + 64; st->body = (uint8_t*)realloc(st->body, st->body_cap); }
            memcpy(st->body + st->body_len, pl + used, frag); st->body_len += frag; used += frag + padl;
            const int64_t acc = (int64_t)frag + st->deferred_wu; st->deferred_wu += frag;
            const int64_t quota = (int64_t)(c->l_stream_window_size / (c->n_pending + 1));
            if (acc >= quota) {
                if (acc > (int64_t)c->l_stream_window_size) { res = res_err(3, sid); break; }
                const int64_t swu = st->deferred_wu; st->deferred_wu = 0;
                if (swu > 0) { write_wu(&w, (uint32_t)sid, swu); const int64_t cw = swu + c->deferred_window_update; c->deferred_window_update = 0; write_wu(&w, 0, cw); }
            }
            if (flags & 0x1) res = end_stream(c, &w, sid);
            break; }
        case 1: {                                                    /* OnHeaders :545-655 */
            if (sid == 0) { res = res_err(1, 0); break; }
            const int has_padding = flags & 0x8, has_priority = flags & 0x20;
            if (length < (has_priority ? 5u : 0u) + (has_padding ? 1u : 0u)) { res = res_err(6, 0); break; }
            uint32_t frag = length, padl = 0;
            if (has_padding) { padl = pl[used++]; frag--; }
            if (has_priority) { used += 5; frag -= 5; }
            if (frag < padl) { res = res_err(6, 0); break; }
            frag -= padl;
            int k;
            if (sid > c->last_received_stream_id) {
                if ((sid & 1) == 0) { res = res_err(1, 0); break; }
                c->last_received_stream_id = sid;
                if (c->n_pending == c->cap) { c->cap = c->cap ? c->cap * 2 : 8; c->streams = (h2_stream*)realloc(c->streams, sizeof(h2_stream) * c->cap); }
                k = (int)c->n_pending++;
                memset(&c->streams[k], 0, sizeof(h2_stream));
                c->streams[k].id = sid; c->streams[k].remote_window_left = (int64_t)c->r_stream_window_size;
            } else {
                k = find_stream(c, sid);
                if (k < 0) { res = res_err(1, 0); break; }
            }
            h2_stream* st = &c->streams[k];
            if (consume_headers(c, st, pl + used, frag) < 0) { res = res_err(1, 0); break; }
            used += frag + padl;
            if (flags & 0x4) { if (flags & 0x1) res = end_stream(c, &w, sid); }
            else if (flags & 0x1) st->stream_ended = 1;
            break; }
        case 2: res = res_err(1, 0); break;                          /* OnPriority :917-921 */
        case 3: {                                                    /* OnResetStream :781-821 */
            if (length != 4) { res = res_err(6, 0); break; }
            used += 4;
            h2_stream dead; if (remove_stream(c, &w, sid, &dead)) stream_free(&dead);
            break; }
        case 4: {                                                    /* OnSettings :848-915 */
            if (sid != 0) { res = res_err(1, 0); break; }
            if (flags & 0x1) { if (length != 0) res = res_err(1, 0); break; }
            const int64_t old_sw = (int64_t)c->r_stream_window_size;
            uint32_t t[6];
            if (!c->remote_settings_received) { t[0] = 4096; t[1] = 0; t[2] = 0xffffffffu; t[3] = 256 * 1024; t[4] = 16384; t[5] = 0xffffffffu; }
            else { t[0] = c->r_header_table_size; t[1] = c->r_enable_push; t[2] = c->r_max_concurrent_streams; t[3] = c->r_stream_window_size; t[4] = c->r_max_frame_size; t[5] = c->r_max_header_list_size; }
            int okp = (length / 6) * 6 == length;                    /* ParseH2Settings :166-211 */
            if (okp) for (uint32_t i = 0; i < length / 6; i++) {
                const uint32_t id = ((uint32_t)pl[used] << 8) | pl[used + 1], value = get32(pl + used + 2);
                used += 6;
                if (id == 1) t[0] = value;
                else if (id == 2) { if (value > 1) { okp = 0; break; } t[1] = value; }
                else if (id == 3) t[2] = value;
                else if (id == 4) { if (value > (uint32_t)MAX_WINDOW) { okp = 0; break; } t[3] = value; }
                else if (id == 5) { if (value > 16777215u || value < 16384u) { okp = 0; break; } t[4] = value; }
                else if (id == 6) t[5] = value;
            }
            if (!c->remote_settings_received) {
                if (!okp) { res = res_err(1, 0); break; }
                c->remote_window_left -= (MAX_WINDOW - 65535);
                c->remote_settings_received = 1;
            }
            c->r_header_table_size = t[0]; c->r_enable_push = t[1]; c->r_max_concurrent_streams = t[2];
            c->r_stream_window_size = t[3]; c->r_max_frame_size = t[4]; c->r_max_header_list_size = t[5];
            if (!okp) { res = res_err(1, 0); break; }
            const int64_t diff = (int64_t)c->r_stream_window_size - old_sw;
            int flow_ok = 1;
            if (diff) for (uint32_t i = 0; i < c->n_pending; i++) if (!add_window(&c->streams[i].remote_window_left, diff)) { flow_ok = 0; break; }
            if (!flow_ok) { res = res_err(3, 0); break; }
            uint8_t* p = room(&w, 9); if (p) put_head(p, 0, 4, 1, 0);
            break; }
        case 5: res = res_err(1, 0); break;                          /* OnPushPromise :923-927 */
        case 6: {                                                    /* OnPing :929-951 */
            if (length != 8) { res = res_err(6, 0); break; }
            if (sid != 0) { res = res_err(1, 0); break; }
            if (flags & 0x1) break;
            uint8_t* p = room(&w, 17);
            if (p) { put_head(p, 8, 6, 1, 0); memcpy(p + 9, pl, 8); }
            used += 8;
            break; }
        case 7: {                                                    /* OnGoAway :958-1004 */
            if (length < 8) { res = res_err(6, 0); break; }
            if (sid != 0) { res = res_err(1, 0); break; }
            if (flags) { res = res_err(1, 0); break; }
            used += length;
            break; }
        case 8: {                                                    /* OnWindowUpdate :1006-1041 */
            if (length != 4) { res = res_err(6, 0); break; }
            const uint32_t inc = get32(pl); used += 4;
            if ((inc & 0x80000000u) || inc == 0) { res = res_err(1, 0); break; }
            if (sid == 0) { if (!add_window(&c->remote_window_left, (int64_t)inc)) res = res_err(3, 0); break; }
            const int k = find_stream(c, sid);
            if (k < 0) break;
            if (!add_window(&c->streams[k].remote_window_left, (int64_t)inc)) res = res_err(3, 0);
            break; }
        case 9: {                                                    /* OnContinuation :657-698 */
            const int k = find_stream(c, sid);
            if (k < 0) { res = res_err(1, 0); break; }
            h2_stream* st = &c->streams[k];
            used += length;
            if (consume_headers(c, st, pl, length) < 0) { res = res_err(1, 0); break; }
            if ((flags & 0x4) && st->stream_ended) res = end_stream(c, &w, sid);
            break; }
        }
        pos += used;
        if (res.kind == 2) {
            if (res.err_stream) {
                uint8_t* p = room(&w, 13);
                if (p) { put_head(p, 4, 3, 0, (uint32_t)res.err_stream); put32(p + 9, res.err); }
                h2_stream dead; if (remove_stream(c, &w, res.err_stream, &dead)) stream_free(&dead);
            } else {
                uint8_t* p = room(&w, 17);
                if (p) { put_head(p, 8, 7, 0, 0); put32(p + 9, (uint32_t)c->last_received_stream_id); put32(p + 13, res.err); }
            }
            last_ok = pos;
            continue;
        }
        last_ok = pos;
        if (res.kind == 1) {
            if (nm < msg_cap) emit_message(cfg, &res.st, &msgs[nm], &blob); else blob.ovf = 1;
            nm++;
            stream_free(&res.st);
        }
    }
    *consumed = last_ok; *n_msgs = nm; *ctrl_len = w.len; *blob_len = blob.len;
    *remote_max_frame_size = c->r_max_frame_size; *remote_stream_window_size = c->r_stream_window_size;
    return perr;
}

/* ---- response side: H2UnsentResponse::AppendAndDestroySelf (:1688-1750) + PackH2Message (:1310-1380) -------------------- */
#include "hpack_tables.h"
static uint8_t lc(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }
static int ci_same(const uint8_t* a, const uint8_t* b, uint32_t n) { for (uint32_t i = 0; i < n; i++) if (lc(a[i]) != lc(b[i])) return 0; return 1; }
/* FindHeaderFromIndexTable / FindNameFromIndexTable (hpack.cpp:676-694): static table first (names resolve to their
 * smallest index: the table is filled in reverse, :250-258), then the latest matching entry of the encode table */
static uint32_t enc_find(const orc_h2_conn* c, const uint8_t* n, uint32_t nl, const uint8_t* v, uint32_t vl, int want_value) {
    for (uint32_t i = 0; i < 61; i++) {
        if (orc_hpack_static_name[i][1] != nl || !ci_same(n, orc_hpack_static_blob + orc_hpack_static_name[i][0], nl)) continue;
        if (want_value && (vl == 0 || orc_hpack_static_value[i][1] != vl || memcmp(v, orc_hpack_static_blob + orc_hpack_static_value[i][0], vl) != 0)) continue;
        return i + 1;
    }
    for (uint32_t i = 0; i < c->enc_count; i++) {
        if (c->enc[i].nl != nl || !ci_same(n, c->enc[i].name, nl)) continue;
        if (want_value && (vl == 0 || c->enc[i].vl != vl || memcmp(v, c->enc[i].value, vl) != 0)) continue;
        return 62 + i;
    }
    return 0;
}
static void enc_add(orc_h2_conn* c, const uint8_t* n, uint32_t nl, const uint8_t* v, uint32_t vl) {       /* IndexTable::AddHeader :146-173 */
    const uint32_t es = nl + vl + 32;
    while (c->enc_count && c->enc_size + es > c->enc_max) {
        struct enc_entry* e = &c->enc[c->enc_count - 1];
        c->enc_size -= e->nl + e->vl + 32; free(e->name); free(e->value); c->enc_count--;
    }
    if (es > c->enc_max) return;
    if (c->enc_count == c->enc_cap) { c->enc_cap = c->enc_cap ? c->enc_cap * 2 : 16; c->enc = (struct enc_entry*)realloc(c->enc, sizeof(struct enc_entry) * c->enc_cap); }
    memmove(&c->enc[1], &c->enc[0], sizeof(struct enc_entry) * c->enc_count);
    c->enc[0].name = (uint8_t*)malloc(nl ? nl : 1); memcpy(c->enc[0].name, n, nl); c->enc[0].nl = nl;
    c->enc[0].value = (uint8_t*)malloc(vl ? vl : 1); memcpy(c->enc[0].value, v, vl); c->enc[0].vl = vl;
    c->enc_count++; c->enc_size += es;
}
static uint8_t* put_int(uint8_t* p, uint8_t msb, uint32_t prefix, uint32_t value) {                      /* EncodeInteger :479-496 */
    const uint32_t lim = (1u << prefix) - 1;
    if (value < lim) { *p++ = (uint8_t)(msb | value); return p; }
    value -= lim; *p++ = (uint8_t)(msb | lim);
    for (; value >= 128;) { *p++ = (uint8_t)((value & 0x7f) | 0x80); value >>= 7; }
    *p++ = (uint8_t)value;
    return p;
}
static uint8_t* encode(orc_h2_conn* c, uint8_t* p, const char* name, const uint8_t* v, uint32_t vl, int never_index) {   /* HPacker::Encode :696-726 */
    const uint8_t* n = (const uint8_t*)name; const uint32_t nl = (uint32_t)strlen(name);
    if (!never_index) {
        const uint32_t idx = enc_find(c, n, nl, v, vl, 1);
        if (idx) return put_int(p, 0x80, 7, idx);
    }
    const uint32_t name_index = enc_find(c, n, nl, NULL, 0, 0);
    if (!never_index) { enc_add(c, n, nl, v, vl); p = put_int(p, 0x40, 6, name_index); }
    else p = put_int(p, 0x10, 4, name_index);
    if (name_index == 0) { p = put_int(p, 0x00, 7, nl); for (uint32_t k = 0; k < nl; k++) *p++ = lc(n[k]); }
    p = put_int(p, 0x00, 7, vl); memcpy(p, v, vl); p += vl;
    return p;
}
uint32_t orc_h2_pack_response(orc_h2_conn* c, const b2_h2_response* R, const uint8_t* bytes, uint8_t* out) {
    uint8_t* o = out;
    const int grpc = R->flags & B2_H2_RESP_GRPC;
    const uint32_t data_size = R->body_len + (grpc ? 5u : 0u);
    if (c->remote_window_left < (int64_t)data_size) {                       /* MinusWindowSize :283-296 -> RST_STREAM(FLOW_CONTROL_ERROR) */
        put_head(o, 4, 3, 0, R->stream_id); put32(o + 9, 3); return 13;
    }
    c->remote_window_left -= (int64_t)data_size;
    const int never = c->r_header_table_size == 0;
    uint8_t* frag = (uint8_t*)malloc(4096); uint8_t* trailer = (uint8_t*)malloc(4096);
    char num[16];
    uint8_t* f = frag;
    snprintf(num, sizeof num, "%d", R->status_code);
    f = encode(c, f, ":status", (const uint8_t*)num, (uint32_t)strlen(num), never);
    if (R->content_type_len) f = encode(c, f, "content-type", bytes + R->content_type_off, R->content_type_len, never);
    uint8_t* t = trailer;
    if (grpc) {
        snprintf(num, sizeof num, "%d", R->grpc_status);
        t = encode(c, t, "grpc-status", (const uint8_t*)num, (uint32_t)strlen(num), never);
        if (R->grpc_message_len) t = encode(c, t, "grpc-message", bytes + R->grpc_message_off, R->grpc_message_len, never);
    }
    const uint32_t fl = (uint32_t)(f - frag), tl = (uint32_t)(t - trailer), mfs = c->r_max_frame_size;
    uint8_t* data = (uint8_t*)malloc(data_size ? data_size : 1);
    if (grpc) { data[0] = 0; put32(data + 1, R->body_len); memcpy(data + 5, bytes + R->body_off, R->body_len); }   /* AddGrpcPrefix */
    else memcpy(data, bytes + R->body_off, R->body_len);
    uint8_t hflags = (data_size == 0 && tl == 0) ? 0x1 : 0;
    if (fl <= mfs) { put_head(o, fl, 1, hflags | 0x4, R->stream_id); o += 9; memcpy(o, frag, fl); o += fl; }
    else {
        put_head(o, mfs, 1, hflags, R->stream_id); o += 9; memcpy(o, frag, mfs); o += mfs;
        for (uint32_t at = mfs; at < fl;) { const uint32_t nn = fl - at < mfs ? fl - at : mfs; put_head(o, nn, 9, at + nn == fl ? 0x4 : 0, R->stream_id); o += 9; memcpy(o, frag + at, nn); o += nn; at += nn; }
    }
    for (uint32_t at = 0; at < data_size;) {
        const uint32_t nn = data_size - at < mfs ? data_size - at : mfs;
        put_head(o, nn, 0, (at + nn == data_size && tl == 0) ? 0x1 : 0, R->stream_id); o += 9;
        memcpy(o, data + at, nn); o += nn; at += nn;
    }
    if (tl) { put_head(o, tl, 1, 0x5, R->stream_id); o += 9; memcpy(o, trailer, tl); o += tl; }
    if (c->deferred_window_update > 0) { const int64_t cw = c->deferred_window_update; c->deferred_window_update = 0; put_head(o, 4, 8, 0, 0); put32(o + 9, (uint32_t)cw); o += 13; }
    free(frag); free(trailer); free(data);
    return (uint32_t)(o - out);
}

/* ---- client side: H2UnsentRequest::New (:1382-1453) + AppendAndDestroySelf (:1496-1592) + PackH2Message (:1310-1380) ------------ */
void orc_h2_conn_set_next_stream_id(orc_h2_conn* c, uint32_t id) { c->last_sent_stream_id = id; }
static uint8_t* encode_n(orc_h2_conn* c, uint8_t* p, const uint8_t* n, uint32_t nl, const uint8_t* v, uint32_t vl, int never_index) {
    char name[512]; if (nl >= sizeof name) nl = sizeof name - 1;
    memcpy(name, n, nl); name[nl] = 0;                                  /* (header names hold no NUL) */
    return encode(c, p, name, v, vl, never_index);
}
int32_t orc_h2_pack_request(orc_h2_conn* c, const b2_h2_request* R, const uint8_t* bytes, uint8_t* out, uint32_t* out_len, uint32_t* stream_id) {
    uint8_t* o = out;
    *stream_id = 0;
    if (!c->preface_sent) {                                             /* ctx == NULL: preface + SerializeH2SettingsFrameAndWU(_unack_local_settings) (:1508-1526) */
        c->preface_sent = 1;
        memcpy(o, "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n", 24); o += 24;
        put_head(o, 12, 4, 0, 0);
        o[9] = 0; o[10] = 2; put32(o + 11, 0);                          /* ENABLE_PUSH = 0 (H2Settings(): enable_push false, http2.cpp:26-34) */
        o[15] = 0; o[16] = 4; put32(o + 17, 256 * 1024);                /* FLAGS_h2_client_stream_window_size */
        put_head(o + 21, 4, 8, 0, 0); put32(o + 30, 1024 * 1024 - 65535);   /* FLAGS_h2_client_connection_window_size - 65535 */
        o += 34;
    }
    *out_len = (uint32_t)(o - out);
    if (c->last_sent_stream_id > 0x7FFFFFFFll) return B2_H2_REQ_RUNOUT; /* AllocateClientStreamId, http2_rpc_protocol.h:399-412 */
    const uint32_t id = (uint32_t)c->last_sent_stream_id; c->last_sent_stream_id += 2;
    *stream_id = id;
    const int grpc = R->flags & B2_H2_REQ_GRPC;
    const uint32_t data_size = R->body_len + (grpc ? 5u : 0u);
    if (data_size) {                                                    /* ConsumeWindowSize (:1199-1219); Init (:1176-1181): the stream starts with the peer's initial window */
        if ((int64_t)c->r_stream_window_size < (int64_t)data_size) return B2_H2_REQ_ELIMIT;
        if (c->remote_window_left < (int64_t)data_size) return B2_H2_REQ_ELIMIT;
        c->remote_window_left -= (int64_t)data_size;
    }
    const int never = c->r_header_table_size == 0;
    uint8_t* frag = (uint8_t*)malloc(8192); uint8_t* f = frag;
    f = encode(c, f, ":method", (const uint8_t*)((R->flags & B2_H2_REQ_GET) ? "GET" : "POST"), (R->flags & B2_H2_REQ_GET) ? 3 : 4, never);
    f = encode(c, f, ":scheme", (const uint8_t*)((R->flags & B2_H2_REQ_HTTPS) ? "https" : "http"), (R->flags & B2_H2_REQ_HTTPS) ? 5 : 4, never);
    f = encode(c, f, ":path", bytes + R->path_off, R->path_len, never);
    f = encode(c, f, ":authority", bytes + R->authority_off, R->authority_len, never);
    if (R->content_type_len) f = encode(c, f, "content-type", bytes + R->content_type_off, R->content_type_len, never);
    if (R->flags & B2_H2_REQ_ACCEPT) f = encode(c, f, "accept", (const uint8_t*)"*/*", 3, never);
    if (R->flags & B2_H2_REQ_USER_AGENT) f = encode(c, f, "user-agent", (const uint8_t*)"brpc/1.0 curl/7.0", 17, never);
    for (uint32_t at = 0; at + 4 <= R->extra_len;) {
        const uint8_t* e = bytes + R->extra_off + at;
        const uint32_t nl = e[0] | ((uint32_t)e[1] << 8), vl = e[2] | ((uint32_t)e[3] << 8);
        if (at + 4 + nl + vl > R->extra_len) break;
        f = encode_n(c, f, e + 4, nl, e + 4 + nl, vl, never);
        at += 4 + nl + vl;
    }
    const uint32_t fl = (uint32_t)(f - frag), mfs = c->r_max_frame_size;
    const uint8_t hflags = data_size == 0 ? 0x1 : 0;                    /* PackH2Message with empty trailers */
    if (fl <= mfs) { put_head(o, fl, 1, hflags | 0x4, id); o += 9; memcpy(o, frag, fl); o += fl; }
    else {
        put_head(o, mfs, 1, hflags, id); o += 9; memcpy(o, frag, mfs); o += mfs;
        for (uint32_t at = mfs; at < fl;) { const uint32_t nn = fl - at < mfs ? fl - at : mfs; put_head(o, nn, 9, at + nn == fl ? 0x4 : 0, id); o += 9; memcpy(o, frag + at, nn); o += nn; at += nn; }
    }
    uint8_t pre[5] = {0, 0, 0, 0, 0}; put32(pre + 1, R->body_len);
    for (uint32_t at = 0; at < data_size;) {
        const uint32_t nn = data_size - at < mfs ? data_size - at : mfs;
        put_head(o, nn, 0, at + nn == data_size ? 0x1 : 0, id); o += 9;
        for (uint32_t k = 0; k < nn; k++) { const uint32_t q = at + k; o[k] = grpc ? (q < 5 ? pre[q] : bytes[R->body_off + q - 5]) : bytes[R->body_off + q]; }
        o += nn; at += nn;
    }
    if (c->deferred_window_update > 0) { const int64_t cw = c->deferred_window_update; c->deferred_window_update = 0; put_head(o, 4, 8, 0, 0); put32(o + 9, (uint32_t)cw); o += 13; }
    free(frag);
    *out_len = (uint32_t)(o - out);
    return B2_H2_REQ_OK;
}

/* the peer's SETTINGS / connection WINDOW_UPDATE as the host's parser mirrors them (OnSettings :848-915, OnWindowUpdate :1006-1041) */
int orc_h2_conn_peer_update(orc_h2_conn* c, const b2_h2_peer_update* u) {
    if (u->set & B2_H2_PEER_HEADER_TABLE_SIZE) c->r_header_table_size = u->header_table_size;
    if (u->set & B2_H2_PEER_MAX_FRAME_SIZE) { if (u->max_frame_size < 16384u || u->max_frame_size > 16777215u) return -1; c->r_max_frame_size = u->max_frame_size; }
    if (u->set & B2_H2_PEER_STREAM_WINDOW) { if (u->stream_window_size > (uint32_t)MAX_WINDOW) return -1; c->r_stream_window_size = u->stream_window_size; }
    if (u->set & B2_H2_PEER_CONN_WINDOW_ADD) {
        if (u->conn_window_add < 0) c->remote_window_left += u->conn_window_add;        /* the first SETTINGS: a plain subtraction (:884) */
        else if (!add_window(&c->remote_window_left, u->conn_window_add)) return -1;
    }
    return 0;
}
