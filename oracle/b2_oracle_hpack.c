/* b2_oracle_hpack.c — CPU ORACLE (test infrastructure): HPACK decode + h2 frame-head scan.
 * Pinned by the RFC 7541 Appendix C vectors the reference asserts (test/brpc_hpack_unittest.cpp:30-550,
 * extracted into tests/golden/hpack_vectors.json) and its fuzz seed corpus. */
#include "b2_oracle.h"
#include "hpack_tables.h"
#include <stdlib.h>
#include <string.h>

#define HP_MAX_ENTRIES 128
struct orc_hpack {
    uint32_t max_size, size;          /* IndexTable::_max_size / _size (entry = name + value + 32, hpack.cpp:117-120) */
    uint32_t count;                   /* entries, newest first at index 0 */
    struct { uint8_t* name; uint32_t nl; uint8_t* value; uint32_t vl; } e[HP_MAX_ENTRIES];
};
orc_hpack* orc_hpack_new(uint32_t max_table_size) { orc_hpack* h = (orc_hpack*)calloc(1, sizeof *h); h->max_size = max_table_size; return h; }
static void hp_pop(orc_hpack* h) {                       /* PopHeader: evict the oldest */
    uint32_t i = h->count - 1;
    h->size -= h->e[i].nl + h->e[i].vl + 32;
    free(h->e[i].name); free(h->e[i].value); h->count--;
}
void orc_hpack_free(orc_hpack* h) { if (!h) return; while (h->count) hp_pop(h); free(h); }
static int hp_add(orc_hpack* h, const uint8_t* n, uint32_t nl, const uint8_t* v, uint32_t vl) {   /* AddHeader :150-177 */
    uint32_t es = nl + vl + 32;
    if (nl == 0) return -1;                               /* reference CHECK-aborts on an empty name */
    while (h->count && h->size + es > h->max_size) hp_pop(h);
    if (es > h->max_size) return 0;
    if (h->count >= HP_MAX_ENTRIES) return -1;
    memmove(&h->e[1], &h->e[0], sizeof h->e[0] * h->count);
    h->e[0].name = (uint8_t*)malloc(nl ? nl : 1); memcpy(h->e[0].name, n, nl); h->e[0].nl = nl;
    h->e[0].value = (uint8_t*)malloc(vl ? vl : 1); memcpy(h->e[0].value, v, vl); h->e[0].vl = vl;
    h->count++; h->size += es;
    return 0;
}
/* HeaderAt: 1..61 static, 62.. dynamic (newest first) */
static int hp_at(const orc_hpack* h, uint32_t index, const uint8_t** n, uint32_t* nl, const uint8_t** v, uint32_t* vl) {
    if (index >= 1 && index <= 61) {
        *n = orc_hpack_static_blob + orc_hpack_static_name[index - 1][0]; *nl = orc_hpack_static_name[index - 1][1];
        *v = orc_hpack_static_blob + orc_hpack_static_value[index - 1][0]; *vl = orc_hpack_static_value[index - 1][1];
        return 1;
    }
    if (index >= 62 && index - 62 < h->count) {
        *n = h->e[index - 62].name; *nl = h->e[index - 62].nl; *v = h->e[index - 62].value; *vl = h->e[index - 62].vl;
        return 1;
    }
    return 0;
}
/* DecodeInteger :531-565.  >0 bytes used, 0 not enough data, -1 malformed */
static int hp_int(const uint8_t* p, uint32_t n, uint32_t prefix, uint32_t* value) {
    if (n == 0) return 0;
    uint64_t tmp = p[0] & ((1u << prefix) - 1);
    if (tmp < ((1u << prefix) - 1)) { *value = (uint32_t)tmp; return 1; }
    uint32_t i = 1; int m = 0; uint8_t cur;
    do {
        if (i >= n) return 0;
        cur = p[i++];
        tmp += (uint64_t)(cur & 0x7f) << m;
        m += 7;
    } while ((cur & 0x80) && tmp < 10u * 1024 * 1024);
    if (tmp >= 10u * 1024 * 1024) return -1;
    *value = (uint32_t)tmp;
    return (int)i;
}
/* Huffman: walk the code bit by bit (HuffmanDecoder::Decode/EndStream :414-468) */
static int hp_huff(const uint8_t* p, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* olen) {
    uint32_t code = 0, depth = 0, o = 0; int padding = 1;
    for (uint32_t i = 0; i < n; i++)
        for (int b = 7; b >= 0; b--) {
            const uint32_t bit = (p[i] >> b) & 1;
            code = (code << 1) | bit; depth++;
            int sym = -1;
            for (int s = 0; s < 257; s++) if (orc_hpack_huff_len[s] == depth && orc_hpack_huff_code[s] == code) { sym = s; break; }
            if (sym >= 0) {
                if (sym == 256) return -1;                     /* EOS inside the stream */
                if (o >= cap) return -1;
                out[o++] = (uint8_t)sym; code = 0; depth = 0; padding = 1;
                continue;
            }
            if (depth >= 30) return -1;                        /* fell off the tree (NULL_NODE) */
            padding = padding && bit;
        }
    if (depth == 0 || (depth <= 7 && padding)) { *olen = o; return 0; }
    return -1;
}
/* DecodeString :606-635.  >0 bytes used, 0 not enough data, -1 error */
static int hp_str(const uint8_t* p, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* olen) {
    if (n == 0) return 0;
    const int huffman = p[0] & 0x80;
    uint32_t length = 0;
    int ib = hp_int(p, n, 7, &length);
    if (ib <= 0) return -1;
    if (length > n - (uint32_t)ib) return 0;
    if (!huffman) { if (length > cap) return -1; memcpy(out, p + ib, length); *olen = length; return ib + (int)length; }
    if (hp_huff(p + ib, length, out, cap, olen) != 0) return -1;
    return ib + (int)length;
}
#define HP_STR_CAP (1u << 16)
/* HPacker::Decode :765-843: one field.  Returns bytes consumed by the field (>0), 0, or -1; *adv is how
 * far the iterator moved (a table size update moves it without producing a field). */
int orc_hpack_field(orc_hpack* h, const uint8_t* p, uint32_t n, uint8_t* name, uint32_t* nl, uint8_t* value, uint32_t* vl, uint32_t* adv) {
    *adv = 0;
    if (n == 0) return 0;
    const uint8_t fb = p[0];
    uint32_t index = 0;
    if (fb & 0x80) {
        int ib = hp_int(p, n, 7, &index);
        if (ib <= 0) return ib;
        const uint8_t *sn, *sv; uint32_t snl, svl;
        if (!hp_at(h, index, &sn, &snl, &sv, &svl)) return -1;
        memcpy(name, sn, snl); *nl = snl; memcpy(value, sv, svl); *vl = svl;
        *adv = (uint32_t)ib; return ib;
    }
    if ((fb >> 5) == 1) {                                    /* 001x: dynamic table size update, then the next field */
        uint32_t max_size = 0;
        int ib = hp_int(p, n, 5, &max_size);
        if (ib <= 0) return ib;
        if (max_size > 4096) return -1;
        if (max_size > h->max_size) h->max_size = max_size;
        else if (max_size < h->max_size) { h->max_size = max_size; while (h->size > h->max_size) hp_pop(h); }
        uint32_t a2 = 0;
        int rc = orc_hpack_field(h, p + ib, n - (uint32_t)ib, name, nl, value, vl, &a2);
        *adv = (uint32_t)ib + a2;
        return rc;
    }
    const int incremental = (fb >> 6) == 1;
    const uint32_t prefix = incremental ? 6 : 4;             /* 01xx / 0001 / 0000 */
    int ib = hp_int(p, n, prefix, &index);
    if (ib <= 0) return -1;
    uint32_t used = (uint32_t)ib;
    if (index != 0) {
        const uint8_t *sn, *sv; uint32_t snl, svl;
        if (!hp_at(h, index, &sn, &snl, &sv, &svl)) return -1;
        memcpy(name, sn, snl); *nl = snl;
    } else {
        int nb = hp_str(p + used, n - used, name, HP_STR_CAP, nl);
        if (nb <= 0) return -1;
        used += (uint32_t)nb;
        for (uint32_t i = 0; i < *nl; i++) if (name[i] >= 'A' && name[i] <= 'Z') name[i] = (uint8_t)(name[i] + 32);   /* tolower :755 */
    }
    int vb = hp_str(p + used, n - used, value, HP_STR_CAP, vl);
    if (vb <= 0) return -1;
    used += (uint32_t)vb;
    if (incremental && hp_add(h, name, *nl, value, *vl) != 0) return -1;
    *adv = used; return (int)used;
}
int orc_hpack_decode_block(orc_hpack* h, const uint8_t* in, uint32_t n, uint8_t* out, uint32_t out_cap,
                           uint32_t* out_len, uint32_t* n_headers) {
    uint8_t* name = (uint8_t*)malloc(HP_STR_CAP); uint8_t* value = (uint8_t*)malloc(HP_STR_CAP);
    uint32_t pos = 0, o = 0, cnt = 0; int status = 0;
    while (pos < n) {
        uint32_t nl = 0, vl = 0, adv = 0;
        int rc = orc_hpack_field(h, in + pos, n - pos, name, &nl, value, &vl, &adv);
        if (rc < 0) { status = -1; break; }
        if (rc == 0) { status = 1; break; }
        if (o + 4 + nl + vl > out_cap) { status = -1; break; }
        out[o] = (uint8_t)nl; out[o + 1] = (uint8_t)(nl >> 8); out[o + 2] = (uint8_t)vl; out[o + 3] = (uint8_t)(vl >> 8);
        memcpy(out + o + 4, name, nl); memcpy(out + o + 4 + nl, value, vl);
        o += 4 + nl + vl; cnt++; pos += adv;
    }
    free(name); free(value);
    *out_len = o; *n_headers = cnt;
    return status;
}

uint32_t orc_h2_scan(const uint8_t* in, uint32_t n, uint32_t max_frame_size, orc_h2_frame* frames, uint32_t cap,
                     uint32_t* consumed, uint32_t* err) {
    uint32_t pos = 0, cnt = 0;
    *err = B2_PARSE_ERROR_NOT_ENOUGH_DATA;
    for (;;) {
        if (n - pos < 3) break;
        const uint32_t length = ((uint32_t)in[pos] << 16) | ((uint32_t)in[pos + 1] << 8) | in[pos + 2];
        if (length > max_frame_size) { *err = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        if ((uint64_t)(n - pos - 3) < 6ull + length) break;
        const uint32_t sid = ((uint32_t)in[pos + 5] << 24) | ((uint32_t)in[pos + 6] << 16) | ((uint32_t)in[pos + 7] << 8) | in[pos + 8];
        if (sid & 0x80000000u) { *err = B2_PARSE_ERROR_ABSOLUTELY_WRONG; break; }
        if (cnt < cap) { frames[cnt].type = in[pos + 3]; frames[cnt].flags = in[pos + 4]; frames[cnt].pad = 0; frames[cnt].stream_id = sid; frames[cnt].payload_off = pos + 9; frames[cnt].payload_len = length; }
        cnt++; pos += 9 + length;
    }
    *consumed = pos;
    return cnt;
}
