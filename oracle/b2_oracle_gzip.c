/*
 * b2_oracle_gzip.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * What brpc's gzip / zlib CompressHandlers hand to the protobuf parser when a request or response body arrives
 * with compress_type GZIP / ZLIB:
 *     policy::GzipDecompress / ZlibDecompress(const IOBuf&, Message*)   src/brpc/policy/gzip_compress.cpp:75-89,99-101,171-174
 *       = google::protobuf::io::GzipInputStream(IOBufAsZeroCopyInputStream(data), GZIP | ZLIB) + msg->ParseFromZeroCopyStream
 *
 * Neither protobuf (MODULE.bazel pins 27.3) nor zlib is part of /root/reference, so this file restates the two published
 * algorithms the call rests on:
 *   - GzipInputStream::Next / Inflate (protobuf src/google/protobuf/io/gzip_stream.cc): one inflate() call per Next() into a
 *     fresh 64 KiB output buffer, windowBits 15|16 (GZIP) or 15 (ZLIB); a call that ends in an error hands NOTHING of that call
 *     to the parser and ends the stream (the parser sees end-of-input, not an error); a truncated input ends the stream after
 *     everything decoded so far was handed over; after Z_STREAM_END the stream is re-initialised and further members are read
 *     (trailing bytes that are no member end the stream silently).
 *   - zlib's inflate() (RFC 1950 / 1951 / 1952 decoding with zlib's own acceptance rules: which code-length sets are
 *     rejected, where the distance check sits relative to a full output buffer, header / trailer checks).
 * The input is taken as ONE block (the body of a message lives in one read block here); with several IOBuf blocks the
 * reference's call boundaries — and with them what survives an erroring call — would depend on the block layout.
 *
 * Pinned by tests/test_oracle_gzip.py against the system zlib driven exactly like GzipInputStream drives it (ctypes), on valid
 * streams of every block type and on corrupted / truncated / concatenated ones.
 */
#include <stdlib.h>
#include <string.h>
#include "b2_oracle.h"

#define GZ_CHUNK 65536u        /* GzipInputStream kDefaultBufferSize */

typedef struct gz_state {
    const uint8_t* in; size_t n; size_t bitpos;          /* input as a bit string (LSB first inside a byte) */
    uint8_t* out; size_t produced, cap;                  /* everything decoded so far */
    size_t delivered;                                    /* what earlier inflate() calls already handed to the parser */
    uint32_t chunk_fill;                                 /* bytes written by the current inflate() call */
    size_t member_start;                                 /* produced at the start of the current member */
    int sizing; size_t limit;                            /* orc_gzip_sizing_bound: no data checks, stop beyond `limit` bytes */
} gz_state;

enum { GZ_OK = 0, GZ_TRUNC = 1, GZ_ERR = 2, GZ_END = 3 };

static size_t bits_left(const gz_state* s) { return s->n * 8 - s->bitpos; }
static uint32_t peek(const gz_state* s, int k) {           /* k <= 32 bits, caller checked availability */
    uint64_t v = 0; size_t byte = s->bitpos >> 3; int sh = (int)(s->bitpos & 7);
    for (int i = 0; i < 6 && byte + i < s->n; i++) v |= (uint64_t)s->in[byte + i] << (8 * i);
    v >>= sh;
    return k == 32 ? (uint32_t)v : (uint32_t)(v & ((1ull << k) - 1));
}
/* NEEDBITS as zlib pulls them: bytes at a time, so "k bits needed" means the byte holding bit k-1 must exist */
static int need(const gz_state* s, int k) { return bits_left(s) >= (size_t)k; }
static void drop(gz_state* s, int k) { s->bitpos += (size_t)k; }
static void byte_align(gz_state* s) { s->bitpos = (s->bitpos + 7) & ~(size_t)7; }

static int put_byte(gz_state* s, uint8_t b) {
    if (s->chunk_fill == GZ_CHUNK) { s->delivered = s->produced; s->chunk_fill = 0; }     /* left == 0: the call returns, Next() hands the chunk over */
    if (s->sizing && s->produced >= s->limit) { s->produced = s->limit + 1; return -1; }
    if (s->produced == s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 4096; uint8_t* p = (uint8_t*)realloc(s->out, nc); if (!p) return -1;
        s->out = p; s->cap = nc;
    }
    s->out[s->produced++] = b; s->chunk_fill++;
    return 0;
}

/* zlib inflate_table's acceptance rules on a set of code lengths (inftrees.c): over-subscribed sets are rejected, incomplete
 * sets too except a single 1-bit code for the literal/length and distance alphabets.  count[] / symbol[] = canonical decoding
 * tables (RFC 1951 3.2.2).  kind: 0 = code-length alphabet, 1 = literal/length, 2 = distance.  Returns 0 ok, -1 rejected;
 * *single = 1 when the set is the accepted incomplete one (then only the code "0" exists) and *empty = 1 for no codes at all */
typedef struct gz_huff { uint16_t count[16]; uint16_t symbol[288]; int single, empty; } gz_huff;
static int build(gz_huff* h, const uint8_t* lens, int n, int kind) {
    memset(h->count, 0, sizeof h->count); h->single = 0; h->empty = 0;
    for (int i = 0; i < n; i++) h->count[lens[i]]++;
    int max = 15; while (max >= 1 && h->count[max] == 0) max--;
    if (max == 0) { h->empty = 1; h->count[0] = (uint16_t)n; return 0; }    /* "no symbols to code at all": a table of invalid codes, not an error */
    int left = 1;
    for (int len = 1; len <= 15; len++) { left <<= 1; left -= h->count[len]; if (left < 0) return -1; }
    if (left > 0 && (kind == 0 || max != 1)) return -1;
    if (left > 0) h->single = 1;
    uint16_t offs[16]; offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + h->count[len]);
    for (int i = 0; i < n; i++) if (lens[i]) h->symbol[offs[lens[i]]++] = (uint16_t)i;
    return 0;
}
/* one symbol.  GZ_OK + *sym, GZ_TRUNC (input ends inside the code), GZ_ERR (a code the table does not hold: only possible for
 * the incomplete / empty sets, which zlib marks invalid after ONE bit) */
static int decode(gz_state* s, const gz_huff* h, int* sym) {
    if (h->empty || h->single) {
        if (!need(s, 1)) return GZ_TRUNC;
        if (h->empty) return GZ_ERR;
        if (peek(s, 1)) return GZ_ERR;
        drop(s, 1); *sym = h->symbol[0]; return GZ_OK;
    }
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        if (!need(s, len)) return GZ_TRUNC;
        code |= (int)((peek(s, len) >> (len - 1)) & 1);
        const int count = h->count[len];
        if (code - count < first) { drop(s, len); *sym = h->symbol[index + (code - first)]; return GZ_OK; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return GZ_ERR;   /* unreachable for a complete set */
}

static const uint16_t kLenBase[29] = { 3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258 };
static const uint8_t  kLenExtra[29] = { 0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0 };
static const uint16_t kDistBase[30] = { 1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577 };
static const uint8_t  kDistExtra[30] = { 0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13 };

/* literal/length + distance symbols of one block (zlib LEN .. MATCH / LIT states) */
static int codes(gz_state* s, const gz_huff* lc, const gz_huff* dc) {
    for (;;) {
        int sym, r = decode(s, lc, &sym);
        if (r != GZ_OK) return r;
        if (sym < 256) { if (put_byte(s, (uint8_t)sym)) return GZ_ERR; continue; }
        if (sym == 256) return GZ_OK;
        if (sym >= 286) return GZ_ERR;                                          /* "invalid literal/length code" (fixed table's 286/287) */
        sym -= 257;
        if (!need(s, kLenExtra[sym])) return GZ_TRUNC;
        const uint32_t len = kLenBase[sym] + peek(s, kLenExtra[sym]); drop(s, kLenExtra[sym]);
        int ds; r = decode(s, dc, &ds);
        if (r != GZ_OK) return r;
        if (ds >= 30) return GZ_ERR;                                            /* "invalid distance code" */
        if (!need(s, kDistExtra[ds])) return GZ_TRUNC;
        const uint32_t dist = kDistBase[ds] + peek(s, kDistExtra[ds]); drop(s, kDistExtra[ds]);
        /* MATCH: `if (left == 0) goto inf_leave` sits in front of the distance check, so a full buffer is handed over first */
        if (s->chunk_fill == GZ_CHUNK) { s->delivered = s->produced; s->chunk_fill = 0; }
        if (dist > s->produced - s->member_start) return GZ_ERR;                /* "invalid distance too far back" */
        for (uint32_t k = 0; k < len; k++) if (put_byte(s, s->out[s->produced - dist])) return GZ_ERR;
    }
}

static int inflate_blocks(gz_state* s) {
    static const uint8_t order[19] = { 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15 };
    gz_huff lc, dc, cl;
    for (;;) {
        if (!need(s, 3)) return GZ_TRUNC;
        const uint32_t hdr = peek(s, 3); drop(s, 3);
        const int last = (int)(hdr & 1), type = (int)(hdr >> 1);
        if (type == 0) {
            byte_align(s);
            if (!need(s, 32)) return GZ_TRUNC;
            const uint32_t v = peek(s, 32);
            if ((v & 0xffff) != ((v >> 16) ^ 0xffff)) return GZ_ERR;           /* "invalid stored block lengths" */
            drop(s, 32);
            uint32_t len = v & 0xffff;
            while (len) {
                if (!need(s, 8)) return GZ_TRUNC;
                if (put_byte(s, (uint8_t)peek(s, 8))) return GZ_ERR;
                drop(s, 8); len--;
            }
        } else if (type == 1) {
            uint8_t lens[288];
            for (int i = 0; i < 144; i++) lens[i] = 8;
            for (int i = 144; i < 256; i++) lens[i] = 9;
            for (int i = 256; i < 280; i++) lens[i] = 7;
            for (int i = 280; i < 288; i++) lens[i] = 8;
            build(&lc, lens, 288, 1);
            for (int i = 0; i < 32; i++) lens[i] = 5;
            build(&dc, lens, 32, 2);
            const int r = codes(s, &lc, &dc); if (r != GZ_OK) return r;
        } else if (type == 2) {
            if (!need(s, 14)) return GZ_TRUNC;
            const uint32_t v = peek(s, 14); drop(s, 14);
            const int nlen = (int)(v & 31) + 257, ndist = (int)((v >> 5) & 31) + 1, ncode = (int)((v >> 10) & 15) + 4;
            if (nlen > 286 || ndist > 30) return GZ_ERR;                         /* "too many length or distance symbols" */
            uint8_t lens[320]; memset(lens, 0, sizeof lens);
            uint8_t cll[19]; memset(cll, 0, sizeof cll);
            for (int i = 0; i < ncode; i++) { if (!need(s, 3)) return GZ_TRUNC; cll[order[i]] = (uint8_t)peek(s, 3); drop(s, 3); }
            if (build(&cl, cll, 19, 0)) return GZ_ERR;                           /* "invalid code lengths set" */
            int have = 0;
            while (have < nlen + ndist) {
                int sym;
                if (cl.empty) { if (!need(s, 1)) return GZ_TRUNC; drop(s, 1); sym = 0; }   /* zlib reads the invalid-code entry's val (0) without looking at its op */
                else {
                    /* the repeat codes' extra bits are pulled together with the code (NEEDBITS(here.bits + n)) before anything is checked */
                    const size_t save = s->bitpos;
                    const int r = decode(s, &cl, &sym); if (r != GZ_OK) return r;
                    if (sym >= 16) {
                        const int eb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
                        if (!need(s, eb)) { s->bitpos = save; return GZ_TRUNC; }
                    }
                }
                if (sym < 16) { lens[have++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (have == 0) return GZ_ERR; val = lens[have - 1]; rep = 3 + (int)peek(s, 2); drop(s, 2); }     /* "invalid bit length repeat" */
                else if (sym == 17) { rep = 3 + (int)peek(s, 3); drop(s, 3); }
                else { rep = 11 + (int)peek(s, 7); drop(s, 7); }
                if (have + rep > nlen + ndist) return GZ_ERR;                    /* "invalid bit length repeat" */
                while (rep--) lens[have++] = (uint8_t)val;
            }
            if (lens[256] == 0) return GZ_ERR;                                   /* "invalid code -- missing end-of-block" */
            if (build(&lc, lens, nlen, 1)) return GZ_ERR;                        /* "invalid literal/lengths set" */
            if (build(&dc, lens + nlen, ndist, 2)) return GZ_ERR;                /* "invalid distances set" */
            const int r = codes(s, &lc, &dc); if (r != GZ_OK) return r;
        } else return GZ_ERR;                                                    /* "invalid block type" */
        if (last) return GZ_OK;
    }
}

static uint32_t crc32_ieee(const uint8_t* p, size_t n, uint32_t crc) {
    crc = ~crc;
    for (size_t i = 0; i < n; i++) { crc ^= p[i]; for (int k = 0; k < 8; k++) crc = (crc >> 1) ^ (0xedb88320u & (0u - (crc & 1))); }
    return ~crc;
}
static uint32_t adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; i++) { a = (a + p[i]) % 65521u; b = (b + a) % 65521u; }
    return (b << 16) | a;
}

/* one member: header, deflate blocks, trailer (zlib HEAD .. DONE) */
static int member(gz_state* s, int format) {
    s->member_start = s->produced; s->chunk_fill = 0;
    const size_t head = s->bitpos >> 3;
    if (format == B2_COMPRESS_TYPE_GZIP) {                                       /* RFC 1952; windowBits 15|16: only the gzip wrapper is accepted */
        if (!need(s, 16)) return GZ_TRUNC;
        if (peek(s, 16) != 0x8b1f) return GZ_ERR;                               /* "incorrect header check" */
        drop(s, 16);
        if (!need(s, 16)) return GZ_TRUNC;
        const uint32_t flags = peek(s, 16); drop(s, 16);
        if ((flags & 0xff) != 8) return GZ_ERR;                                 /* "unknown compression method" */
        if (flags & 0xe000) return GZ_ERR;                                      /* "unknown header flags set" */
        if (!need(s, 32)) return GZ_TRUNC;
        drop(s, 32);                                                            /* mtime */
        if (!need(s, 16)) return GZ_TRUNC;
        drop(s, 16);                                                            /* xfl, os */
        if (flags & 0x0400) {
            if (!need(s, 16)) return GZ_TRUNC;
            const uint32_t xlen = peek(s, 16); drop(s, 16);
            if (bits_left(s) < (size_t)xlen * 8) { s->bitpos = s->n * 8; return GZ_TRUNC; }
            drop(s, (int)xlen * 8);
        }
        for (int f = 0x0800; f <= 0x1000; f <<= 1) if (flags & f) {              /* FNAME, FCOMMENT: zero-terminated */
            for (;;) { if (!need(s, 8)) return GZ_TRUNC; const uint32_t c = peek(s, 8); drop(s, 8); if (!c) break; }
        }
        if (flags & 0x0200) {
            if (!need(s, 16)) return GZ_TRUNC;
            const size_t here = s->bitpos >> 3;
            if (peek(s, 16) != (crc32_ieee(s->in + head, here - head, 0) & 0xffff)) return GZ_ERR;   /* "header crc mismatch" */
            drop(s, 16);
        }
    } else {                                                                     /* RFC 1950 */
        if (!need(s, 16)) return GZ_TRUNC;
        const uint32_t h = peek(s, 16);
        if ((((h & 0xff) << 8) + (h >> 8)) % 31) return GZ_ERR;                 /* "incorrect header check" */
        if ((h & 0xf) != 8) return GZ_ERR;                                      /* "unknown compression method" */
        if (((h >> 4) & 0xf) + 8 > 15) return GZ_ERR;                           /* "invalid window size" */
        drop(s, 16);
        if (h & 0x2000) {                                                       /* FDICT: inflate() returns Z_NEED_DICT, which GzipInputStream treats as an error */
            if (!need(s, 32)) return GZ_TRUNC;
            return GZ_ERR;
        }
    }
    const int r = inflate_blocks(s);
    if (r != GZ_OK) return r;
    byte_align(s);
    if (!need(s, 32)) return GZ_TRUNC;
    const uint32_t t = peek(s, 32);
    const uint8_t* mo = s->out + s->member_start; const size_t mn = s->produced - s->member_start;
    if (format == B2_COMPRESS_TYPE_GZIP) {
        if (!s->sizing && t != crc32_ieee(mo, mn, 0)) return GZ_ERR;            /* "incorrect data check" */
        drop(s, 32);
        if (!need(s, 32)) return GZ_TRUNC;
        if (peek(s, 32) != (uint32_t)(mn & 0xffffffffu)) return GZ_ERR;         /* "incorrect length check" */
        drop(s, 32);
    } else {
        const uint32_t be = (t >> 24) | ((t >> 8) & 0xff00) | ((t << 8) & 0xff0000) | (t << 24);
        if (!s->sizing && be != adler32(mo, mn)) return GZ_ERR;                 /* "incorrect data check" */
        drop(s, 32);
    }
    return GZ_END;
}

/* The bytes GzipInputStream(format) yields for `in` before it reports end-of-stream.  *out is malloc'ed (free with
 * orc_free); returns 0, or -1 when out of memory. */
int orc_gzip_input_stream(const uint8_t* in, size_t n, int format, uint8_t** out, size_t* out_len) {
    gz_state s; memset(&s, 0, sizeof s); s.in = in; s.n = n;
    for (;;) {
        /* GzipInputStream::Inflate: a call with no input left asks the sub-stream, which is at its end */
        if ((s.bitpos >> 3) >= n && (s.bitpos & 7) == 0) { s.delivered = s.produced; break; }
        const int r = member(&s, format);
        if (r == GZ_ERR) break;                                                  /* this call's chunk is never handed over */
        s.delivered = s.produced;
        if (r == GZ_TRUNC) break;
    }
    *out = s.out; *out_len = s.delivered;
    if (!s.out) *out = (uint8_t*)malloc(1);
    return *out ? 0 : -1;
}
void orc_free(void* p) { free(p); }

/* The DEVICE's sizing pass (brpc_b200/csrc/b2_inflate.cuh, gz_input_stream<false>): the same walk without the data checks, which
 * need the bytes — an upper bound of what the stream hands over, used to reserve the reply slot.  Returns the bound, or limit + 1
 * as soon as the stream produces more than `limit` bytes (such a body is left to the host). */
size_t orc_gzip_sizing_bound(const uint8_t* in, size_t n, int format, size_t limit) {
    gz_state s; memset(&s, 0, sizeof s); s.in = in; s.n = n; s.sizing = 1; s.limit = limit;
    for (;;) {
        if ((s.bitpos >> 3) >= n && (s.bitpos & 7) == 0) break;
        const int r = member(&s, format);
        if (s.produced > limit) break;
        if (r == GZ_ERR || r == GZ_TRUNC) break;
    }
    free(s.out);
    return s.produced;
}
