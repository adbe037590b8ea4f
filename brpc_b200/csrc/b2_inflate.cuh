// b2_inflate.cuh — what brpc's gzip / zlib CompressHandlers hand to the protobuf parser, on the device.
//
//   policy::GzipDecompress / ZlibDecompress(const IOBuf&, Message*)   src/brpc/policy/gzip_compress.cpp:75-89,99-101,171-174
//     = google::protobuf::io::GzipInputStream(data, GZIP | ZLIB) + ParseFromZeroCopyStream
//
// protobuf and zlib are third-party to the reference; the behaviour restated here is theirs: GzipInputStream runs ONE inflate()
// per Next() into a fresh 64 KiB buffer (windowBits 15|16 for GZIP, 15 for ZLIB); a call that ends in an error hands nothing of
// that call to the parser and ends the stream silently, a truncated input ends it after everything decoded so far was handed
// over, and after a member's end further members are read.  inflate()'s own acceptance rules (which code-length sets it rejects,
// when a full output buffer returns before a check) decide what "that call" holds.  The body is one block (it lives in one
// read block here).
//
// One thread decodes (DEFLATE is a serial bit stream); requests of this kind are rare on the path this library accelerates and are
// served by the slow-path kernel, one warp per message, so many of them still decode side by side.  kWrite = false is the sizing pass
// the decode stage runs to reserve the reply slot: no bytes are produced and the trailer checks (which need them) count as passed, so
// it returns an upper bound of what the real pass hands over.
#pragma once
#include <stdint.h>

namespace b2 {

constexpr uint32_t kGzChunk = 65536;                 // GzipInputStream kDefaultBufferSize
// bodies beyond these are left to the host (B2_MSG_UNSUPPORTED): one thread decodes, so the work per message is bounded
constexpr uint32_t kGzMaxIn = 1u << 20, kGzMaxOut = 1u << 20;

struct GzHuff { uint16_t count[16]; uint16_t symbol[288]; uint8_t single, empty; };

struct GzState {
    const uint8_t* in; uint32_t n, pos;              // next input byte
    uint64_t hold; uint32_t nbits;                   // bit buffer (LSB first), as zlib keeps it
    uint8_t* out; uint32_t cap;
    uint32_t produced, delivered, chunk_fill, member_start;
    bool overflow;                                   // more than cap / kGzMaxOut bytes
};
enum { kGzOk = 0, kGzTrunc = 1, kGzErr = 2, kGzEnd = 3 };

__device__ __forceinline__ bool gz_need(GzState& s, uint32_t k) {       // NEEDBITS(k), k <= 32
    while (s.nbits < k) { if (s.pos >= s.n) return false; s.hold |= (uint64_t)s.in[s.pos++] << s.nbits; s.nbits += 8; }
    return true;
}
__device__ __forceinline__ uint32_t gz_peek(const GzState& s, uint32_t k) { return k >= 32 ? (uint32_t)s.hold : (uint32_t)s.hold & ((1u << k) - 1u); }
__device__ __forceinline__ void gz_drop(GzState& s, uint32_t k) { s.hold >>= k; s.nbits -= k; }
__device__ __forceinline__ void gz_align(GzState& s) { gz_drop(s, s.nbits & 7u); }
__device__ __forceinline__ uint32_t gz_byte_pos(const GzState& s) { return s.pos - (s.nbits >> 3); }

template <bool kWrite>
__device__ __forceinline__ void gz_put(GzState& s, uint32_t b) {
    if (s.chunk_fill == kGzChunk) { s.delivered = s.produced; s.chunk_fill = 0; }    // left == 0: inflate() returns, Next() hands the chunk over
    if (s.produced >= s.cap) { s.overflow = true; return; }
    if (kWrite) s.out[s.produced] = (uint8_t)b;
    s.produced++; s.chunk_fill++;
}

// inflate_table's verdict on a set of code lengths (zlib inftrees.c): over-subscribed sets are rejected, incomplete ones too except a
// single 1-bit code in the literal/length and distance alphabets; no codes at all is a table of invalid codes, not an error.
// kind: 0 code lengths, 1 literal/length, 2 distance
__device__ __noinline__ bool gz_build(GzHuff& h, const uint8_t* lens, int n, int kind) {
    for (int i = 0; i < 16; i++) h.count[i] = 0;
    h.single = 0; h.empty = 0;
    for (int i = 0; i < n; i++) h.count[lens[i]]++;
    int max = 15; while (max >= 1 && h.count[max] == 0) max--;
    if (max == 0) { h.empty = 1; return true; }
    int left = 1;
    for (int len = 1; len <= 15; len++) { left <<= 1; left -= h.count[len]; if (left < 0) return false; }
    if (left > 0 && (kind == 0 || max != 1)) return false;
    if (left > 0) h.single = 1;
    uint16_t offs[16]; offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + h.count[len]);
    for (int i = 0; i < n; i++) if (lens[i]) h.symbol[offs[lens[i]]++] = (uint16_t)i;
    return true;
}
// one symbol (canonical decoding, RFC 1951 3.2.2); the incomplete / empty sets answer after ONE bit like zlib's invalid-code entries
__device__ __forceinline__ int gz_decode(GzState& s, const GzHuff& h, int& sym) {
    if (h.empty | h.single) {
        if (!gz_need(s, 1)) return kGzTrunc;
        if (h.empty || (s.hold & 1)) return kGzErr;
        gz_drop(s, 1); sym = h.symbol[0]; return kGzOk;
    }
    int code = 0, first = 0, index = 0;
    for (uint32_t len = 1; len <= 15; len++) {
        if (!gz_need(s, len)) return kGzTrunc;
        code |= (int)((s.hold >> (len - 1)) & 1);
        const int count = h.count[len];
        if (code - count < first) { gz_drop(s, len); sym = h.symbol[index + (code - first)]; return kGzOk; }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return kGzErr;
}

__device__ __forceinline__ uint32_t gz_len_base(int s) { return s < 8 ? 3u + s : s == 28 ? 258u : ((4u + (s & 3)) << ((s >> 2) - 1)) + 3u; }
__device__ __forceinline__ uint32_t gz_len_extra(int s) { return (s < 8 || s == 28) ? 0u : (uint32_t)(s >> 2) - 1u; }
__device__ __forceinline__ uint32_t gz_dist_base(int s) { return s < 4 ? 1u + s : ((2u + (s & 1)) << ((s >> 1) - 1)) + 1u; }
__device__ __forceinline__ uint32_t gz_dist_extra(int s) { return s < 4 ? 0u : (uint32_t)(s >> 1) - 1u; }

// literal/length + distance symbols of one block (zlib's LEN .. MATCH / LIT states)
template <bool kWrite>
__device__ __noinline__ int gz_codes(GzState& s, const GzHuff& lc, const GzHuff& dc) {
    for (;;) {
        int sym; int r = gz_decode(s, lc, sym);
        if (r != kGzOk) return r;
        if (sym < 256) { gz_put<kWrite>(s, (uint32_t)sym); if (s.overflow) return kGzErr; continue; }
        if (sym == 256) return kGzOk;
        if (sym >= 286) return kGzErr;                                  // "invalid literal/length code"
        sym -= 257;
        const uint32_t le = gz_len_extra(sym);
        if (!gz_need(s, le)) return kGzTrunc;
        const uint32_t len = gz_len_base(sym) + gz_peek(s, le); gz_drop(s, le);
        int ds; r = gz_decode(s, dc, ds);
        if (r != kGzOk) return r;
        if (ds >= 30) return kGzErr;                                    // "invalid distance code"
        const uint32_t de = gz_dist_extra(ds);
        if (!gz_need(s, de)) return kGzTrunc;
        const uint32_t dist = gz_dist_base(ds) + gz_peek(s, de); gz_drop(s, de);
        // MATCH: a full buffer returns before the distance is checked
        if (s.chunk_fill == kGzChunk) { s.delivered = s.produced; s.chunk_fill = 0; }
        if (dist > s.produced - s.member_start) return kGzErr;          // "invalid distance too far back"
        for (uint32_t k = 0; k < len; k++) {
            gz_put<kWrite>(s, kWrite ? s.out[s.produced - dist] : 0u);
            if (s.overflow) return kGzErr;
        }
    }
}

template <bool kWrite>
__device__ __noinline__ int gz_blocks(GzState& s) {
    GzHuff lc, dc;
    uint8_t lens[320];
    for (;;) {
        if (!gz_need(s, 3)) return kGzTrunc;
        const uint32_t hdr = gz_peek(s, 3); gz_drop(s, 3);
        const uint32_t last = hdr & 1u, type = hdr >> 1;
        if (type == 0) {
            gz_align(s);
            if (!gz_need(s, 32)) return kGzTrunc;
            const uint32_t v = gz_peek(s, 32);
            if ((v & 0xffffu) != ((v >> 16) ^ 0xffffu)) return kGzErr;  // "invalid stored block lengths"
            gz_drop(s, 32);
            for (uint32_t len = v & 0xffffu; len; len--) {
                if (!gz_need(s, 8)) return kGzTrunc;
                gz_put<kWrite>(s, gz_peek(s, 8)); gz_drop(s, 8);
                if (s.overflow) return kGzErr;
            }
        } else if (type == 1) {
            for (int i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            gz_build(lc, lens, 288, 1);
            for (int i = 0; i < 32; i++) lens[i] = 5;
            gz_build(dc, lens, 32, 2);
            const int r = gz_codes<kWrite>(s, lc, dc); if (r != kGzOk) return r;
        } else if (type == 2) {
            if (!gz_need(s, 14)) return kGzTrunc;
            const uint32_t v = gz_peek(s, 14); gz_drop(s, 14);
            const int nlen = (int)(v & 31u) + 257, ndist = (int)((v >> 5) & 31u) + 1, ncode = (int)((v >> 10) & 15u) + 4;
            if (nlen > 286 || ndist > 30) return kGzErr;                // "too many length or distance symbols"
            for (int i = 0; i < 320; i++) lens[i] = 0;
            {
                uint8_t cll[19];
                for (int i = 0; i < 19; i++) cll[i] = 0;
                for (int i = 0; i < ncode; i++) {
                    if (!gz_need(s, 3)) return kGzTrunc;
                    constexpr uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
                    cll[order[i]] = (uint8_t)gz_peek(s, 3); gz_drop(s, 3);
                }
                if (!gz_build(lc, cll, 19, 0)) return kGzErr;            // "invalid code lengths set"   (lc doubles as the code-length table)
            }
            int have = 0;
            while (have < nlen + ndist) {
                int sym;
                if (lc.empty) { if (!gz_need(s, 1)) return kGzTrunc; gz_drop(s, 1); sym = 0; }
                else {
                    // NEEDBITS(here.bits + extra) comes before DROPBITS: when the extra bits are missing nothing was consumed
                    const uint64_t h0 = s.hold; const uint32_t n0 = s.nbits, p0 = s.pos;
                    const int r = gz_decode(s, lc, sym); if (r != kGzOk) return r;
                    if (sym >= 16 && !gz_need(s, sym == 16 ? 2u : sym == 17 ? 3u : 7u)) { s.hold = h0; s.nbits = n0; s.pos = p0; return kGzTrunc; }
                }
                if (sym < 16) { lens[have++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (have == 0) return kGzErr; val = lens[have - 1]; rep = 3 + (int)gz_peek(s, 2); gz_drop(s, 2); }   // "invalid bit length repeat"
                else if (sym == 17) { rep = 3 + (int)gz_peek(s, 3); gz_drop(s, 3); }
                else { rep = 11 + (int)gz_peek(s, 7); gz_drop(s, 7); }
                if (have + rep > nlen + ndist) return kGzErr;
                while (rep--) lens[have++] = (uint8_t)val;
            }
            if (lens[256] == 0) return kGzErr;                          // "invalid code -- missing end-of-block"
            if (!gz_build(lc, lens, nlen, 1)) return kGzErr;             // "invalid literal/lengths set"
            if (!gz_build(dc, lens + nlen, ndist, 2)) return kGzErr;     // "invalid distances set"
            const int r = gz_codes<kWrite>(s, lc, dc); if (r != kGzOk) return r;
        } else return kGzErr;                                           // "invalid block type"
        if (last) return kGzOk;
    }
}

__device__ __forceinline__ uint32_t gz_crc32_byte(uint32_t crc, uint32_t b) {        // IEEE 802.3 (reflected 0xedb88320), a nibble at a time
    constexpr uint32_t t[16] = { 0x00000000u, 0x1db71064u, 0x3b6e20c8u, 0x26d930acu, 0x76dc4190u, 0x6b6b51f4u, 0x4db26158u, 0x5005713cu,
                                 0xedb88320u, 0xf00f9344u, 0xd6d6a3e8u, 0xcb61b38cu, 0x9b64c2b0u, 0x86d3d2d4u, 0xa00ae278u, 0xbdbdf21cu };
    crc ^= b;
    crc = (crc >> 4) ^ t[crc & 15u];
    crc = (crc >> 4) ^ t[crc & 15u];
    return crc;
}
__device__ __noinline__ uint32_t gz_crc32(const uint8_t* p, uint32_t n) {
    uint32_t crc = 0xffffffffu;
    for (uint32_t i = 0; i < n; i++) crc = gz_crc32_byte(crc, p[i]);
    return ~crc;
}
__device__ __noinline__ uint32_t gz_adler32(const uint8_t* p, uint32_t n) {
    uint32_t a = 1, b = 0;
    while (n) {
        uint32_t k = n < 5552u ? n : 5552u; n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521u; b %= 65521u;
    }
    return (b << 16) | a;
}

// one member: header, deflate blocks, trailer (zlib's HEAD .. DONE)
template <bool kWrite>
__device__ __noinline__ int gz_member(GzState& s, int format) {
    s.member_start = s.produced; s.chunk_fill = 0;
    const uint32_t head = gz_byte_pos(s);
    if (format == B2_COMPRESS_TYPE_GZIP) {                              // RFC 1952; only the gzip wrapper is accepted
        if (!gz_need(s, 16)) return kGzTrunc;
        if (gz_peek(s, 16) != 0x8b1fu) return kGzErr;                   // "incorrect header check"
        gz_drop(s, 16);
        if (!gz_need(s, 16)) return kGzTrunc;
        const uint32_t flags = gz_peek(s, 16); gz_drop(s, 16);
        if ((flags & 0xffu) != 8u) return kGzErr;                       // "unknown compression method"
        if (flags & 0xe000u) return kGzErr;                             // "unknown header flags set"
        if (!gz_need(s, 32)) return kGzTrunc;
        gz_drop(s, 32);                                                 // mtime
        if (!gz_need(s, 16)) return kGzTrunc;
        gz_drop(s, 16);                                                 // xfl, os
        if (flags & 0x0400u) {
            if (!gz_need(s, 16)) return kGzTrunc;
            uint32_t xlen = gz_peek(s, 16); gz_drop(s, 16);
            while (xlen) { if (!gz_need(s, 8)) return kGzTrunc; gz_drop(s, 8); xlen--; }
        }
        for (uint32_t f = 0x0800u; f <= 0x1000u; f <<= 1) if (flags & f) {
            for (;;) { if (!gz_need(s, 8)) return kGzTrunc; const uint32_t c = gz_peek(s, 8); gz_drop(s, 8); if (!c) break; }
        }
        if (flags & 0x0200u) {
            if (!gz_need(s, 16)) return kGzTrunc;
            const uint32_t here = gz_byte_pos(s);
            if (gz_peek(s, 16) != (gz_crc32(s.in + head, here - head) & 0xffffu)) return kGzErr;    // "header crc mismatch"
            gz_drop(s, 16);
        }
    } else {                                                            // RFC 1950
        if (!gz_need(s, 16)) return kGzTrunc;
        const uint32_t h = gz_peek(s, 16);
        if ((((h & 0xffu) << 8) + (h >> 8)) % 31u) return kGzErr;       // "incorrect header check"
        if ((h & 0xfu) != 8u) return kGzErr;                            // "unknown compression method"
        if (((h >> 4) & 0xfu) + 8u > 15u) return kGzErr;                // "invalid window size"
        gz_drop(s, 16);
        if (h & 0x2000u) { if (!gz_need(s, 32)) return kGzTrunc; return kGzErr; }        // FDICT: Z_NEED_DICT ends the stream
    }
    const int r = gz_blocks<kWrite>(s);
    if (r != kGzOk) return r;
    gz_align(s);
    if (!gz_need(s, 32)) return kGzTrunc;
    const uint32_t t = gz_peek(s, 32);
    const uint32_t mn = s.produced - s.member_start;
    if (format == B2_COMPRESS_TYPE_GZIP) {
        if (kWrite && t != gz_crc32(s.out + s.member_start, mn)) return kGzErr;          // "incorrect data check"
        gz_drop(s, 32);
        if (!gz_need(s, 32)) return kGzTrunc;
        if (gz_peek(s, 32) != mn) return kGzErr;                        // "incorrect length check"
        gz_drop(s, 32);
    } else {
        if (kWrite && __byte_perm(t, 0, 0x0123) != gz_adler32(s.out + s.member_start, mn)) return kGzErr;
        gz_drop(s, 32);
    }
    return kGzEnd;
}

// The bytes GzipInputStream(format) yields for in[0, n): written to out[0, cap) when kWrite; returns how many.  *too_big = the stream
// produces more than cap bytes (sizing pass: more than kGzMaxOut) — the caller leaves such a body to the host.
template <bool kWrite>
__device__ __noinline__ uint32_t gz_input_stream(const uint8_t* in, uint32_t n, int format, uint8_t* out, uint32_t cap, bool* too_big) {
    GzState s; s.in = in; s.n = n; s.pos = 0; s.hold = 0; s.nbits = 0; s.out = out; s.cap = cap;
    s.produced = 0; s.delivered = 0; s.chunk_fill = 0; s.member_start = 0; s.overflow = false;
    for (;;) {
        if (s.pos >= s.n && s.nbits == 0) { s.delivered = s.produced; break; }        // the sub-stream is at its end when the next member would start
        const int r = gz_member<kWrite>(s, format);
        if (s.overflow) break;
        if (r == kGzErr) break;                                         // the erroring call's chunk is never handed over
        s.delivered = s.produced;
        if (r == kGzTrunc) break;
    }
    *too_big = s.overflow;
    return kWrite ? s.delivered : s.produced;
}

}  // namespace b2
