// ref_leaf_shim.cc — extern "C" doorway into the UNMODIFIED reference leaf code
// (src/butil/crc32c.cc and src/butil/third_party/snappy/*.cc), compiled from
// /root/reference where it lies by oracle/Makefile into oracle/_ref/libref_leaf.so.
// Test infrastructure only; nothing from the reference is copied into this repo.
#include <stddef.h>
#include <stdint.h>
#include <string>
#include "butil/crc32c.h"
#include "butil/third_party/snappy/snappy.h"

extern "C" {
uint32_t ref_crc32c_extend(uint32_t init, const char* data, size_t n) {
    return butil::crc32c::Extend(init, data, n);          // crc32c.cc:451-454
}
uint32_t ref_crc32c_mask(uint32_t crc) { return butil::crc32c::Mask(crc); }
uint32_t ref_crc32c_unmask(uint32_t m) { return butil::crc32c::Unmask(m); }
size_t ref_snappy_max_compressed_length(size_t n) {
    return butil::snappy::MaxCompressedLength(n);         // snappy.cc:55-77
}
int ref_snappy_compress(const char* in, size_t n, char* out, size_t* out_len) {
    butil::snappy::RawCompress(in, n, out, out_len);      // snappy.cc:875-956 via ByteArraySource
    return 1;
}
int ref_snappy_uncompressed_length(const char* in, size_t n, size_t* result) {
    return butil::snappy::GetUncompressedLength(in, n, result) ? 1 : 0;
}
int ref_snappy_uncompress(const char* in, size_t n, char* out, size_t cap, size_t* got) {
    size_t ulen = 0;
    if (!butil::snappy::GetUncompressedLength(in, n, &ulen)) return 0;
    if (ulen > cap) return 0;
    if (!butil::snappy::RawUncompress(in, n, out)) return 0;   // snappy.cc:1526-1552 semantics
    *got = ulen;
    return 1;
}
}
