// Test harness (not product): warp-cooperative device functions of b2_kernels.cuh on an emulated 32-lane warp (warp_emul_prelude.h).
#include "warp_emul_prelude.h"
#include "kernels_host.cuh"
#include <string>
#include <vector>
using namespace b2;
namespace b2 { __attribute__((aligned(128))) uint8_t fused_raw[16], pack_smem_raw[16], small_raw[16]; __attribute__((aligned(16))) uint8_t s_rings[16]; uint32_t sm[4]; }

extern "C" {
// butil::crc32c::Extend through warp_crc32c_update (all 32 lanes must agree): tables = the host-built crc_adv blob the library uploads
uint32_t wh_crc32c(const uint32_t* crc_adv, uint32_t init_crc, const uint8_t* p, uint32_t n, uint32_t* lanes_agree) {
    uint32_t res[32];
    CrcTabs ct; ct.hot = crc_adv; ct.tree = crc_adv + kCrcHotWords; ct.ring = nullptr;
    we_run_warp([&](unsigned lane) { res[lane] = warp_crc32c_update(init_crc ^ 0xffffffffu, p, n, lane, ct) ^ 0xffffffffu; });
    *lanes_agree = 1; for (int l = 1; l < 32; l++) if (res[l] != res[0]) *lanes_agree = 0;
    return res[0];
}
}

// ---- tables exactly as b2_ctx_create builds them (brpc_b200/csrc/b2_api.cu: crc_table_init + the warp-CRC operators) -------------------------
static std::vector<uint32_t> g_adv;
static void wh_tables() {
    if (!g_adv.empty()) return;
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82f63b78u & (0u - (c & 1u))); c_crc_table[i] = c; }
    g_adv.resize(kCrcHotWords + kCrcTreeWords);
    auto adv = [&](uint32_t x, int bytes) { for (int k = 0; k < bytes; k++) x = c_crc_table[x & 0xff] ^ (x >> 8); return x; };
    for (int k = 0; k < 16; k++) for (uint32_t b = 0; b < 256; b++) g_adv[k * 256 + b] = adv(c_crc_table[b], k);
    for (int j = 0; j < 4; j++) for (uint32_t b = 0; b < 256; b++) g_adv[(16 + j) * 256 + b] = adv(b << (8 * j), 512);
    for (int t = 0; t < 5; t++) for (int j = 0; j < 4; j++) for (uint32_t b = 0; b < 256; b++) g_adv[kCrcHotWords + (t * 4 + j) * 256 + b] = adv(b << (8 * j), 16 << t);
}
extern "C" {
uint32_t wh_crc32c_extend(uint32_t init_crc, const uint8_t* p, uint32_t n, uint32_t* lanes_agree) { wh_tables(); return wh_crc32c(g_adv.data(), init_crc, p, n, lanes_agree); }
// butil::snappy::RawUncompress / RawCompress through the warp primitives
int wh_snappy_uncompress(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* produced) {
    int ok[32]; uint32_t prod[32];
    we_run_warp([&](unsigned lane) { uint32_t p = 0; ok[lane] = warp_snappy_decode(in, n, out, cap, lane, p, nullptr) ? 1 : 0; prod[lane] = p; });
    *produced = prod[0];
    for (int l = 1; l < 32; l++) if (ok[l] != ok[0] || prod[l] != prod[0]) return -1;
    return ok[0];
}
uint32_t wh_snappy_compress(const uint8_t* in, uint32_t n, uint8_t* out) {
    std::vector<uint16_t> table(kSnappyMaxTable);
    uint32_t len[32];
    we_run_warp([&](unsigned lane) { len[lane] = warp_snappy_compress(in, n, out, table.data(), lane); });
    for (int l = 1; l < 32; l++) if (len[l] != len[0]) return 0xffffffffu;
    return len[0];
}

// ---- a batch through decode_one (lane 0, one message at a time) -> slot scan -> pack_one on the emulated warp: what k_small / k_decode +
// k_pack_slow do with it, for EVERY kind of message (the bandwidth kernels only move bytes pack_one would place the same way)
struct wh_ctx { DevMethod methods[4]; uint32_t n_methods; DevConfig C; };
wh_ctx* wh_create(uint64_t max_body, uint32_t proto_mask, uint32_t stream_handler, const char* identity) {
    wh_tables();
    wh_ctx* k = new wh_ctx; memset(k, 0, sizeof *k);
    k->C.max_body_size = max_body ? max_body : (64ull << 20); k->C.proto_mask = proto_mask; k->C.stream_handler = stream_handler;
    k->C.tile_bytes = 8192; k->C.tile_shift = 13; k->C.spec_k = 16; k->C.pull_vecs = 8;
    if (identity) { k->C.identity_len = (uint32_t)strlen(identity); memcpy(k->C.identity, identity, k->C.identity_len); }
    return k;
}
void wh_destroy(wh_ctx* k) { delete k; }
void wh_add_method(wh_ctx* k, const char* service_full, const char* service_short, const char* method, const char* request_type, int handler, int echo_att, int r_cks, int r_cmp) {
    DevMethod& m = k->methods[k->n_methods++];
    const std::string full = std::string(service_full) + "." + method;
    m.full_method_len = (uint32_t)full.size(); memcpy(m.full_method, full.data(), full.size());
    m.service_short_len = (uint32_t)strlen(service_short); memcpy(m.service_short, service_short, m.service_short_len);
    m.service_full_len = (uint32_t)strlen(service_full); memcpy(m.service_full, service_full, m.service_full_len);
    m.request_type_len = (uint32_t)strlen(request_type); memcpy(m.request_type, request_type, m.request_type_len);
    m.handler = handler; m.echo_attachment = echo_att; m.response_checksum_type = r_cks; m.response_compress_type = r_cmp;
    k->C.n_methods = k->n_methods;
}
// fo_raw[i] / run_of[i]: the frame table (from the oracle's cut: the front kernels are tested elsewhere).  msgs / resp as the ABI returns them.
int wh_process(wh_ctx* k, const uint8_t* bytes, const b2_run* runs, uint32_t n_runs, const uint32_t* fo_raw, const uint32_t* run_of, uint32_t n,
               b2_msg_desc* msgs, uint8_t* resp, uint32_t resp_cap, uint32_t* resp_used) {
    std::vector<MsgAux> aux(n + 1); std::vector<PackJob> jobs(n + 1); std::vector<uint32_t> slot(n + 2, 0), scan_tmp(64, 0), totals(16, 0), frame_run(run_of, run_of + n);
    std::vector<uint4> refs(n + 1); std::vector<uint8_t> heads((size_t)(n + 1) * kHeadBytes), unz(2 * (size_t)resp_cap + 64); std::vector<uint16_t> tab(kSnappyMaxTable);
    std::vector<uint32_t> frame_off(fo_raw, fo_raw + n);
    BatchPtrs B; memset(&B, 0, sizeof B);
    B.bytes = bytes; B.runs = runs; B.n_runs = n_runs; B.frame_off = frame_off.data(); B.frame_run = frame_run.data(); B.msgs = msgs; B.aux = aux.data(); B.jobs = jobs.data();
    B.slot = slot.data(); B.scan_tmp = scan_tmp.data(); B.refs = refs.data(); B.heads = heads.data(); B.resp = resp; B.unz = unz.data(); B.snappy_tab = tab.data();
    B.methods = k->methods; B.totals = totals.data(); B.crc_adv = g_adv.data(); B.max_msgs = n; B.max_resp = resp_cap;
    threadIdx.x = 0; blockIdx.x = 0; blockDim.x = 32; gridDim.x = 1;
    for (uint32_t i = 0; i < n; i++) decode_one<false>(B, k->C, i, fo_raw[i], bytes + (fo_raw[i] & 0x7fffffffu), heads.data() + (size_t)i * kHeadBytes, 0xffffffffu);
    uint32_t off = 0;
    for (uint32_t i = 0; i < n; i++) { const uint32_t l = slot[i]; slot[i] = off; off += l; }          // k_scan_blocks
    if (off > resp_cap) return -2;
    *resp_used = off; totals[0] = n; totals[1] = off;
    CrcTabs ct; ct.hot = g_adv.data(); ct.tree = g_adv.data() + kCrcHotWords; ct.ring = nullptr;
    const DevConfig C = k->C;
    we_run_warp([&](unsigned lane) { for (uint32_t i = 0; i < n; i++) { pack_one(B, C, i, lane, ct); __syncwarp(); } });
    return 0;
}
}
