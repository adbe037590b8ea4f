// Test harness (not product): k_small — the whole latency path in one CTA of 512 threads (cut loop per run, frame table, decode rounds with their
// shared-memory row staging, slot scan, pack) — on an emulated thread block (block_emul_prelude.h), out of the generated host-compilable copy of
// b2_kernels.cuh.  The same code k_ring runs per batch.
#include "block_emul_prelude.h"
#include "kernels_host.cuh"
#include <string>
using namespace b2;
namespace b2 { __attribute__((aligned(128))) uint8_t fused_raw[16], pack_smem_raw[16], small_raw[sizeof(SmallSmem) + 128]; __attribute__((aligned(16))) uint8_t s_rings[16]; uint32_t sm[4]; }

static std::vector<uint32_t> g_adv;
static void sh_tables() {
    if (!g_adv.empty()) return;
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0x82f63b78u & (0u - (c & 1u))); c_crc_table[i] = c; }
    g_adv.resize(kCrcHotWords + kCrcTreeWords);
    auto adv = [&](uint32_t x, int bytes) { for (int k = 0; k < bytes; k++) x = c_crc_table[x & 0xff] ^ (x >> 8); return x; };
    for (int k = 0; k < 16; k++) for (uint32_t b = 0; b < 256; b++) g_adv[k * 256 + b] = adv(c_crc_table[b], k);
    for (int j = 0; j < 4; j++) for (uint32_t b = 0; b < 256; b++) g_adv[(16 + j) * 256 + b] = adv(b << (8 * j), 512);
    for (int t = 0; t < 5; t++) for (int j = 0; j < 4; j++) for (uint32_t b = 0; b < 256; b++) g_adv[kCrcHotWords + (t * 4 + j) * 256 + b] = adv(b << (8 * j), 16 << t);
}
extern "C" {
struct sh_ctx { DevMethod methods[4]; uint32_t n_methods; DevConfig C; };
sh_ctx* sh_create(uint64_t max_body, uint32_t proto_mask, uint32_t by_ref, uint32_t stream_handler, const char* identity) {
    sh_tables();
    sh_ctx* k = new sh_ctx; memset(k, 0, sizeof *k);
    k->C.max_body_size = max_body ? max_body : (64ull << 20); k->C.proto_mask = proto_mask; k->C.by_ref = by_ref; k->C.stream_handler = stream_handler;
    k->C.tile_bytes = 8192; k->C.tile_shift = 13; k->C.spec_k = 16; k->C.pull_vecs = 8;
    if (identity) { k->C.identity_len = (uint32_t)strlen(identity); memcpy(k->C.identity, identity, k->C.identity_len); }
    return k;
}
void sh_destroy(sh_ctx* k) { delete k; }
void sh_add_method(sh_ctx* k, const char* service_full, const char* service_short, const char* method, const char* request_type, int handler, int echo_att, int r_cks, int r_cmp) {
    DevMethod& m = k->methods[k->n_methods++];
    const std::string full = std::string(service_full) + "." + method;
    m.full_method_len = (uint32_t)full.size(); memcpy(m.full_method, full.data(), full.size());
    m.service_short_len = (uint32_t)strlen(service_short); memcpy(m.service_short, service_short, m.service_short_len);
    m.service_full_len = (uint32_t)strlen(service_full); memcpy(m.service_full, service_full, m.service_full_len);
    m.request_type_len = (uint32_t)strlen(request_type); memcpy(m.request_type, request_type, m.request_type_len);
    m.handler = handler; m.echo_attachment = echo_att; m.response_checksum_type = r_cks; m.response_compress_type = r_cmp;
    k->C.n_methods = k->n_methods;
}
// one batch through k_small: the buffers b2_batch_upload / make_ptrs give it (capacities max_msgs / max_resp), outputs as the ABI returns them.
// returns the kernel's overflow flags (totals[2]); n_msgs / resp_bytes from totals[0] / [1]
uint32_t sh_k_small(sh_ctx* k, const uint8_t* bytes, const b2_run* runs, uint32_t n_runs, uint32_t max_msgs, uint32_t max_resp,
                    b2_run_status* rs, b2_msg_desc* msgs, uint32_t* refs4, uint8_t* resp, uint32_t* n_msgs, uint32_t* resp_bytes) {
    std::vector<MsgAux> aux(max_msgs + 1); std::vector<PackJob> jobs(max_msgs + 1); std::vector<uint32_t> slot(max_msgs + 2, 0), scan_tmp(64, 0), totals(16, 0);
    std::vector<uint32_t> frame_off(max_msgs + 1), frame_run(max_msgs + 1), frame_row(max_msgs + 1), slow_idx(max_msgs + 1);
    std::vector<uint8_t> heads((size_t)(max_msgs + 1) * kHeadBytes), unz(2 * (size_t)max_resp + 64); std::vector<uint16_t> tab((size_t)(kSmallThreads / 32) * kSnappyMaxTable);   // one hash table per warp (b2_ctx_create: kSnappyWarps of them)
    std::vector<unsigned long long> counters(16, 0);
    BatchPtrs B; memset(&B, 0, sizeof B);
    B.bytes = bytes; B.runs = runs; B.n_runs = n_runs; B.run_status = rs; B.frame_off = frame_off.data(); B.frame_run = frame_run.data(); B.frame_row = frame_row.data();
    B.msgs = msgs; B.aux = aux.data(); B.jobs = jobs.data(); B.refs = reinterpret_cast<uint4*>(refs4); B.slow_idx = slow_idx.data(); B.heads = heads.data(); B.slot = slot.data();
    B.scan_tmp = scan_tmp.data(); B.resp = resp; B.unz = unz.data(); B.snappy_tab = tab.data(); B.counters = counters.data(); B.totals = totals.data();
    B.methods = k->methods; B.crc_adv = g_adv.data(); B.max_msgs = max_msgs; B.max_resp = max_resp;
    const DevConfig C = k->C;
    be_run_block(kSmallThreads, [&]() { k_small(B, C); });
    *n_msgs = totals[0]; *resp_bytes = totals[1];
    return totals[2];
}
}
