// Test harness (not product): the per-message device functions of brpc_b200/csrc/b2_kernels.cuh — decode_one (k_decode / k_small's decoder) and
// fused_fast_echo (k_fused's exact-shape decoder) — called from the CPU suite on single messages, through the generated host-compilable copy of
// the kernels file (tests/cpp/gen_kernels_host.py).  One lane, one message: what a lane of the kernels does.
#include "kernels_host_prelude.h"
#include "kernels_host.cuh"
#include <string>
#include <vector>
using namespace b2;
// the kernels' dynamic shared-memory arrays (declared `extern` in the generated copy): never touched here, the linker just wants them
namespace b2 { __attribute__((aligned(128))) uint8_t fused_raw[16], pack_smem_raw[16], small_raw[16]; __attribute__((aligned(16))) uint8_t s_rings[16]; uint32_t sm[4]; }

extern "C" {
struct kh_out {
    b2_msg_desc d; MsgAux a; PackJob job; uint32_t slot; uint32_t ref[4]; uint8_t head[kHeadBytes];
    uint32_t fast_ok, fast_prefix, fast_rs;          // fused_fast_echo: accepted?, prefix length, reply start (batch offset)
    b2_msg_desc fast_d;
};
struct kh_ctx { DevMethod methods[4]; uint32_t n_methods; DevConfig C; };
kh_ctx* kh_create(uint64_t max_body, uint32_t proto_mask, uint32_t by_ref, uint32_t stream_handler, const char* identity) {
    kh_ctx* k = new kh_ctx; memset(k, 0, sizeof *k);
    k->C.max_body_size = max_body ? max_body : (64ull << 20); k->C.proto_mask = proto_mask; k->C.by_ref = by_ref; k->C.stream_handler = stream_handler;
    k->C.tile_bytes = 8192; k->C.tile_shift = 13; k->C.spec_k = 16; k->C.pull_vecs = 8;
    if (identity) { k->C.identity_len = (uint32_t)strlen(identity); memcpy(k->C.identity, identity, k->C.identity_len); }
    return k;
}
void kh_destroy(kh_ctx* k) { delete k; }
// what b2_register_method writes into the device table
void kh_add_method(kh_ctx* k, const char* service_full, const char* service_short, const char* method, const char* request_type, int handler, int echo_att, int r_cks, int r_cmp) {
    DevMethod& m = k->methods[k->n_methods++];
    const std::string full = std::string(service_full) + "." + method;
    m.full_method_len = (uint32_t)full.size(); memcpy(m.full_method, full.data(), full.size());
    m.service_short_len = (uint32_t)strlen(service_short); memcpy(m.service_short, service_short, m.service_short_len);
    m.service_full_len = (uint32_t)strlen(service_full); memcpy(m.service_full, service_full, m.service_full_len);
    m.request_type_len = (uint32_t)strlen(request_type); memcpy(m.request_type, request_type, m.request_type_len);
    m.handler = handler; m.echo_attachment = echo_att; m.response_checksum_type = r_cks; m.response_compress_type = r_cmp;
    k->C.n_methods = k->n_methods;
}
// one message: bytes = the whole batch buffer (padded by the caller), fo_raw = its frame offset (| bit 31 when not baidu_std), run = its run
void kh_decode(kh_ctx* k, const uint8_t* bytes, uint32_t fo_raw, const b2_run* run, kh_out* out) {
    memset(out, 0, sizeof *out);
    uint32_t frame_run = 0, totals[16] = { 0 }; uint32_t ref4[4] = { 0, 0, 0, 0 };
    BatchPtrs B; memset(&B, 0, sizeof B);
    B.bytes = bytes; B.runs = run; B.n_runs = 1; B.frame_run = &frame_run; B.msgs = &out->d; B.aux = &out->a; B.jobs = &out->job; B.slot = &out->slot;
    B.refs = reinterpret_cast<uint4*>(ref4); B.heads = out->head; B.methods = k->methods; B.totals = totals; B.max_msgs = 1; B.max_resp = 0xfffffff0u;
    const uint32_t fo = fo_raw & 0x7fffffffu;
    // (srow: the frame's first bytes as staged by decode_round — here simply the frame itself, all of it "staged")
    decode_one<false>(B, k->C, 0, fo_raw, bytes + fo, out->head, 0xffffffffu);
    memcpy(out->ref, ref4, 16);
    // the exact-shape decoder of k_fused on a private copy of the frame (it writes the reply prefix in place)
    const uint32_t body = load_be32(bytes + fo + 4);
    if (!(fo_raw >> 31) && !(run->flags & (B2_RUN_CLIENT | B2_RUN_RPC_DUMP)) && load_le32(bytes + fo) == kMagicPRPC && body < (64u << 20)) {
        std::vector<uint8_t> img(bytes + fo, bytes + fo + 12 + body + 16);
        b2_msg_desc fd; memset(&fd, 0, sizeof fd);
        BatchPtrs F = B; F.msgs = &fd;
        DecodeOut o; o.fast = false; o.slow = false; o.prefix = 0; o.rs = 0;
        if (fused_fast_echo(F, k->methods, k->n_methods, 0, fo, 0, img.data(), 12 + body, nullptr, o)) {
            out->fast_ok = 1; out->fast_prefix = o.prefix; out->fast_rs = o.rs; out->fast_d = fd;
            // hand the patched prefix back through `head` is not possible (in place): the caller asks for it separately
        }
    }
}
// the reply fused_fast_echo builds in place, for a message it accepts: returns its length (0 = declined), bytes to `reply`
uint32_t kh_fast_reply(kh_ctx* k, const uint8_t* bytes, uint32_t fo, uint8_t* reply, uint32_t cap) {
    const uint32_t body = load_be32(bytes + fo + 4);
    if (load_le32(bytes + fo) != kMagicPRPC || body >= (64u << 20)) return 0;
    std::vector<uint8_t> img(bytes + fo, bytes + fo + 12 + body + 16);
    b2_msg_desc fd; BatchPtrs F; memset(&F, 0, sizeof F); F.msgs = &fd; F.max_msgs = 1;
    DecodeOut o; o.fast = false; o.slow = false; o.prefix = 0; o.rs = 0;
    if (!fused_fast_echo(F, k->methods, k->n_methods, 0, fo, 0, img.data(), 12 + body, nullptr, o)) return 0;
    if (fd.resp_len > cap) return 0;
    memcpy(reply, img.data() + (fd.resp_off - fo), fd.resp_len);
    return fd.resp_len;
}
}
extern "C" uint32_t kh_sizeof_out(void) { return (uint32_t)sizeof(kh_out); }
// the error reply lane 0 of the pack stage writes for a message decode_one answered with an error: returns its length
extern "C" uint32_t kh_error_reply(kh_ctx* k, const uint8_t* bytes, const kh_out* o, uint8_t* reply) {
    return pack_error_reply(reply, k->C, k->methods, o->d, o->a, bytes + (o->d.frame_off & 0x7fffffffu));
}
// k_tile_walk's speculative walk (register fast path, header prefetch) and the plain restatement it must equal, from the same entry:
// out = {exit, count, kind, last_proto} x 2, offs = the offsets each of them emitted
struct EmitVec { uint32_t* out; uint32_t cap; uint32_t* n; void operator()(uint32_t i, const Step& s) const { if (i < cap) out[i] = s.frame_pos | ((uint32_t)(s.index != 1) << 31); *n = i + 1; } };
extern "C" void kh_walks(const uint8_t* run, uint32_t len, uint32_t entry, uint32_t tile_end, uint64_t max_body, int client, uint32_t mask,
                         uint32_t* out, uint32_t* offs_spec, uint32_t* offs_plain, uint32_t cap) {
    TileRec a, b; memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    uint32_t na = 0, nb = 0;
    EmitVec ea = { offs_spec, cap, &na }, eb = { offs_plain, cap, &nb };
    walk_tile_spec(run, len, entry, tile_end, max_body, client != 0, a, ea, mask);
    walk_tile<true>(run, len, entry, -1, tile_end, max_body, client != 0, b, eb, mask);
    out[0] = a.exit; out[1] = a.count; out[2] = a.kind; out[3] = (uint32_t)(int)a.last_proto;
    out[4] = b.exit; out[5] = b.count; out[6] = b.kind; out[7] = (uint32_t)(int)b.last_proto;
}

// ---- the front of the staged pipeline on the host: k_tile_walk (thread per tile) -> k_resolve (a block of ONE thread per run: its block-wide
// steps degenerate correctly, the cross-run prefix is done here instead of by its last-CTA epilogue) -> k_frame_table.  The speculative entries
// are GIVEN (k_tile_search is warp-cooperative and not what is under test): k_resolve must produce the sequential cut whatever they are.
extern "C" int kh_front(const uint8_t* bytes, const b2_run* runs, uint32_t n_runs, uint32_t tile_shift, uint32_t proto_mask, uint64_t max_body,
                        const uint32_t* entries /* per tile, run-relative, kNone = none */, uint32_t n_tiles_expected,
                        b2_run_status* rs_out, uint32_t* frame_off, uint32_t* frame_run, uint32_t cap, uint32_t* n_msgs_out, uint32_t* n_rewalked) {
    const uint32_t tile = 1u << tile_shift;
    std::vector<uint32_t> rtb(n_runs + 1, 0);
    for (uint32_t r = 0; r < n_runs; r++) rtb[r + 1] = rtb[r] + (uint32_t)(((uint64_t)runs[r].length + tile - 1) >> tile_shift);
    const uint32_t nt = rtb[n_runs];
    if (nt != n_tiles_expected) return -1;
    std::vector<uint4> ti(nt ? nt : 1);
    for (uint32_t r = 0; r < n_runs; r++) for (uint32_t t = rtb[r]; t < rtb[r + 1]; t++) ti[t] = make_uint4(runs[r].offset, runs[r].length, t - rtb[r], r | (runs[r].flags << 24));
    std::vector<TileRec> tiles(nt ? nt : 1); std::vector<uint32_t> tile_base(nt + 1, 0), scratch(3 * (size_t)nt + 3, 0), spec((size_t)kSpecK * nt + 1, 0), totals(16, 0);
    for (uint32_t t = 0; t < nt; t++) { memset(&tiles[t], 0, sizeof(TileRec)); tiles[t].entry = entries[t]; }
    BatchPtrs B; memset(&B, 0, sizeof B);
    B.bytes = bytes; B.runs = runs; B.run_tile_base = rtb.data(); B.tile_info = ti.data(); B.tiles = tiles.data(); B.tile_base = tile_base.data();
    B.tile_scratch = scratch.data(); B.tile_spec = spec.data(); B.run_status = rs_out; B.frame_off = frame_off; B.frame_run = frame_run; B.totals = totals.data();
    B.n_runs = n_runs; B.n_tiles = nt; B.max_msgs = cap; B.max_resp = 0xfffffff0u;
    DevConfig C; memset(&C, 0, sizeof C);
    C.max_body_size = max_body ? max_body : (64ull << 20); C.tile_bytes = tile; C.tile_shift = tile_shift; C.spec_k = kSpecK; C.proto_mask = proto_mask;
    blockDim.x = 1; threadIdx.x = 0; gridDim.x = 0x7fffffffu;                   // (the last-CTA epilogues never fire: the prefix over runs is done below)
    for (uint32_t t = 0; t < nt; t++) { blockIdx.x = t; k_tile_walk(B, C); }
    for (uint32_t r = 0; r < n_runs; r++) { blockIdx.x = r; k_resolve(B, C, 1u); }
    uint32_t total = 0, rew = 0;
    for (uint32_t r = 0; r < n_runs; r++) { rs_out[r].first_msg = total; total += rs_out[r].n_msgs; }
    for (uint32_t t = 0; t < nt; t++) rew += (tiles[t].kind & kKindRewalked) ? 1u : 0u;
    totals[0] = total;
    *n_msgs_out = total; *n_rewalked = rew;
    if (total > cap) return -2;
    for (uint32_t g = 0; g < nt * kSpecK; g++) { blockIdx.x = g; k_frame_table(B, C); }
    return 0;
}
