// C++ tests of the host side (brpc_b200/host).  `host_test cpu` needs no GPU: b2::IOBuf behaves like
// the slice of butil::IOBuf the path uses (cases modelled on test/iobuf_unittest.cpp: block geometry
// :60-61, append/cut/pop, zero-copy sharing, user data :1593-1723, the counting allocator installed
// through blockmem_allocate/deallocate :40-100).  `host_test gpu` drives GpuInputMessenger end to end
// and checks every response byte against the oracle (tests may link the oracle).
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../brpc_b200/host/input_messenger.h"
#include "../../brpc_b200/host/h2_messenger.h"
#include "../../brpc_b200/host/protocol.h"
#include "../../oracle/b2_oracle.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

extern "C" {
typedef struct b2press_spec { const char* service; const char* method; uint32_t payload_bytes, attachment_bytes; int32_t payload_kind, checksum_type; uint64_t seed; } b2press_spec;
size_t b2press_frame(const b2press_spec* s, uint64_t index, uint8_t* out, size_t cap);
}

static int g_live_blocks = 0, g_total_allocs = 0;
static void* counting_alloc(size_t n) { g_live_blocks++; g_total_allocs++; return malloc(n); }
static void counting_free(void* p) { g_live_blocks--; free(p); }

static void test_iobuf() {
    using b2::IOBuf;
    b2::iobuf::blockmem_allocate = counting_alloc; b2::iobuf::blockmem_deallocate = counting_free;
    {
        IOBuf b;
        CHECK(b.empty() && b.length() == 0 && b.backing_block_num() == 0);
        std::string s(20000, 'x'); for (size_t i = 0; i < s.size(); i++) s[i] = (char)('a' + i % 26);
        CHECK(b.append(s) == 0);
        CHECK(b.length() == 20000 && b.backing_block_num() == 3);               // 8160 + 8160 + 3680
        CHECK(b.backing_block(0).second == IOBuf::DEFAULT_PAYLOAD && IOBuf::DEFAULT_PAYLOAD == 8192 - 32);
        CHECK(b.backing_block(2).second == 20000 - 2 * 8160 && b.backing_block(3).first == nullptr);
        CHECK(b.to_string() == s);
        char hdr[12]; CHECK(b.copy_to(hdr, 12) == 12 && memcmp(hdr, s.data(), 12) == 0);
        CHECK(b.copy_to(hdr, 12, 8155) == 12 && memcmp(hdr, s.data() + 8155, 12) == 0);   // spans two blocks
        CHECK(b.copy_to(hdr, 12, 19995) == 5 && b.copy_to(hdr, 1, 20000) == 0);
        char aux[16]; CHECK(b.fetch(aux, 8) == b.backing_block(0).first);          // contiguous: no copy
        IOBuf head; CHECK(b.cutn(&head, 8150) == 8150);
        CHECK(b.fetch(aux, 16) == aux && memcmp(aux, s.data() + 8150, 16) == 0);   // straddles: copied
        IOBuf meta, payload;                                                        // ParseRpcMessage's two cutn
        CHECK(b.cutn(&meta, 46) == 46 && b.cutn(&payload, 1027) == 1027);
        CHECK(meta.to_string() == s.substr(8150, 46) && payload.to_string() == s.substr(8196, 1027));
        CHECK(meta.backing_block_num() == 2 && payload.backing_block_num() == 1);
        CHECK(payload.block_nshared(0) >= 3);                                       // shared, not copied
        CHECK(b.length() == 20000 - 8150 - 46 - 1027);
        CHECK(b.pop_front(100) == 100 && b.pop_back(50) == 50 && b.to_string() == s.substr(8150 + 46 + 1027 + 100, b.length()));
        CHECK(b.pop_front(1 << 30) == 20000 - 8150 - 46 - 1027 - 150 && b.empty());
        IOBuf all; all.append(head); all.append(meta.movable()); all.append(payload);
        CHECK(meta.empty() && all.length() == 8150 + 46 + 1027 && all.to_string() == s.substr(0, all.length()));
        CHECK(all.backing_block_num() == 2);                                        // adjacent refs of one block merge
        IOBuf copy(all); std::string t; CHECK(copy.cutn(&t, 5) == 5 && t == s.substr(0, 5) && all.length() == 9223);
        char c; CHECK(copy.cut1(&c) && c == s[5]);
        IOBuf mv(all.movable()); CHECK(all.empty() && mv.length() == 9223);
        swap(all, mv); CHECK(mv.empty() && all.length() == 9223);
        // user data: referenced, not copied; deleter runs when the last reference goes away
        static int deleted = 0; static char user[64] = "user-owned-bytes";
        { IOBuf u; CHECK(u.append_user_data(user, 16, [](void* p) { deleted += (p == user); }) == 0);
          IOBuf u2(u); CHECK(u.backing_block(0).first == user && u2.to_string() == "user-owned-bytes");
          u.clear(); CHECK(deleted == 0); }
        CHECK(deleted == 1);
        // appending small pieces fills the tail block instead of allocating
        IOBuf small; const int before = g_total_allocs;
        for (int i = 0; i < 1000; i++) small.append("12345678", 8);
        CHECK(small.length() == 8000 && g_total_allocs == before + 1 && small.backing_block_num() == 1);
    }
    CHECK(g_live_blocks == 0);                                                      // every block went back (iobuf_unittest.cpp:94-100)
    b2::iobuf::blockmem_allocate = b2::iobuf::default_alloc; b2::iobuf::blockmem_deallocate = b2::iobuf::default_free;
    printf("iobuf ok (%d blocks allocated and freed)\n", g_total_allocs);
}

// IOPortal / writev cuts / byte iterator over a socketpair (test/iobuf_unittest.cpp cut_into_fd / append_from_fd cases :462-520, :1400-1460)
#include <errno.h>
#include <fcntl.h>
#include <sys/socket.h>
static void test_portal_and_fd() {
    using b2::IOBuf; using b2::IOPortal;
    b2::iobuf::blockmem_allocate = counting_alloc; b2::iobuf::blockmem_deallocate = counting_free;
    const int live_before = g_live_blocks;
    {
        int sv[2]; CHECK(socketpair(AF_UNIX, SOCK_STREAM, 0, sv) == 0);
        CHECK(fcntl(sv[1], F_SETFL, fcntl(sv[1], F_GETFL) | O_NONBLOCK) == 0);
        CHECK(fcntl(sv[0], F_SETFL, fcntl(sv[0], F_GETFL) | O_NONBLOCK) == 0);     // a full socket buffer must not hang the test
        std::string s(50000, 0); for (size_t i = 0; i < s.size(); i++) s[i] = (char)(i * 131 + 7);
        IOBuf out; out.append(s.substr(0, 20000)); IOBuf out2; out2.append(s.substr(20000, 17000)); IOBuf out3; out3.append(s.substr(37000));
        // three queued replies in one writev, then the rest piece by piece
        IOBuf* pieces[3] = { &out, &out2, &out3 };
        size_t sent = 0; IOPortal in; std::string got;
        while (sent < s.size() || in.length() + got.size() < s.size()) {
            if (sent < s.size()) { const ssize_t nw = IOBuf::cut_multiple_into_file_descriptor(sv[0], pieces, 3); if (nw > 0) sent += (size_t)nw; else CHECK(errno == EAGAIN || errno == EWOULDBLOCK); }
            for (;;) {                                               // Socket::DoRead until EAGAIN, like OnNewMessages
                const ssize_t nr = in.append_from_file_descriptor(sv[1], 12345);
                if (nr < 0) { CHECK(errno == EAGAIN || errno == EWOULDBLOCK); break; }
                CHECK(nr > 0 && nr <= 12345);
            }
            if (in.length() > 30000) { std::string part; in.cutn(&part, 11111); got += part; }   // the messenger pops consumed bytes in between
        }
        CHECK(out.empty() && out2.empty() && out3.empty() && sent == s.size());
        got += in.to_string();
        CHECK(got == s);
        CHECK(in.backing_block_num() <= in.length() / 8160 + 2);     // reads fill blocks, they do not allocate one per readv
        // an unread portal keeps spare blocks for the next read and gives them back at the end
        IOPortal idle; CHECK(idle.append_from_file_descriptor(sv[1], 100) < 0 && idle.cached_block_num() == 1 && idle.empty());
        // single-buffer cut with a size hint: at most the refs needed to cover it go into the writev
        IOBuf w; for (int i = 0; i < 5; i++) w.append(s.substr(i * 8160, 8160));
        CHECK(w.backing_block_num() == 5);
        const ssize_t nw = w.cut_into_file_descriptor(sv[0], 9000); CHECK(nw == 2 * 8160 && w.length() == 3 * 8160);
        IOPortal drain; while (drain.length() < (size_t)nw) CHECK(drain.append_from_file_descriptor(sv[1], 1 << 16) > 0);
        CHECK(drain.to_string() == s.substr(0, 2 * 8160));
        // byte iterator: the h2 frame head walk (LoadUint8 / LoadUint32 / copy_and_forward / forward)
        b2::IOBufBytesIterator it(drain);
        CHECK(it.bytes_left() == 2 * 8160 && *it == (uint8_t)s[0]);
        ++it; CHECK(*it == (uint8_t)s[1]);
        char buf[16]; CHECK(it.copy_and_forward(buf, 16) == 16 && memcmp(buf, s.data() + 1, 16) == 0);
        CHECK(it.forward(8160) == 8160 && *it == (uint8_t)s[17 + 8160] && it.bytes_left() == 2 * 8160 - 17 - 8160);
        CHECK(it.forward(1 << 20) == 2 * 8160 - 17 - 8160 && !it && it.bytes_left() == 0);
        // Socket: read-size policy and the read-until-EAGAIN loop (input_messenger.cpp:243-261, :346-372)
        b2::Socket sock(7);
        CHECK(sock.once_read() == 4096 && sock.avg_msg_size() == 0);
        sock.OnMessageCut(1085); CHECK(sock.avg_msg_size() == 1085 && sock.once_read() == 16 * 1085);
        sock.OnMessageCut(2085); CHECK(sock.avg_msg_size() == (1085 * 9 + 2085) / 10);
        for (int i = 0; i < 200; i++) sock.OnMessageCut(1 << 20);
        CHECK(sock.once_read() == 524288);
        IOBuf big; big.append(s); big.append(s);
        size_t pushed = 0; while (!big.empty()) { const ssize_t nw = big.cut_into_file_descriptor(sv[0]); if (nw < 0) break; pushed += (size_t)nw; }
        bool eof = true;
        const ssize_t nread = sock.ReadUntilWouldBlock(sv[1], &eof);
        CHECK(nread == (ssize_t)pushed && !eof && !sock.Failed() && sock._read_buf.length() == pushed);
        CHECK(sock._read_buf.to_string() == (s + s).substr(0, pushed));
        close(sv[0]);
        CHECK(sock.ReadUntilWouldBlock(sv[1], &eof) == 0 && eof);                    // peer closed: EOF
        close(sv[1]);
    }
    CHECK(g_live_blocks == live_before);
    b2::iobuf::blockmem_allocate = b2::iobuf::default_alloc; b2::iobuf::blockmem_deallocate = b2::iobuf::default_free;
    printf("portal ok (readv into blocks, writev cuts, byte iterator)\n");
}


// IOBufAsZeroCopy{Input,Output}Stream, IOBufCutter, IOBufAppender (cases after test/iobuf_unittest.cpp:1240-1420 "as_zero_copy_*", "cutter", "appender")
static void test_iobuf_adapters() {
    using namespace b2;
    std::string s(20000, 'x'); for (size_t i = 0; i < s.size(); i++) s[i] = (char)('A' + i % 53);
    IOBuf b; b.append(s);
    {   // input stream: blocks come out in order; BackUp returns the tail of the last block; Skip crosses blocks; ByteCount follows
        IOBufAsZeroCopyInputStream in(b);
        const void* d; int n; std::string got;
        CHECK(in.Next(&d, &n) && n == 8160 && memcmp(d, s.data(), 8160) == 0 && in.ByteCount() == 8160);
        in.BackUp(100); CHECK(in.ByteCount() == 8060);
        CHECK(in.Next(&d, &n) && n == 100 && memcmp(d, s.data() + 8060, 100) == 0);
        CHECK(in.Skip(8160 + 10) && in.ByteCount() == 8160 * 2 + 10);
        CHECK(in.Next(&d, &n) && n == (int)(20000 - 8160 * 2 - 10) && memcmp(d, s.data() + 8160 * 2 + 10, (size_t)n) == 0);
        CHECK(!in.Next(&d, &n) && in.ByteCount() == 20000 && !in.Skip(1) && in.Skip(0));
        CHECK(b.length() == 20000);                                         // the source is never modified
    }
    {   // output stream: Next hands out block space, BackUp gives back what was not written; appending goes on where the stream stopped
        IOBuf o; o.append("head:");
        IOBufAsZeroCopyOutputStream out(&o);
        void* d; int n; size_t written = 0;
        while (written < 12000) { CHECK(out.Next(&d, &n) && n > 0); const size_t k = std::min((size_t)n, 12000 - written); memcpy(d, s.data() + written, k); written += k; if (k < (size_t)n) out.BackUp(n - (int)k); }
        CHECK(out.ByteCount() == 12000 && o.length() == 12005 && o.to_string() == "head:" + s.substr(0, 12000));
        o.append("|tail"); CHECK(o.to_string() == "head:" + s.substr(0, 12000) + "|tail");
        IOBuf e; IOBufAsZeroCopyOutputStream eo(&e); CHECK(eo.Next(&d, &n)); eo.BackUp(n); CHECK(e.empty() && eo.ByteCount() == 0);
    }
    {   // cutter
        IOBuf c(b); IOBufCutter cut(&c);
        char ch; CHECK(cut.fetch1() && *(const char*)cut.fetch1() == s[0] && cut.cut1(&ch) && ch == s[0] && cut.remaining_bytes() == 19999);
        char hdr[12]; CHECK(cut.copy_to(hdr, 12) == 12 && memcmp(hdr, s.data() + 1, 12) == 0 && cut.remaining_bytes() == 19999);
        IOBuf meta; std::string str; CHECK(cut.cutn(&meta, 8200) == 8200 && cut.cutn(&str, 50) == 50 && cut.pop_front(49) == 49);
        CHECK(meta.to_string() == s.substr(1, 8200) && str == s.substr(8201, 50) && cut.remaining_bytes() == 20000 - 1 - 8200 - 50 - 49);
        char rest[16]; CHECK(cut.cutn(rest, 16) == 16 && memcmp(rest, s.data() + 8300, 16) == 0);
        CHECK(cut.pop_front(1 << 30) == 20000 - 8316 && cut.remaining_bytes() == 0 && !cut.cut1(&ch) && cut.fetch1() == nullptr);
    }
    {   // appender
        IOBufAppender ap; std::string want;
        for (int i = 0; i < 9000; i++) { CHECK(ap.push_back((char)('a' + i % 26)) == 0); want.push_back((char)('a' + i % 26)); }
        CHECK(ap.append("-mid-", 5) == 0 && ap.append_decimal(-1234567) == 0 && ap.append(s) == 0); want += "-mid--1234567" + s;
        CHECK(ap.buf().length() == want.size() && ap.buf().to_string() == want);
        CHECK(ap.push_back('!') == 0); want.push_back('!');                  // keeps appending after buf()
        IOBuf target; target.append("old"); ap.move_to(target);
        CHECK(target.to_string() == want && ap.buf().empty());
    }
    printf("adapters ok (zero-copy input/output streams, cutter, appender)\n");
}

static int g_host_msgs = 0;
static void HostProcess(b2::InputMessageBase* base) {
    b2::MostCommonMessage* m = static_cast<b2::MostCommonMessage*>(base);
    CHECK(m->meta.length() == m->desc.meta_size && m->payload.length() == m->desc.body_size - m->desc.meta_size);
    g_host_msgs++;
    delete m;
}

static void test_messenger_gpu() {
    // pinned IOBuf blocks, as a brpc integration would install them
    b2::iobuf::blockmem_allocate = b2_block_alloc; b2::iobuf::blockmem_deallocate = b2_block_free;
    b2_options opt; memset(&opt, 0, sizeof opt);
    opt.device = 0; opt.max_batch_bytes = 8 << 20; opt.max_msgs = 1 << 16; opt.max_runs = 256; opt.max_body_size = 4 << 20;
    b2::GpuInputMessenger messenger(opt);
    b2_method echo = { "example.EchoService", "EchoService", "Echo", "example.EchoRequest", B2_HANDLER_ECHO, 1, 0, 0 };
    b2_method other = { "example.Other", "Other", "Call", "example.OtherRequest", B2_HANDLER_HOST, 0, 0, 0 };
    CHECK(messenger.AddMethod(echo) == 0 && messenger.AddMethod(other) == 1);
    messenger.SetHostProcess(HostProcess);

    const int kSockets = 16, kFrames = 40;
    std::vector<std::string> streams(kSockets);
    std::vector<uint8_t> f(1 << 17);
    for (int s = 0; s < kSockets; s++)
        for (int i = 0; i < kFrames; i++) {
            b2press_spec sp = { (s == 3 && i % 5 == 0) ? "example.Other" : "example.EchoService", (s == 3 && i % 5 == 0) ? "Call" : "Echo",
                                (uint32_t)(16 << (i % 9)), (uint32_t)(i % 3 == 0 ? 21 : 0), i % 2, i % 4 == 1, 20260921 };
            const size_t n = b2press_frame(&sp, ((uint64_t)s << 32) + i, f.data(), f.size());
            CHECK(n > 0);
            streams[s].append((const char*)f.data(), n);
        }
    streams[7][5000] ^= 0x40;   // corruption somewhere in socket 7's stream
    // feed in uneven chunks over several rounds, like reads that stop mid-frame
    std::vector<b2::Socket*> socks; std::vector<size_t> pos(kSockets, 0);
    for (int s = 0; s < kSockets; s++) socks.push_back(messenger.AddSocket(1000 + s));
    unsigned seed = 12345; int rounds = 0, total_msgs = 0;
    for (bool more = true; more; rounds++) {
        more = false;
        for (int s = 0; s < kSockets; s++) {
            seed = seed * 1103515245u + 12345u;
            const size_t n = std::min(streams[s].size() - pos[s], (size_t)(seed >> 16) % 9000);
            socks[s]->_read_buf.append(streams[s].data() + pos[s], n); pos[s] += n;
            if (pos[s] < streams[s].size()) more = true;
        }
        const int n = messenger.ProcessNewMessages(socks);
        CHECK(n >= 0); total_msgs += n;
    }
    // expectation: the oracle over each whole stream
    orc_config cfg; memset(&cfg, 0, sizeof cfg);
    b2_method ms[2] = { echo, other }; cfg.methods = ms; cfg.n_methods = 2;
    int checked = 0;
    for (int s = 0; s < kSockets; s++) {
        b2_run run = { 0, 0, (uint32_t)streams[s].size(), -1, 0 };
        b2_run_status rs; std::vector<b2_msg_desc> msgs(4096); std::vector<uint8_t> resp(streams[s].size() * 2 + (1 << 16));
        uint32_t nm = 0, rb = 0;
        CHECK(orc_process_batch(&cfg, (const uint8_t*)streams[s].data(), run.length, &run, 1, &rs, msgs.data(), 4096, &nm,
                                resp.data(), (uint32_t)resp.size(), &rb) == 0);
        if (rs.parse_error != B2_PARSE_ERROR_NOT_ENOUGH_DATA) {         // corrupted stream: the socket must be closed
            CHECK(socks[s]->Failed());
            continue;
        }
        CHECK(!socks[s]->Failed() && socks[s]->in_msgs() == nm && socks[s]->_read_buf.length() == streams[s].size() - rs.consumed);
        CHECK(socks[s]->_write_buf.length() == rb);
        CHECK(socks[s]->_write_buf.to_string() == std::string((const char*)resp.data(), rb));
        checked++;
    }
    CHECK(checked >= kSockets - 2 && g_host_msgs == 8);
    printf("messenger ok: %d sockets, %d messages in %d rounds, %d streams byte-identical to the oracle, %d host-handled\n",
           kSockets, total_msgs, rounds, checked, g_host_msgs);
}


// ---- seam (1): the Protocol table with the GPU-backed parse entry, driven like InputMessenger::CutInputMessage drives a handler ----
static void test_protocol_shim_gpu() {
    using namespace b2;
    CHECK(RegisterProtocol(PROTOCOL_B2_GPU, policy::GpuProtocol()) == 0);
    CHECK(RegisterProtocol(PROTOCOL_B2_GPU, policy::GpuProtocol()) == -1);            // once-only per type (protocol.cpp:90-93)
    Protocol none = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, CONNECTION_TYPE_ALL, "none" };
    CHECK(RegisterProtocol((ProtocolType)101, none) == -1 && RegisterProtocol((ProtocolType)200, policy::GpuProtocol()) == -1);
    const Protocol* proto = FindProtocol(PROTOCOL_B2_GPU);
    CHECK(proto && proto->support_server() && !proto->support_client() && FindProtocol(PROTOCOL_BAIDU_STD) == nullptr);
    b2_options opt; memset(&opt, 0, sizeof opt);
    opt.device = 0; opt.max_batch_bytes = 4 << 20; opt.max_msgs = 1 << 14; opt.max_runs = 8;
    b2_ctx* ctx = nullptr; CHECK(b2_ctx_create(&opt, &ctx) == B2_OK);
    b2_method echo = { "example.EchoService", "EchoService", "Echo", "example.EchoRequest", B2_HANDLER_ECHO, 1, 0, 0 };
    CHECK(b2_register_method(ctx, &echo) == 0);
    CHECK(b2_set_protocols(ctx, (1u << 1) | (1u << 2) | (1u << 3) | (1u << 12)) == B2_OK);
    policy::GpuParser parser(ctx, 4 << 20);
    // a stream of baidu_std echo requests, one hulu frame, one nshead frame, then garbage
    std::string stream; std::vector<uint8_t> f(1 << 16); std::vector<std::pair<int, size_t>> expect;   // (protocol, bytes behind the header)
    for (int i = 0; i < 30; i++) {
        b2press_spec sp = { "example.EchoService", "Echo", (uint32_t)(10 + 37 * i), 0, 1, 0, 20260921 };
        const size_t n = b2press_frame(&sp, 900 + i, f.data(), f.size());
        stream.append((const char*)f.data(), n); expect.push_back(std::make_pair(1, n - 12));
        if (i == 10) { std::string h("HULU"); uint32_t b = 25, m = 5; h.append((const char*)&b, 4); h.append((const char*)&m, 4); h.append(25, 'h'); stream += h; expect.push_back(std::make_pair(3, (size_t)25)); }
        if (i == 20) { std::string h(36, '\0'); uint32_t mg = 0xfb709394u, bl = 9; memcpy(&h[24], &mg, 4); memcpy(&h[32], &bl, 4); h.append(9, 'n'); stream += h; expect.push_back(std::make_pair(12, (size_t)9)); }
    }
    stream += "XXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXXX";
    Socket sock(42); IOBuf source; size_t fed = 0, got = 0; unsigned seed = 7; int fatal = -1;
    orc_config cfg; memset(&cfg, 0, sizeof cfg); cfg.methods = &echo; cfg.n_methods = 1; cfg.protocols = (1u << 1) | (1u << 2) | (1u << 3) | (1u << 12);
    while (fatal < 0) {
        ParseResult r = proto->parse(&source, &sock, false, &parser);
        if (r.is_ok()) {
            policy::GpuMessage* m = static_cast<policy::GpuMessage*>(r.message());
            CHECK(got < expect.size() && m->desc.protocol == expect[got].first && m->meta.length() + m->payload.length() == expect[got].second + (m->desc.protocol == 12 ? 36 : 0));
            if (m->desc.protocol == 1) CHECK(m->desc.status == B2_MSG_ECHOED && !m->reply.empty());
            got++;
            proto->process_request(m);                                   // writes the device-packed reply into sock._write_buf
        } else if (r.error() == PARSE_ERROR_NOT_ENOUGH_DATA) {
            if (fed == stream.size()) break;
            seed = seed * 1103515245u + 12345u;
            const size_t n = std::min(stream.size() - fed, (size_t)1 + (seed >> 16) % 3000);
            source.append(stream.data() + fed, n); fed += n;
        } else fatal = (int)r.error();
    }
    CHECK(got == expect.size() && fatal == PARSE_ERROR_TRY_OTHERS);          // the garbage tail: no handler takes it, the socket would be closed
    // every reply byte == the oracle's response stream for the same bytes
    b2_run run = { 0, 0, (uint32_t)stream.size(), -1, 0 }; b2_run_status rs; std::vector<b2_msg_desc> msgs(256); std::vector<uint8_t> resp(1 << 20); uint32_t nm = 0, rb = 0;
    CHECK(orc_process_batch(&cfg, (const uint8_t*)stream.data(), run.length, &run, 1, &rs, msgs.data(), 256, &nm, resp.data(), (uint32_t)resp.size(), &rb) == 0);
    CHECK(nm == expect.size() && sock._write_buf.to_string() == std::string((const char*)resp.data(), rb));
    b2_ctx_destroy(ctx);
    printf("protocol shim ok: %zu messages of 3 protocols cut through Protocol::parse, replies byte-identical to the oracle\n", got);
}

// ---- h2 / gRPC through GpuH2Messenger, every written byte against the oracle (same chunking fed to both) ----
static std::string h2_frame(int type, int flags, uint32_t sid, const std::string& payload) {
    std::string f; const uint32_t n = (uint32_t)payload.size();
    f.push_back((char)(n >> 16)); f.push_back((char)(n >> 8)); f.push_back((char)n); f.push_back((char)type); f.push_back((char)flags);
    f.push_back((char)(sid >> 24)); f.push_back((char)(sid >> 16)); f.push_back((char)(sid >> 8)); f.push_back((char)sid);
    return f + payload;
}
static std::string hp_lit(const std::string& n, const std::string& v) {      // literal header field without indexing, new name (RFC 7541 6.2.2)
    std::string o; o.push_back(0); o.push_back((char)n.size()); o += n; o.push_back((char)v.size()); o += v; return o;
}
static int g_h2_host = 0;
static void H2HostProcess(b2::InputMessageBase* base) { b2::H2Message* m = static_cast<b2::H2Message*>(base); CHECK(!m->headers.empty()); g_h2_host++; delete m; }

static void test_h2_messenger_gpu() {
    b2_options opt; memset(&opt, 0, sizeof opt);
    opt.device = 0; opt.max_batch_bytes = 8 << 20; opt.max_msgs = 1 << 14; opt.max_runs = 64; opt.max_resp_bytes = 32 << 20;
    b2::GpuH2Messenger messenger(opt);
    b2_method echo = { "example.EchoService", "EchoService", "Echo", "example.EchoRequest", B2_HANDLER_ECHO, 1, 0, 0 };
    CHECK(messenger.AddMethod(echo) == 0);
    messenger.SetHostProcess(H2HostProcess);
    const int kConns = 6, kCalls = 12;
    std::vector<std::string> streams(kConns);
    for (int c = 0; c < kConns; c++) {
        std::string& st = streams[c];
        st = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"; st += h2_frame(4, 0, 0, "");
        for (int k = 0; k < kCalls; k++) {
            const uint32_t sid = 1 + 2 * k;
            const bool other = (c == 2 && k % 4 == 1);                               // not a device-served method: goes to the host callback
            std::string hb = hp_lit(":method", "POST") + hp_lit(":scheme", "http") + hp_lit(":path", other ? "/example.Other/Call" : "/example.EchoService/Echo") +
                             hp_lit("content-type", k % 3 == 0 ? "application/grpc+proto" : "application/grpc") + hp_lit("te", "trailers");
            std::string msg(10 + 37 * k + (k == 8 ? 9000 : 0), (char)('a' + k));
            std::string pb = "\x0a"; { uint32_t n = (uint32_t)msg.size(); while (n >= 0x80) { pb.push_back((char)(n | 0x80)); n >>= 7; } pb.push_back((char)n); } pb += msg;
            std::string body; body.push_back(0); body.push_back((char)(pb.size() >> 24)); body.push_back((char)(pb.size() >> 16)); body.push_back((char)(pb.size() >> 8)); body.push_back((char)pb.size()); body += pb;
            st += h2_frame(1, 0x4, sid, hb);
            if (k % 5 == 2) { st += h2_frame(0, 0, sid, body.substr(0, 7)); st += h2_frame(0, 0x1, sid, body.substr(7)); }     // two DATA frames: body assembled in the slot
            else if (body.size() > 8000) { st += h2_frame(0, 0, sid, body.substr(0, 5000)); st += h2_frame(0, 0x1, sid, body.substr(5000)); }
            else st += h2_frame(0, 0x1, sid, body);
            if (k % 4 == 3) st += h2_frame(6, 0, 0, "pingpong");
            if (k % 6 == 5) st += h2_frame(8, 0, 0, std::string("\x00\x01\x00\x00", 4));
        }
    }
    std::vector<b2::Socket*> socks; std::vector<size_t> pos(kConns, 0);
    std::vector<orc_h2_conn*> oc(kConns); std::vector<std::string> obuf(kConns), expect(kConns);
    for (int c = 0; c < kConns; c++) { socks.push_back(messenger.AddConnection(500 + c)); oc[c] = orc_h2_conn_new(); }
    orc_config cfg; memset(&cfg, 0, sizeof cfg); b2_method ms[1] = { echo }; cfg.methods = ms; cfg.n_methods = 1;
    unsigned seed = 777; int rounds = 0, total = 0;
    std::vector<b2_h2_msg> om(256); std::vector<uint8_t> octrl(1 << 16), oblob(1 << 20), opack(1 << 18);
    for (bool more = true; more; rounds++) {
        more = false;
        for (int c = 0; c < kConns; c++) {
            seed = seed * 1103515245u + 12345u;
            const size_t n = std::min(streams[c].size() - pos[c], (size_t)(seed >> 16) % 5000);
            socks[c]->_read_buf.append(streams[c].data() + pos[c], n); obuf[c].append(streams[c].data() + pos[c], n); pos[c] += n;
            if (pos[c] < streams[c].size()) more = true;
        }
        const int n = messenger.ProcessNewMessages(socks);
        CHECK(n >= 0); total += n;
        for (int c = 0; c < kConns; c++) {                       // the same round through the oracle
            if (obuf[c].empty()) continue;
            uint32_t cons = 0, nm = 0, cl = 0, bl = 0, mfs = 0, sws = 0;
            const uint32_t err = orc_h2_consume(oc[c], &cfg, (const uint8_t*)obuf[c].data(), (uint32_t)obuf[c].size(), &cons, om.data(), 256, &nm,
                                                octrl.data(), (uint32_t)octrl.size(), &cl, oblob.data(), (uint32_t)oblob.size(), &bl, &mfs, &sws);
            CHECK(err == B2_PARSE_ERROR_NOT_ENOUGH_DATA);
            expect[c].append((const char*)octrl.data(), cl);
            for (uint32_t k = 0; k < nm; k++) {
                if (om[k].method_idx != 0) continue;            // the host callback's business
                std::string ct;
                for (uint32_t q = 0; q < om[k].headers_len;) {
                    const uint8_t* p = oblob.data() + om[k].headers_off + q; const uint32_t nl = p[0] | (p[1] << 8), vl = p[2] | (p[3] << 8);
                    if (nl == 12 && memcmp(p + 4, "content-type", 12) == 0) ct.assign((const char*)p + 4 + nl, vl);
                    q += 4 + nl + vl;
                }
                std::string blob = ct + std::string((const char*)oblob.data() + om[k].msg_off, om[k].msg_len);
                b2_h2_response r; memset(&r, 0, sizeof r);
                r.stream_id = om[k].stream_id; r.status_code = 200; r.flags = B2_H2_RESP_GRPC; r.content_type_len = (uint32_t)ct.size();
                r.body_off = (uint32_t)ct.size(); r.body_len = om[k].msg_len;
                const uint32_t pn = orc_h2_pack_response(oc[c], &r, (const uint8_t*)blob.data(), opack.data());
                expect[c].append((const char*)opack.data(), pn);
            }
            obuf[c].erase(0, cons);
        }
    }
    for (int c = 0; c < kConns; c++) {
        CHECK(!socks[c]->Failed() && socks[c]->_read_buf.length() == obuf[c].size());
        CHECK(socks[c]->_write_buf.length() == expect[c].size());
        CHECK(socks[c]->_write_buf.to_string() == expect[c]);
        orc_h2_conn_free(oc[c]);
    }
    CHECK(total == kConns * kCalls && g_h2_host == 3);
    printf("h2 messenger ok: %d connections, %d gRPC calls in %d rounds, every written byte identical to the oracle, %d host-handled\n", kConns, total, rounds, g_h2_host);
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    test_iobuf();
    test_portal_and_fd();
    test_iobuf_adapters();
    if (mode == "gpu") { test_messenger_gpu(); test_h2_messenger_gpu(); test_protocol_shim_gpu(); }
    return 0;
}
