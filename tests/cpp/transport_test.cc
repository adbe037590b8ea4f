// C++ tests / bench of b2::GpuTransport and the Socket write queue (brpc_b200/host/gpu_transport.h, input_messenger.h).
//   transport_test queue  (no GPU)  Socket::Write from many threads over a socketpair with a tiny send buffer: StartWrite /
//                                   KeepWrite / IsWriteComplete keep every request whole and every producer's order
//                                   (modelled on test/brpc_socket_unittest.cpp's multi-threaded write tests)
//   transport_test gpu              K connections over socketpairs, request streams fed in uneven chunks, pipelined rounds;
//                                   every byte the clients read back == the oracle's response stream (tests may link the oracle)
//   transport_test bench [mib] [rounds] [input] [resp] [groups]   pre-filled read regions (the message-processing path without socket syscalls on
//                                   the read side, like the reference arm), replies gathered by writev into /dev/null; one JSON line
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <time.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>
#include "../../brpc_b200/host/gpu_transport.h"
#include "../../oracle/b2_oracle.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

extern "C" {
typedef struct b2press_spec { const char* service; const char* method; uint32_t payload_bytes, attachment_bytes; int32_t payload_kind, checksum_type; uint64_t seed; } b2press_spec;
size_t b2press_frame(const b2press_spec* s, uint64_t index, uint8_t* out, size_t cap);
uint64_t b2press_fill_run(const b2press_spec* s, uint64_t* index, uint8_t* out, size_t run_bytes);
}
// run on (and first-touch pinned memory from) the CPUs next to the GPU: zero-copy PCIe reads that cross the socket interconnect lose most
// of their rate (bench.py does the same through NVML)
static std::vector<int> g_local_cpus;          // CPUs next to the GPU, in sysfs order (the first half of a socket's list = one hardware thread per core)
static void pin_thread_to_local_cpu(int k) {
    if (g_local_cpus.empty()) return;
    cpu_set_t one; CPU_ZERO(&one); CPU_SET(g_local_cpus[(size_t)k % g_local_cpus.size()], &one);
    sched_setaffinity(0, sizeof one, &one);
}
static void pin_to_gpu_numa(int device) {
    char bus[32] = {0};
    if (b2_device_pci_bus_id(device, bus, sizeof bus) != B2_OK) { fprintf(stderr, "(no NUMA pinning: bus id unknown)\n"); return; }
    for (char* p = bus; *p; p++) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
    char path[128]; snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r"); if (!f) return;
    char list[512] = {0}; if (!fgets(list, sizeof list, f)) { fclose(f); return; } fclose(f);
    cpu_set_t set; CPU_ZERO(&set); int n = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a, b; if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; c++) { CPU_SET(c, &set); g_local_cpus.push_back(c); n++; } } else if (sscanf(tok, "%d", &a) == 1) { CPU_SET(a, &set); g_local_cpus.push_back(a); n++; }
    }
    if (n) sched_setaffinity(0, sizeof set, &set);
    fprintf(stderr, "(pinned to %d CPUs next to GPU %d)\n", n, device);
}
static double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }
static void set_nonblock(int fd) { fcntl(fd, F_SETFL, fcntl(fd, F_GETFL) | O_NONBLOCK); }

static void test_write_queue() {
    int sv[2]; CHECK(socketpair(AF_UNIX, SOCK_STREAM, 0, sv) == 0);
    int sz = 4096; setsockopt(sv[0], SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
    set_nonblock(sv[0]);
    b2::Socket sock(7); sock.set_fd(sv[0]);
    const int kThreads = 8, kPer = 400;
    // every message: [u8 thread][u32 seq][u32 len][len bytes of (thread ^ seq ^ i)]
    std::vector<std::thread> prod;
    for (int t = 0; t < kThreads; t++) prod.emplace_back([&, t]() {
        for (int q = 0; q < kPer; q++) {
            const uint32_t len = 1 + (uint32_t)((t * 131 + q * 977) % 3000);
            std::string m(9 + len, '\0');
            m[0] = (char)t; memcpy(&m[1], &q, 4); memcpy(&m[5], &len, 4);
            for (uint32_t i = 0; i < len; i++) m[9 + i] = (char)(t ^ q ^ i);
            b2::IOBuf b; b.append(m.substr(0, m.size() / 2)); b.append(m.substr(m.size() / 2));      // two blocks' worth of refs
            CHECK(sock.Write(&b) == 0 && b.empty());
        }
    });
    std::vector<int> next(kThreads, 0); size_t total = 0; int got = 0;
    std::string acc; char buf[65536];
    while (got < kThreads * kPer) {
        const ssize_t n = read(sv[1], buf, sizeof buf);
        CHECK(n > 0); acc.append(buf, (size_t)n); total += (size_t)n;
        size_t p = 0;
        while (acc.size() - p >= 9) {
            const int t = (unsigned char)acc[p]; int q; uint32_t len; memcpy(&q, &acc[p + 1], 4); memcpy(&len, &acc[p + 5], 4);
            CHECK(t < kThreads && len <= 3000);
            if (acc.size() - p < 9 + len) break;
            CHECK(q == next[t]);                                         // per-producer order kept
            for (uint32_t i = 0; i < len; i++) CHECK(acc[p + 9 + i] == (char)(t ^ q ^ i));   // requests are never interleaved
            next[t]++; got++; p += 9 + len;
        }
        acc.erase(0, p);
    }
    for (auto& th : prod) th.join();
    CHECK(acc.empty() && sock.write_queue_empty() && !sock.Failed() && sock.keepwrite_rounds() > 0);
    // a dead peer fails the socket and drops what is queued
    close(sv[1]);
    b2::IOBuf big; big.append(std::string(1 << 20, 'z'));
    sock.Write(&big);
    b2::IOBuf more; more.append("x"); sock.Write(&more);
    CHECK(sock.Failed() && sock.write_queue_empty());
    close(sv[0]);
    printf("write queue ok: %d producers x %d requests, %zu bytes, order and integrity kept, %llu KeepWrite rounds\n", kThreads, kPer, total,
           (unsigned long long)sock.keepwrite_rounds());
}

static std::vector<std::string> make_streams(int k, int frames) {
    std::vector<std::string> streams(k); std::vector<uint8_t> f(1 << 17);
    for (int s = 0; s < k; s++)
        for (int i = 0; i < frames; i++) {
            b2press_spec sp = { "example.EchoService", i % 11 == 10 ? "Nope" : "Echo", (uint32_t)(16 << (i % 9)), (uint32_t)(i % 3 == 0 ? 21 : 0), i % 2, i % 4 == 1, 20260921 };
            const size_t n = b2press_frame(&sp, ((uint64_t)s << 32) + i, f.data(), f.size());
            CHECK(n > 0); streams[s].append((const char*)f.data(), n);
        }
    return streams;
}

static void test_transport_gpu(int input_mode, int resp_mode, bool use_sink = false) {
    b2::GpuTransport::Options o; memset(&o.ctx, 0, sizeof o.ctx);
    o.ctx.device = 0; o.ctx.max_batch_bytes = 20 << 20; o.ctx.max_msgs = 1 << 16; o.ctx.max_runs = 64;
    o.pipeline = 3; o.region_bytes = 1 << 20; o.max_connections = 16; o.input_mode = input_mode; o.resp_mode = resp_mode;
    b2::GpuTransport tr(o);
    b2_method echo = { "example.EchoService", "EchoService", "Echo", "example.EchoRequest", B2_HANDLER_ECHO, 1, 0, 0 };
    CHECK(tr.AddMethod(echo) == 0);
    const int K = 10;
    std::vector<std::string> streams = make_streams(K, 60);
    std::vector<int> cli(K); std::vector<b2::GpuTransport::Conn*> conns(K);
    for (int s = 0; s < K; s++) {
        int sv[2]; CHECK(socketpair(AF_UNIX, SOCK_STREAM, 0, sv) == 0);
        set_nonblock(sv[0]); set_nonblock(sv[1]);
        int big = 4 << 20; setsockopt(sv[0], SOL_SOCKET, SO_SNDBUF, &big, sizeof big); setsockopt(sv[1], SOL_SOCKET, SO_SNDBUF, &big, sizeof big);
        cli[s] = sv[0]; conns[s] = tr.AddConnection(500 + s, sv[1]); CHECK(conns[s]);
    }
    std::vector<size_t> pos(K, 0); std::vector<std::string> got(K);
    unsigned seed = 99; int total = 0, rounds = 0; char rb[1 << 16];
    auto drain = [&]() { for (int s = 0; s < K; s++) for (;;) { const ssize_t n = read(cli[s], rb, sizeof rb); if (n <= 0) break; got[s].append(rb, (size_t)n); } };
    // the reply sink: the gather list goes to writev as it is (zero-length entries included), resuming after partial writes
    if (use_sink) tr.SetReplySink([&](b2::GpuTransport::Conn* c, const struct iovec* v, size_t n) {
        std::vector<struct iovec> w(v, v + n); size_t at = 0;
        while (at < n) {
            const ssize_t k = writev(c->fd, w.data() + at, (int)std::min<size_t>(1024, n - at));
            if (k < 0) { CHECK(errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR); drain(); continue; }
            size_t left = (size_t)k;
            while (at < n && left >= w[at].iov_len) { left -= w[at].iov_len; at++; }
            if (at < n && left) { w[at].iov_base = (char*)w[at].iov_base + left; w[at].iov_len -= left; }
            if (k == 0) { while (at < n && w[at].iov_len == 0) at++; }
        }
    });
    for (bool more = true; more || rounds % 3; rounds++) {
        more = false;
        const uint32_t g = rounds % tr.pipeline();
        const int c = tr.Collect(g); CHECK(c >= 0); total += c;                    // the batch this group submitted three rounds ago
        for (int s = 0; s < K; s++) {
            if (conns[s]->group != g) { if (pos[s] < streams[s].size()) more = true; continue; }
            seed = seed * 1103515245u + 12345u;
            const size_t n = std::min(streams[s].size() - pos[s], (size_t)(seed >> 16) % 20000);
            size_t w = 0; while (w < n) { const ssize_t k = write(cli[s], streams[s].data() + pos[s] + w, n - w); if (k <= 0) break; w += (size_t)k; }
            pos[s] += w;
            if (pos[s] < streams[s].size()) more = true;
            bool eof; tr.ReadUntilWouldBlock(conns[s], &eof);
        }
        CHECK(tr.Submit(g) >= 0);
        drain();
    }
    for (uint32_t g = 0; g < tr.pipeline(); g++) { const int c = tr.Collect(g); CHECK(c >= 0); total += c; }
    for (int r = 0; r < 6; r++) { for (uint32_t g = 0; g < tr.pipeline(); g++) { CHECK(tr.Submit(g) >= 0); const int c = tr.Collect(g); CHECK(c >= 0); total += c; } drain(); }
    orc_config cfg; memset(&cfg, 0, sizeof cfg); cfg.methods = &echo; cfg.n_methods = 1;
    for (int s = 0; s < K; s++) {
        b2_run run = { 0, 0, (uint32_t)streams[s].size(), -1, 0 };
        b2_run_status rs; std::vector<b2_msg_desc> msgs(4096); std::vector<uint8_t> resp(streams[s].size() * 2 + (1 << 16)); uint32_t nm = 0, rbn = 0;
        CHECK(orc_process_batch(&cfg, (const uint8_t*)streams[s].data(), run.length, &run, 1, &rs, msgs.data(), 4096, &nm, resp.data(), (uint32_t)resp.size(), &rbn) == 0);
        CHECK(!conns[s]->sock.Failed() && conns[s]->sock.in_msgs() == nm && conns[s]->fill == streams[s].size() - rs.consumed);
        // the oracle packs replies back to back in message order: that is the byte stream the client must have read
        std::string want; for (uint32_t m = 0; m < nm; m++) want.append((const char*)resp.data() + msgs[m].resp_off, msgs[m].resp_len);
        CHECK(got[s].size() == want.size() && got[s] == want);
    }
    for (int s = 0; s < K; s++) close(cli[s]);
    printf("transport ok (input=%d resp=%d): %d connections, %d messages in %d pipelined rounds, reply streams byte-identical to the oracle\n", input_mode, resp_mode, K, total, rounds);
}

static int bench(int run_mib, int rounds, int input_mode, int resp_mode, int groups, int per_thread) {
    const int K = 64;
    pin_to_gpu_numa(0);
    b2::GpuTransport::Options o; memset(&o.ctx, 0, sizeof o.ctx);
    const uint32_t region = (uint32_t)run_mib << 20;
    const uint32_t per_group = (uint32_t)(K + groups - 1) / groups;
    o.ctx.device = 0;
    o.ctx.max_batch_bytes = input_mode == B2_INPUT_PULL ? per_group * region + (1u << 20) : (uint32_t)K * region + (1u << 20) /* COPY moves the arena span that holds the group's regions */;
    o.ctx.max_msgs = (uint32_t)((uint64_t)per_group * region / 1000 + 4096);
    o.ctx.max_runs = K; o.ctx.max_resp_bytes = o.ctx.max_batch_bytes + (64u << 20);
    o.pipeline = (uint32_t)groups; o.region_bytes = region; o.max_connections = K; o.input_mode = input_mode; o.resp_mode = resp_mode;
    b2::GpuTransport tr(o);
    b2_method echo = { "example.EchoService", "EchoService", "Echo", "example.EchoRequest", B2_HANDLER_ECHO, 1, 0, 0 };
    CHECK(tr.AddMethod(echo) == 0);
    const int devnull = open("/dev/null", O_WRONLY); CHECK(devnull >= 0);
    std::vector<b2::GpuTransport::Conn*> conns(K);
    std::vector<std::vector<uint8_t>> fresh(K);
    b2press_spec sp = { "example.EchoService", "Echo", 1024, 0, 0, 0, 20260921 };
    const size_t run_bytes = region - 4096;
    for (int s = 0; s < K; s++) {
        conns[s] = tr.AddConnection(s, devnull);
        fresh[s].resize(run_bytes);
        uint64_t idx = ((uint64_t)s << 32);
        b2press_fill_run(&sp, &idx, fresh[s].data(), run_bytes);
    }
    const bool no_write = getenv("B2_BENCH_NOWRITE") != nullptr;                       // (experiment: what the writev syscalls cost)
    struct alignas(64) Tally { uint64_t bytes = 0, iov = 0, msgs = 0; };               // one cache line per group: the groups' threads share nothing
    std::vector<Tally> tally(groups);
    tr.SetReplySink([&](b2::GpuTransport::Conn* c, const struct iovec* v, size_t n) {   // what KeepWrite does: one writev per <= 1024 references
        Tally& t = tally[c->group];
        if (!no_write) for (size_t i = 0; i < n; i += 1024) { const ssize_t w = writev(c->fd, v + i, (int)std::min<size_t>(1024, n - i)); if (w > 0) t.bytes += (uint64_t)w; }
        t.iov += n;
    });
    auto refill = [&](uint32_t g) { for (int s = 0; s < K; s++) if (conns[s]->group == g) { conns[s]->fill = 0; tr.Feed(conns[s], fresh[s].data(), run_bytes); } };
    // warm-up; it also tells where each connection's last complete frame ends: the timed rounds submit exactly that much, so nothing is
    // left to move to the region's front and the regions stay as they are (a real read would append behind the tail)
    std::vector<uint32_t> whole(K, 0);
    for (int g = 0; g < groups; g++) { refill(g); CHECK(tr.Submit(g) > 0); CHECK(tr.Collect(g) > 0); }
    for (int s = 0; s < K; s++) { whole[s] = (uint32_t)run_bytes - conns[s]->fill; CHECK(whole[s] > run_bytes / 2 && !conns[s]->sock.Failed()); }
    auto arm = [&](uint32_t g) { for (int s = 0; s < K; s++) if (conns[s]->group == g) conns[s]->fill = whole[s]; };
    for (int g = 0; g < groups; g++) { refill(g); arm(g); CHECK(tr.Submit(g) > 0); CHECK(tr.Collect(g) > 0); }
    // one host thread per PAIR of groups (each group owns a context): while the thread delivers one group's replies the other group's
    // batch is on the GPU — transfers, kernels and host work overlap inside a thread as well as across threads
    const double t0 = now_s();
    std::vector<std::thread> th;
    for (int g0 = 0; g0 < groups; g0 += per_thread) th.emplace_back([&, g0]() {
        const int g1 = (per_thread == 2 && g0 + 1 < groups) ? g0 + 1 : -1;
        if (!getenv("B2_BENCH_NOPIN")) pin_thread_to_local_cpu(1 + g0 / per_thread);     // one core per delivering thread
        arm(g0); CHECK(tr.Submit(g0) > 0);
        for (int r = 0; r < rounds; r++) {
            if (g1 >= 0) { arm(g1); CHECK(tr.Submit(g1) > 0); }
            int c = tr.Collect(g0); CHECK(c > 0); tally[g0].msgs += (uint64_t)c;
            if (r + 1 < rounds) { arm(g0); CHECK(tr.Submit(g0) > 0); }
            if (g1 >= 0) { c = tr.Collect(g1); CHECK(c > 0); tally[g1].msgs += (uint64_t)c; }
        }
    });
    for (auto& t : th) t.join();
    const double dt = now_s() - t0;
    uint64_t tm = 0, tb = 0, ti = 0; for (int g = 0; g < groups; g++) { tm += tally[g].msgs; tb += tally[g].bytes; ti += tally[g].iov; }
    double ws = 0, ds = 0, ss = 0; for (int g = 0; g < groups; g++) { ws += tr.stats(g).wait_s; ds += tr.stats(g).deliver_s; ss += tr.stats(g).submit_s; }
    fprintf(stderr, "(per group over the whole run incl. warm-up: submit %.1f ms, wait %.1f ms, deliver %.1f ms)\n", ss / groups * 1e3, ws / groups * 1e3, ds / groups * 1e3);
    printf("{\"via\": \"b2::GpuTransport (C++)\", \"msgs_per_s\": %.1f, \"rounds_per_group\": %d, \"groups\": %d, \"host_threads\": %d, \"connections\": %d, \"run_mib\": %d, \"input_mode\": %d, \"resp_mode\": %d, "
           "\"reply_bytes_written\": %llu, \"iovecs\": %llu, \"seconds\": %.4f}\n", tm / dt, rounds, groups, (groups + per_thread - 1) / per_thread, K, run_mib, input_mode, resp_mode,
           (unsigned long long)tb, (unsigned long long)ti, dt);
    close(devnull);
    return 0;
}

int main(int argc, char** argv) {
    signal(SIGPIPE, SIG_IGN);                      // as brpc's global init does: a dead peer is an EPIPE from writev, not a signal
    const std::string mode = argc > 1 ? argv[1] : "queue";
    if (mode == "queue") { test_write_queue(); return 0; }
    if (mode == "gpu") { test_transport_gpu(B2_INPUT_PULL, B2_RESP_BY_REF); test_transport_gpu(B2_INPUT_COPY, B2_RESP_COPY); test_transport_gpu(B2_INPUT_COPY, B2_RESP_BY_REF);
                         test_transport_gpu(B2_INPUT_PULL, B2_RESP_IOVEC, true); test_transport_gpu(B2_INPUT_COPY, B2_RESP_IOVEC, true); test_transport_gpu(B2_INPUT_PULL, B2_RESP_IOVEC, false);
                         test_transport_gpu(B2_INPUT_PULL, B2_RESP_BY_REF, true); return 0; }
    if (mode == "bench") return bench(argc > 2 ? atoi(argv[2]) : 4, argc > 3 ? atoi(argv[3]) : 20, argc > 4 ? atoi(argv[4]) : B2_INPUT_PULL, argc > 5 ? atoi(argv[5]) : B2_RESP_BY_REF, argc > 6 ? atoi(argv[6]) : 4, argc > 7 && atoi(argv[7]) == 1 ? 1 : 2);
    fprintf(stderr, "usage: transport_test queue|gpu|bench\n"); return 2;
}
