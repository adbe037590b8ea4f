// b2::GpuTransport — the host side of the GPU message path as a brpc Transport would run it:
//   Transport                          src/brpc/transport.h:26-64 (one per socket mode, chosen by TransportFactory)
//   RdmaEndpoint::PollCq               src/brpc/rdma/rdma_endpoint.cpp:1470-1591: the precedent — an endpoint that owns REGISTERED
//                                      receive memory, polls completions and hands ready bytes to InputMessenger::ProcessNewMessage
//   rdma::block_pool                   src/brpc/rdma/block_pool.h:74-105: registered memory carved per connection
//   InputMessenger::OnNewMessages      src/brpc/input_messenger.cpp:324-389: read until EAGAIN, then cut
//   Socket::Write / KeepWrite          src/brpc/socket.cpp:1604-1889: replies leave as IOBuf references gathered by writev
//
// Every connection reads straight into its own region of ONE pinned + mapped arena (b2_block_alloc): no IOBuf blocks, no gather
// copy.  Connections are split into `pipeline` groups; a round over group g is
//     ReadUntilWouldBlock (each connection)  ->  Submit(g)  ...  Collect(g)  ->  replies written, consumed bytes popped
// and while group g's batch is on the GPU the other groups are being read / collected, so the PCIe transfers, the kernels and the
// host's socket work overlap (b2_batch_submit / b2_batch_collect, one context per group).  A connection belongs to exactly one
// group and has at most one batch in flight: its input is processed serially, as brpc guarantees per socket.  Groups share nothing
// but the arena: each may be driven (Read / Submit / Collect) by its own thread.
// With the default modes (B2_INPUT_PULL + B2_RESP_BY_REF) no payload byte moves on the host or across PCIe: the kernels read
// headers and metas in place, and an echo reply is {prefix reference into the batch's pinned reply block, payload reference into
// the connection's own read region}, written out by one writev per <= 128 replies.
#pragma once
#include <sys/uio.h>
#include <time.h>
#include <unistd.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "input_messenger.h"

namespace b2 {

class GpuTransport {
public:
    struct Options {
        b2_options ctx;                      // per group; max_batch_bytes bounds the bytes one group submits per round
        uint32_t pipeline = 3;               // groups == contexts == batches in flight
        uint32_t region_bytes = 4u << 20;    // registered read region per connection: >= max_body_size + 12 of the frames it must carry
        uint32_t max_connections = 256;
        int input_mode = B2_INPUT_PULL, resp_mode = B2_RESP_BY_REF;
    };
    struct Conn {
        Socket sock; uint8_t* base; uint32_t cap, fill, group; int fd;
        explicit Conn(uint64_t id) : sock(id), base(nullptr), cap(0), fill(0), group(0), fd(-1) {}
    };
    // replies of one connection, in order, as an iovec list (2 entries per by-reference reply): the writev that KeepWrite would issue.
    // Default sink: IOBuf references + Socket::Write.  A sink must be DONE with the list when it returns — the entries point into the batch's
    // pinned reply block and the connection's read region, both reused by the next round: what a writev could not take (EAGAIN) has to be
    // copied or written before returning.  With B2_RESP_IOVEC the list is the device's own (entries of length 0 for unanswered messages).
    typedef std::function<void(Conn*, const struct iovec*, size_t)> ReplySink;
    typedef void (*Process)(InputMessageBase* msg);

    explicit GpuTransport(const Options& o) : _opt(o) {
        if (o.pipeline == 0 || o.pipeline > 16) throw std::runtime_error("GpuTransport: pipeline must be 1..16");
        const size_t arena = (size_t)o.region_bytes * o.max_connections;
        if (arena >= (1ull << 31)) throw std::runtime_error("GpuTransport: arena must stay below 2 GiB (32-bit batch offsets)");
        _arena = static_cast<uint8_t*>(b2_block_alloc(arena));
        if (!_arena) throw std::runtime_error("GpuTransport: b2_block_alloc failed");
        _arena_bytes = arena;
        for (uint32_t g = 0; g < o.pipeline; g++) {
            b2_ctx* c = nullptr;
            if (b2_ctx_create(&o.ctx, &c) != B2_OK || b2_set_modes(c, o.input_mode, o.resp_mode) != B2_OK) { Destroy(); throw std::runtime_error(std::string("GpuTransport: ") + b2_last_error()); }
            _ctx.push_back(c);
        }
        _groups.resize(o.pipeline); _inflight.assign(o.pipeline, 0); _runs.resize(o.pipeline); _live.resize(o.pipeline); _iov.resize(o.pipeline); _pop.resize(o.pipeline); _stats.resize(o.pipeline);
        _outstanding.reset(new std::atomic<int>[o.pipeline]); for (uint32_t g = 0; g < o.pipeline; g++) _outstanding[g].store(0);
    }
    ~GpuTransport() { Destroy(); }
    GpuTransport(const GpuTransport&) = delete;

    int AddMethod(const b2_method& m) { int r = -1; for (b2_ctx* c : _ctx) r = b2_register_method(c, &m); return r; }
    void SetServerIdentity(const char* ip_port) { for (b2_ctx* c : _ctx) b2_set_server_identity(c, ip_port); }
    void SetHostProcess(Process p) { _process = p; }
    void SetReplySink(ReplySink s) { _sink = s; }
    uint32_t pipeline() const { return _opt.pipeline; }

    Conn* AddConnection(uint64_t socket_id, int fd) {
        if (_conns.size() >= _opt.max_connections) return nullptr;
        std::unique_ptr<Conn> c(new Conn(socket_id));
        c->base = _arena + (size_t)_conns.size() * _opt.region_bytes; c->cap = _opt.region_bytes; c->fd = fd;
        c->group = (uint32_t)(_conns.size() % _opt.pipeline);
        c->sock.set_fd(fd);
        _groups[c->group].push_back(c.get());
        _conns.push_back(std::move(c));
        return _conns.back().get();
    }
    // Socket::DoRead into the registered region until the fd would block (the loop of OnNewMessages without the per-read parse)
    ssize_t ReadUntilWouldBlock(Conn* c, bool* eof) {
        ssize_t total = 0; *eof = false;
        while (c->fill < c->cap) {
            const ssize_t nr = ::read(c->fd, c->base + c->fill, c->cap - c->fill);
            if (nr > 0) { c->fill += (uint32_t)nr; total += nr; continue; }
            if (nr == 0) { *eof = true; break; }
            if (errno == EINTR) continue;
            if (errno != EAGAIN && errno != EWOULDBLOCK) c->sock.SetFailed(errno, "Fail to read");
            break;
        }
        return total;
    }
    // test / bench entry: bytes that "arrived" on the connection
    size_t Feed(Conn* c, const void* data, size_t n) { const size_t k = n < c->cap - c->fill ? n : c->cap - c->fill; memcpy(c->base + c->fill, data, k); c->fill += (uint32_t)k; return k; }

    // Enqueue group g's pending bytes on the GPU.  Returns the number of runs submitted (0 = nothing pending), -1 on an ABI error.
    int Submit(uint32_t g) {
        if (_inflight[g]) return -1;
        const double t_sub0 = mono_s();
        std::vector<b2_run>& runs = _runs[g]; std::vector<Conn*>& live = _live[g];
        runs.clear(); live.clear();
        uint64_t bytes = 0;
        for (Conn* c : _groups[g]) {
            if (c->sock.Failed() || c->fill == 0) continue;
            if (bytes + c->fill > _opt.ctx.max_batch_bytes || runs.size() >= _opt.ctx.max_runs) break;      // the rest waits for the next round
            b2_run r; r.socket_id = c->sock.id(); r.offset = (uint32_t)(c->base - _arena); r.length = c->fill;
            r.preferred_proto = c->sock.preferred_index(); r.flags = 0;
            runs.push_back(r); live.push_back(c); bytes += c->fill;
        }
        if (runs.empty()) return 0;
        // (regions are region_bytes apart, a multiple of 16; PULL reads them where they are, COPY moves the arena span that holds them)
        const uint32_t span_end = runs.back().offset + runs.back().length;
        const int rc = b2_batch_submit(_ctx[g], _arena, span_end, runs.data(), (uint32_t)runs.size());
        if (rc != B2_OK) return -1;
        _stats[g].submit_s += mono_s() - t_sub0;
        _inflight[g] = 1;
        return (int)runs.size();
    }
    // Wait for group g's batch; deliver its messages (replies through Socket::Write / the sink, host-handled ones through the
    // process callback), pop what was consumed.  Returns the number of messages cut, -1 on an ABI error.
    int Collect(uint32_t g) {
        if (!_inflight[g]) return 0;
        _inflight[g] = 0;
        b2_batch_result res;
        const double t_wait0 = mono_s();
        if (b2_batch_collect(_ctx[g], &res) != B2_OK) return -1;
        const double t_wait1 = mono_s();
        _stats[g].wait_s += t_wait1 - t_wait0;
        std::vector<b2_run>& runs = _runs[g]; std::vector<Conn*>& live = _live[g];
        // one external block over the batch's pinned reply area, one per connection region: every reply is two references
        std::atomic<int>* outstanding = &_outstanding[g];
        std::vector<struct iovec>& _iov = this->_iov[g]; std::vector<std::pair<Conn*, uint32_t>>& _pop = this->_pop[g];
        IOBuf::Block* resp_blk = nullptr;
        if (!_sink && res.resp_bytes) { outstanding->fetch_add(1); resp_blk = IOBuf::create_external_block(const_cast<uint8_t*>(res.resp), res.resp_bytes, [outstanding](void*) { outstanding->fetch_sub(1); }); }
        for (uint32_t k = 0; k < res.n_runs; k++) {
            Conn* c = live[k]; Socket* s = &c->sock;
            const b2_run_status& st = res.runs[k];
            s->AddInputBytes(st.consumed); s->AddInputMessages(st.n_msgs); s->set_preferred_index(st.preferred_proto);
            if (res.iov && _sink) {
                // B2_RESP_IOVEC: the device wrote the gather list of this connection's replies (what cut_multiple_into_file_descriptor
                // would assemble, iobuf.cpp:954-992): no per-message work unless the run holds messages the device did not answer
                s->OnMessagesCut(st.consumed, st.n_msgs);
                if (st.n_msgs) _sink(c, reinterpret_cast<const struct iovec*>(res.iov) + 2 * (size_t)st.first_msg, 2 * (size_t)st.n_msgs);
                if (st.n_unanswered)
                    for (uint32_t m = st.first_msg; m < st.first_msg + st.n_msgs; m++) {
                        const b2_msg_desc& d = res.msgs[m];
                        if (d.status == B2_MSG_BAD_META) s->SetFailed(1003 /*EREQUEST*/, "Fail to parse RpcMeta");
                        else if (d.status != B2_MSG_ECHOED && d.status != B2_MSG_ERROR_REPLIED && _process) HandOver(s, d);
                    }
                FinishRun(c, s, st, _pop);
                continue;
            }
            IOBuf::Block* reg_blk = nullptr;
            if (!_sink && st.n_msgs) { outstanding->fetch_add(1); reg_blk = IOBuf::create_external_block(c->base, c->fill, [outstanding](void*) { outstanding->fetch_sub(1); }); }
            _iov.clear();
            IOBuf out;
            for (uint32_t m = st.first_msg; m < st.first_msg + st.n_msgs; m++) {
                const b2_msg_desc& d = res.msgs[m];
                s->OnMessageCut(12u + d.body_size);
                if (d.status == B2_MSG_ECHOED || d.status == B2_MSG_ERROR_REPLIED) {
                    b2_resp_ref from_iov;                                              // (B2_RESP_IOVEC without a sink: the pair says the same)
                    if (res.iov) {
                        const b2_iovec& a = res.iov[2 * (size_t)m]; const b2_iovec& b = res.iov[2 * (size_t)m + 1];
                        from_iov.prefix_len = (uint32_t)a.iov_len; from_iov.src_len = (uint32_t)b.iov_len; from_iov.reserved = 0;
                        from_iov.src_off = b.iov_len ? (uint32_t)(static_cast<const uint8_t*>(b.iov_base) - _arena) : 0;
                    }
                    const b2_resp_ref* rf = res.iov ? &from_iov : res.refs ? res.refs + m : nullptr;
                    const uint32_t plen = (rf && rf->src_len) ? rf->prefix_len : d.resp_len;
                    if (_sink) {
                        struct iovec v; v.iov_base = const_cast<uint8_t*>(res.resp) + d.resp_off; v.iov_len = plen; _iov.push_back(v);
                        if (rf && rf->src_len) { v.iov_base = _arena + rf->src_off; v.iov_len = rf->src_len; _iov.push_back(v); }
                    } else {
                        out.append_block_range(resp_blk, d.resp_off, plen);
                        if (rf && rf->src_len) out.append_block_range(reg_blk, rf->src_off - runs[k].offset, rf->src_len);
                        if (out.backing_block_num() >= 192) s->Write(&out);            // (one writev takes up to 256 references)
                    }
                } else if (d.status == B2_MSG_BAD_META) {
                    s->SetFailed(1003 /*EREQUEST*/, "Fail to parse RpcMeta");          // baidu_rpc_protocol.cpp:577-582
                } else if ((d.status == B2_MSG_HOST || d.status == B2_MSG_STREAM_FRAME || d.status == B2_MSG_UNSUPPORTED) && _process) HandOver(s, d);
            }
            if (_sink) { if (!_iov.empty()) _sink(c, _iov.data(), _iov.size()); }
            else if (!out.empty()) s->Write(&out);
            IOBuf::release_external_block(reg_blk);
            FinishRun(c, s, st, _pop);
        }
        IOBuf::release_external_block(resp_blk);
        // the references point into the reply block and the read regions: both are reused next round, so the writes must be done
        // (KeepWrite runs to completion on the default executor; a caller-set executor must drain before the next Submit)
        while (outstanding->load(std::memory_order_acquire) != 0) sched_yield();
        for (auto& pc : _pop) {                                                         // pop_front(consumed): the partial tail moves to the region's start
            Conn* c = pc.first; const uint32_t used = pc.second;
            if (used && used < c->fill) memmove(c->base, c->base + used, c->fill - used);
            c->fill -= used;
        }
        _pop.clear();
        _stats[g].deliver_s += mono_s() - t_wait1; _stats[g].batches++;
        return (int)res.n_msgs;
    }
    // where a group's host thread spent its time: in b2_batch_submit, waiting in b2_batch_collect, delivering messages
    struct GroupStats { double submit_s = 0, wait_s = 0, deliver_s = 0; uint64_t batches = 0; };
    const GroupStats& stats(uint32_t g) const { return _stats[g]; }
    const std::vector<std::unique_ptr<Conn>>& connections() const { return _conns; }
    uint8_t* arena() const { return _arena; }

private:
    // a message the device left to the host (B2_HANDLER_HOST methods, stream frames, other codecs / protocols): handed to the process callback
    void HandOver(Socket* s, const b2_msg_desc& d) {
        MostCommonMessage* msg = new MostCommonMessage;
        msg->socket = s; msg->desc = d;
        const uint8_t* f = _arena + d.frame_off;                                        // (copied: the region is reused after this round)
        msg->meta.append(f + 12, d.meta_size); msg->payload.append(f + 12 + d.meta_size, d.body_size - d.meta_size);
        _process(msg);
    }
    static void FinishRun(Conn* c, Socket* s, const b2_run_status& st, std::vector<std::pair<Conn*, uint32_t>>& pop) {
        if (st.parse_error != B2_PARSE_ERROR_NOT_ENOUGH_DATA)                           // input_messenger.cpp:227-239
            s->SetFailed(22 /*EINVAL*/, std::string("Close socket: ") + ParseErrorToString((ParseError)st.parse_error));
        else if (st.consumed == 0 && c->fill == c->cap)
            s->SetFailed(22, std::string("Close socket: ") + ParseErrorToString(PARSE_ERROR_TOO_BIG_DATA) + " (one frame exceeds the connection's read region)");
        pop.push_back(std::make_pair(c, st.consumed));
    }
    static double mono_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + (double)t.tv_nsec * 1e-9; }
    void Destroy() { for (b2_ctx* c : _ctx) b2_ctx_destroy(c); _ctx.clear(); if (_arena) { b2_block_free(_arena); _arena = nullptr; } }
    Options _opt; uint8_t* _arena = nullptr; size_t _arena_bytes = 0;
    std::vector<b2_ctx*> _ctx; std::vector<std::unique_ptr<Conn>> _conns; std::vector<std::vector<Conn*>> _groups;
    std::vector<char> _inflight; std::vector<std::vector<b2_run>> _runs; std::vector<std::vector<Conn*>> _live;
    std::vector<std::vector<struct iovec>> _iov; std::vector<std::vector<std::pair<Conn*, uint32_t>>> _pop;      // per group: groups may be driven by different threads
    std::vector<GroupStats> _stats;
    std::unique_ptr<std::atomic<int>[]> _outstanding; Process _process = nullptr; ReplySink _sink;
};

}  // namespace b2
