// b2::GpuInputMessenger — C++ host side of the GPU message path, mirroring the reference types
// it stands in for (same names, same contracts):
//   ParseError / ParseResult        src/brpc/parse_result.h:25-74
//   InputMessageBase                src/brpc/input_message_base.h:29-64
//   MostCommonMessage               src/brpc/policy/most_common_message.h:33-49
//   Socket (_read_buf, preferred_index, SetFailed, Write)   src/brpc/socket.h
//   InputMessenger::OnNewMessages / ProcessNewMessage       src/brpc/input_messenger.cpp:206-389
// One ProcessNewMessages() call == one OnNewMessages round over every readable socket, with the
// cut loop, RpcMeta decode, echo service and response packing done by the CUDA kernels behind
// include/b2rpc.h.  There is no parsing in this file.
#pragma once
#include <errno.h>
#include <poll.h>
#include <sched.h>
#include <stdint.h>
#include <atomic>
#include <functional>
#include <string.h>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/b2rpc.h"
#include "iobuf.h"

namespace b2 {

enum ParseError {
    PARSE_OK = 0,
    PARSE_ERROR_TRY_OTHERS,
    PARSE_ERROR_NOT_ENOUGH_DATA,
    PARSE_ERROR_TOO_BIG_DATA,
    PARSE_ERROR_NO_RESOURCE,
    PARSE_ERROR_ABSOLUTELY_WRONG,
};
inline const char* ParseErrorToString(ParseError e) {
    switch (e) {
    case PARSE_OK: return "ok";
    case PARSE_ERROR_TRY_OTHERS: return "try other protocols";
    case PARSE_ERROR_NOT_ENOUGH_DATA: return "not enough data";
    case PARSE_ERROR_TOO_BIG_DATA: return "too big data";
    case PARSE_ERROR_NO_RESOURCE: return "no resource for the message";
    case PARSE_ERROR_ABSOLUTELY_WRONG: return "absolutely wrong message";
    }
    return "unknown ParseError";
}

class Socket;
struct InputMessageBase {
    virtual ~InputMessageBase() {}
    Socket* socket = nullptr;
    int64_t received_us = 0;
};
struct MostCommonMessage : public InputMessageBase {
    IOBuf meta;
    IOBuf payload;
    b2_msg_desc desc;     // the decoded RpcMeta / StreamFrameMeta fields, produced on the GPU
};

class Socket {
public:
    explicit Socket(uint64_t id) : _id(id) {}
    uint64_t id() const { return _id; }
    IOPortal _read_buf;                  // bytes read from the fd, not yet cut (Socket::_read_buf)
    IOBuf _write_buf;                    // what Socket::Write would hand to writev, in order
    // Socket::DoRead (src/brpc/socket.cpp:2042-2122): one readv of at most `size_hint' bytes into the read buffer's blocks
    ssize_t DoRead(int fd, size_t size_hint) { return _read_buf.append_from_file_descriptor(fd, size_hint); }
    // Socket::DoWrite (:1856-1889): one writev over the queued reply blocks; returns what writev returned
    ssize_t DoWrite(int fd) { return _write_buf.cut_into_file_descriptor(fd); }
    // The read-size policy of InputMessenger::OnNewMessages (input_messenger.cpp:346-353): 16 x the average message size,
    // clamped to [MIN_ONCE_READ 4 KiB, MAX_ONCE_READ 512 KiB]; the average is the windowed mean ProcessNewMessage keeps (:243-261).
    static const size_t MIN_ONCE_READ = 4096, MAX_ONCE_READ = 524288, MSG_SIZE_WINDOW = 10;
    size_t once_read() const { const size_t n = _avg_msg_size * 16; return n < MIN_ONCE_READ ? MIN_ONCE_READ : n > MAX_ONCE_READ ? MAX_ONCE_READ : n; }
    void OnMessageCut(size_t msg_bytes) { _avg_msg_size = _avg_msg_size ? (_avg_msg_size * (MSG_SIZE_WINDOW - 1) + msg_bytes) / MSG_SIZE_WINDOW : msg_bytes; }
    // n messages of `bytes` in total were cut at once (the device cut them): the same smoothing, applied with their mean size
    void OnMessagesCut(size_t bytes, size_t n) { if (!n) return; const size_t mean = bytes / n; for (size_t k = n < 64 ? n : 64; k; k--) OnMessageCut(mean); }
    size_t avg_msg_size() const { return _avg_msg_size; }
    // Read until the fd would block (or EOF / an error): the inner loop of OnNewMessages without the per-read parse — parsing is
    // batched over all readable sockets by the messenger.  Returns the bytes read; *eof is set on a zero-length read.
    ssize_t ReadUntilWouldBlock(int fd, bool* eof) {
        ssize_t total = 0; *eof = false;
        for (;;) {
            const ssize_t nr = DoRead(fd, once_read());
            if (nr > 0) { total += nr; continue; }
            if (nr == 0) { *eof = true; break; }
            if (errno == EINTR) continue;
            if (errno != EAGAIN && errno != EWOULDBLOCK) SetFailed(errno, "Fail to read");
            break;
        }
        if (_read_buf.empty()) _read_buf.return_cached_blocks();   // good timing to give spare blocks back (:245-251)
        return total;
    }
    int preferred_index() const { return _preferred_index; }
    void set_preferred_index(int i) { _preferred_index = i; }
    bool Failed() const { return _failed; }
    int error_code() const { return _error_code; }
    const std::string& error_text() const { return _error_text; }
    void SetFailed(int error_code, const std::string& text) { bool f = false; if (_failed.compare_exchange_strong(f, true)) { _error_code = error_code; _error_text = text; } }
    // ---- Socket::Write (src/brpc/socket.cpp:1604-1679) -> StartWrite (:1681-1778) -> KeepWrite (:1782-1866) -> DoWrite (:1868-1889).
    // Many threads may Write() to one socket; writes are wait-free for all but one of them: every request is pushed onto an
    // atomic stack (_write_head).  The thread that finds the stack empty becomes THE writer: it writes its own request once, in
    // place; whatever is left (partial write, or requests other threads pushed meanwhile) is finished by KeepWrite, which reverses
    // the newly pushed part of the stack into FIFO order (IsWriteComplete, :1121-1180) and gathers up to 256 requests into one
    // writev (IOBuf::cut_multiple_into_file_descriptor).  In brpc KeepWrite runs in a new bthread; here it runs on the executor
    // set with SetKeepWriteExecutor (default: the writer's own thread, polling the fd while it would block).
    // A socket without an fd (tests, or a transport that ships _write_buf itself) just queues into _write_buf.
    // (`next` is written by the pusher after its exchange and read by the writer that waits for it to leave UNCONNECTED: an atomic here,
    //  where butil gets away with a plain pointer on the platforms it supports)
    struct WriteRequest { IOBuf data; std::atomic<WriteRequest*> next; };
    void set_fd(int fd) { _fd = fd; }
    int fd() const { return _fd; }
    typedef std::function<void(std::function<void()>)> KeepWriteExecutor;
    void SetKeepWriteExecutor(KeepWriteExecutor e) { _keepwrite_exec = e; }
    int Write(IOBuf* data) {
        if (_failed) return -1;
        if (_fd < 0) { _write_buf.append(data->movable()); return 0; }
        if (data->empty()) return 0;
        WriteRequest* req = new WriteRequest; req->data.swap(*data); req->next = unconnected();
        return StartWrite(req);
    }
    uint64_t keepwrite_rounds() const { return _keepwrite_rounds.load(); }
    bool write_queue_empty() const { return _write_head.load(std::memory_order_acquire) == nullptr; }
private:
    static WriteRequest* unconnected() { return reinterpret_cast<WriteRequest*>(uintptr_t(-1)); }   // WriteRequest::UNCONNECTED
    int StartWrite(WriteRequest* req) {
        WriteRequest* const prev_head = _write_head.exchange(req, std::memory_order_release);
        if (prev_head != nullptr) { req->next = prev_head; return 0; }           // someone is writing: it will find this request
        req->next = nullptr;                                                      // we are the writer
        const ssize_t nw = req->data.cut_into_file_descriptor(_fd);               // write once in the calling thread
        if (nw < 0 && errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) { const int e = errno; SetFailed(e, "Fail to write"); ReleaseAllFailedWriteRequests(req); return -1; }
        if (IsWriteComplete(req, true, nullptr)) { delete req; return 0; }
        if (_keepwrite_exec) _keepwrite_exec([this, req]() { KeepWrite(req); }); else KeepWrite(req);
        return 0;
    }
    // true when everything up to and including old_head is written and nobody pushed more; otherwise the newly pushed requests are
    // linked behind old_head in the order they were pushed and *new_tail is the last of them
    bool IsWriteComplete(WriteRequest* old_head, bool singular_node, WriteRequest** new_tail) {
        WriteRequest* new_head = old_head; WriteRequest* desired = nullptr; bool return_when_no_more = true;
        if (!old_head->data.empty() || !singular_node) { desired = old_head; return_when_no_more = false; }
        if (_write_head.compare_exchange_strong(new_head, desired, std::memory_order_acquire)) { if (new_tail) *new_tail = old_head; return return_when_no_more; }
        WriteRequest* tail = nullptr; WriteRequest* p = new_head;                 // someone pushed: new_head -> ... -> old_head, reverse it
        do {
            while (p->next == unconnected()) sched_yield();                      // its pusher is between the exchange and the link
            WriteRequest* const saved_next = p->next; p->next = tail; tail = p; p = saved_next;
        } while (p != old_head);
        old_head->next = tail;
        if (new_tail) *new_tail = new_head;
        return false;
    }
    void KeepWrite(WriteRequest* req) {
        WriteRequest* cur_tail = nullptr;
        for (;;) {
            _keepwrite_rounds.fetch_add(1, std::memory_order_relaxed);
            if (req->next != nullptr && req->data.empty()) { WriteRequest* const saved = req; req = req->next; delete saved; }
            IOBuf* pieces[IOBuf::IOV_MAX_REFS]; size_t n = 0;                     // DoWrite: up to 256 queued requests in one writev
            for (WriteRequest* p = req; p != nullptr && n < IOBuf::IOV_MAX_REFS; p = p->next) pieces[n++] = &p->data;
            const ssize_t nw = IOBuf::cut_multiple_into_file_descriptor(_fd, pieces, n);
            if (nw < 0) {
                if (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) { struct pollfd pf = { _fd, POLLOUT, 0 }; poll(&pf, 1, 100); }   // WaitEpollOut
                else { const int e = errno; SetFailed(e, "Fail to keep-write"); ReleaseAllFailedWriteRequests(req); return; }
            }
            while (req->next != nullptr && req->data.empty()) { WriteRequest* const saved = req; req = req->next; delete saved; }
            if (cur_tail == nullptr) for (cur_tail = req; cur_tail->next != nullptr; cur_tail = cur_tail->next) {}
            if (IsWriteComplete(cur_tail, req == cur_tail, &cur_tail)) { delete req; return; }
        }
    }
    void ReleaseAllFailedWriteRequests(WriteRequest* req) {                        // (:1056-1090) drop everything queued, now and until the stack is empty
        WriteRequest* cur_tail = nullptr;
        for (;;) {
            while (req->next != nullptr) { WriteRequest* const saved = req; req = req->next; delete saved; }
            req->data.clear();
            if (cur_tail == nullptr) cur_tail = req;
            if (IsWriteComplete(cur_tail, req == cur_tail, &cur_tail)) { delete req; return; }
        }
    }
public:
    void AddInputBytes(size_t n) { _in_bytes += n; }
    void AddInputMessages(size_t n) { _in_msgs += n; }
    uint64_t in_bytes() const { return _in_bytes; }
    uint64_t in_msgs() const { return _in_msgs; }
private:
    uint64_t _id; int _preferred_index = -1; std::atomic<bool> _failed{false}; int _error_code = 0; std::string _error_text;
    uint64_t _in_bytes = 0, _in_msgs = 0; size_t _avg_msg_size = 0;
    int _fd = -1; std::atomic<WriteRequest*> _write_head{nullptr}; std::atomic<uint64_t> _keepwrite_rounds{0}; KeepWriteExecutor _keepwrite_exec;
};

class GpuInputMessenger {
public:
    typedef void (*Process)(InputMessageBase* msg);   // InputMessageHandler::Process, input_messenger.h:51-57

    explicit GpuInputMessenger(const b2_options& opt) : _cap(opt.max_batch_bytes) {
        const uint64_t max_body = opt.max_body_size ? opt.max_body_size : (64ull << 20);
        if (max_body + 12 + 32 > _cap) throw std::runtime_error("GpuInputMessenger: max_batch_bytes must hold one frame of max_body_size (+44 bytes)");
        if (b2_ctx_create(&opt, &_ctx) != B2_OK) throw std::runtime_error(std::string("b2_ctx_create: ") + b2_last_error());
        _batch = static_cast<uint8_t*>(b2_block_alloc(_cap));
        if (!_batch) { b2_ctx_destroy(_ctx); throw std::runtime_error("b2_block_alloc failed"); }
    }
    ~GpuInputMessenger() { b2_block_free(_batch); b2_ctx_destroy(_ctx); }
    GpuInputMessenger(const GpuInputMessenger&) = delete;

    int AddMethod(const b2_method& m) { return b2_register_method(_ctx, &m); }   // what Server::AddService feeds
    void SetServerIdentity(const char* ip_port) { b2_set_server_identity(_ctx, ip_port); }
    // process callback for messages whose user code runs on the host (B2_MSG_HOST) and for stream frames
    void SetHostProcess(Process p) { _process = p; }
    Socket* AddSocket(uint64_t id) { auto& s = _sockets[id]; if (!s) s.reset(new Socket(id)); return s.get(); }
    b2_ctx* ctx() { return _ctx; }

    // One round over the readable sockets.  Returns the number of messages cut, -1 on an ABI error.
    // Every readable socket is served: when the batch buffer fills up the batch is processed and the round goes on with the next
    // one; a socket with more pending bytes than a batch holds submits a prefix (the cut loop answers NOT_ENOUGH_DATA for a frame
    // cut short, the rest follows next round) and is failed with TOO_BIG_DATA only when a single frame can never fit; a batch
    // that exceeds the context's message / response capacity is split and retried.
    int ProcessNewMessages(const std::vector<Socket*>& readable) {
        int total_msgs = 0; size_t i = 0;
        while (i < readable.size()) {
            std::vector<b2_run> runs; std::vector<Socket*> live; std::vector<bool> truncated;
            size_t total = 0;
            for (; i < readable.size(); i++) {
                Socket* s = readable[i];
                if (s->Failed() || s->_read_buf.empty()) continue;
                size_t n = s->_read_buf.length(); bool cut = false;
                if (n + 16 > _cap) { n = (_cap - 16) & ~(size_t)15; cut = true; }
                if (total + n + 16 > _cap) { if (runs.empty()) { n = (_cap - 16 - total) & ~(size_t)15; cut = true; } else break; }   // next batch of this round
                s->_read_buf.copy_to(_batch + total, n, 0);       // gather the (pinned) blocks into the batch buffer
                b2_run r; r.socket_id = s->id(); r.offset = (uint32_t)total; r.length = (uint32_t)n;
                r.preferred_proto = s->preferred_index(); r.flags = 0;
                runs.push_back(r); live.push_back(s); truncated.push_back(cut);
                total = (total + n + 15) & ~(size_t)15;
            }
            if (runs.empty()) break;
            const int n = ProcessRuns(runs, live, truncated, 0, runs.size(), total);
            if (n < 0) return -1;
            total_msgs += n;
        }
        return total_msgs;
    }

private:
    // runs [lo, hi) of the gathered batch through the ABI; on B2_E_CAPACITY the range is halved (a single run: its length)
    int ProcessRuns(std::vector<b2_run>& runs, std::vector<Socket*>& live, std::vector<bool>& truncated, size_t lo, size_t hi, size_t total) {
        b2_batch_result res;
        const int rc = b2_process_batch(_ctx, _batch, (uint32_t)total, runs.data() + lo, (uint32_t)(hi - lo), &res);
        if (rc == B2_E_CAPACITY) {
            if (hi - lo > 1) {
                const size_t mid = lo + (hi - lo) / 2;
                const int a = ProcessRuns(runs, live, truncated, lo, mid, total); if (a < 0) return a;
                const int b = ProcessRuns(runs, live, truncated, mid, hi, total); if (b < 0) return b;
                return a + b;
            }
            if (runs[lo].length <= 4096) return -1;
            runs[lo].length = (runs[lo].length / 2) & ~15u; truncated[lo] = true;       // fewer messages / response bytes per call
            return ProcessRuns(runs, live, truncated, lo, hi, total);
        }
        if (rc == B2_E_CUDA || rc == B2_E_NO_DEVICE) {
            // a device error must not stall the connections — the bytes stay in _read_buf and the sockets of this
            // batch are handed to the host parse path (brpc's own Protocol::parse) when one is set; without one they are failed
            for (size_t k = lo; k < hi; k++) { if (_host_parse) _host_parse(live[k]); else live[k]->SetFailed(5 /*EIO*/, std::string("GPU path failed: ") + b2_last_error()); }
            return 0;
        }
        if (rc != B2_OK) return -1;
        for (uint32_t k = 0; k < res.n_runs; k++) {
            Socket* s = live[lo + k];
            const b2_run_status& st = res.runs[k];
            s->AddInputBytes(st.consumed); s->AddInputMessages(st.n_msgs);
            s->set_preferred_index(st.preferred_proto);
            for (uint32_t m = st.first_msg; m < st.first_msg + st.n_msgs; m++) { s->OnMessageCut(12u + res.msgs[m].body_size); Deliver(s, res.msgs[m], res.resp, runs[lo + k]); }
            s->_read_buf.pop_front(st.consumed);              // exactly what the handlers cut (protocol.h:82-92)
            if (st.parse_error != B2_PARSE_ERROR_NOT_ENOUGH_DATA)   // input_messenger.cpp:227-239
                s->SetFailed(22 /*EINVAL*/, std::string("Close socket: ") + ParseErrorToString((ParseError)st.parse_error));
            else if (truncated[lo + k] && st.consumed == 0 && runs[lo + k].length + 32 >= _cap)
                s->SetFailed(22, std::string("Close socket: ") + ParseErrorToString(PARSE_ERROR_TOO_BIG_DATA) + " (one frame exceeds the batch capacity)");
        }
        return (int)res.n_msgs;
    }
public:
    // the host-side fallback of SURVEY §5.3 (a GPU error degrades to the CPU parse for the affected sockets)
    typedef void (*HostParse)(Socket* s);
    void SetHostParse(HostParse p) { _host_parse = p; }

private:
    void Deliver(Socket* s, const b2_msg_desc& d, const uint8_t* resp, const b2_run& run) {
        switch (d.status) {
        case B2_MSG_ECHOED:
        case B2_MSG_ERROR_REPLIED: {
            IOBuf out; out.append(resp + d.resp_off, d.resp_len);
            s->Write(&out);
            break; }
        case B2_MSG_BAD_META:
            s->SetFailed(1003 /*EREQUEST*/, "Fail to parse RpcMeta");         // baidu_rpc_protocol.cpp:577-582
            break;
        case B2_MSG_HOST:
        case B2_MSG_STREAM_FRAME:
        case B2_MSG_UNSUPPORTED: {
            if (!_process) break;
            MostCommonMessage* msg = new MostCommonMessage;   // MostCommonMessage::Get()
            msg->socket = s; msg->desc = d;
            IOBuf whole(s->_read_buf);                        // share the blocks, then trim to the frame's meta / payload
            whole.pop_front(d.frame_off - run.offset + 12);
            whole.cutn(&msg->meta, d.meta_size);
            whole.cutn(&msg->payload, d.body_size - d.meta_size);
            _process(msg);                                    // callee destroys it (DestroyingPtr in the reference)
            break; }
        default: break;                                       // BAD_STREAM_META: frame dropped
        }
    }
    b2_ctx* _ctx = nullptr; uint8_t* _batch = nullptr; size_t _cap; Process _process = nullptr; HostParse _host_parse = nullptr;
    std::unordered_map<uint64_t, std::unique_ptr<Socket>> _sockets;
};

}  // namespace b2
