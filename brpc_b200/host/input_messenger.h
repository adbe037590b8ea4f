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
#include <stdint.h>
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
    void SetFailed(int error_code, const std::string& text) { if (!_failed) { _failed = true; _error_code = error_code; _error_text = text; } }
    int Write(IOBuf* data) { if (_failed) return -1; _write_buf.append(data->movable()); return 0; }
    void AddInputBytes(size_t n) { _in_bytes += n; }
    void AddInputMessages(size_t n) { _in_msgs += n; }
    uint64_t in_bytes() const { return _in_bytes; }
    uint64_t in_msgs() const { return _in_msgs; }
private:
    uint64_t _id; int _preferred_index = -1; bool _failed = false; int _error_code = 0; std::string _error_text;
    uint64_t _in_bytes = 0, _in_msgs = 0; size_t _avg_msg_size = 0;
};

class GpuInputMessenger {
public:
    typedef void (*Process)(InputMessageBase* msg);   // InputMessageHandler::Process, input_messenger.h:51-57

    explicit GpuInputMessenger(const b2_options& opt) : _cap(opt.max_batch_bytes) {
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
    int ProcessNewMessages(const std::vector<Socket*>& readable) {
        std::vector<b2_run> runs; std::vector<Socket*> live;
        size_t total = 0;
        for (Socket* s : readable) {
            if (s->Failed() || s->_read_buf.empty()) continue;
            const size_t n = s->_read_buf.length();
            if (total + n + 16 > _cap) break;                 // the rest waits for the next round
            s->_read_buf.copy_to(_batch + total, n, 0);       // gather the (pinned) blocks into the batch buffer
            b2_run r; r.socket_id = s->id(); r.offset = (uint32_t)total; r.length = (uint32_t)n;
            r.preferred_proto = s->preferred_index(); r.flags = 0;
            runs.push_back(r); live.push_back(s);
            total = (total + n + 15) & ~(size_t)15;
        }
        if (runs.empty()) return 0;
        b2_batch_result res;
        if (b2_process_batch(_ctx, _batch, (uint32_t)total, runs.data(), (uint32_t)runs.size(), &res) != B2_OK) return -1;
        for (uint32_t i = 0; i < res.n_runs; i++) {
            Socket* s = live[i];
            const b2_run_status& st = res.runs[i];
            s->AddInputBytes(st.consumed); s->AddInputMessages(st.n_msgs);
            s->set_preferred_index(st.preferred_proto);
            for (uint32_t m = st.first_msg; m < st.first_msg + st.n_msgs; m++) { s->OnMessageCut(12u + res.msgs[m].body_size); Deliver(s, res.msgs[m], res.resp, runs[i]); }
            s->_read_buf.pop_front(st.consumed);              // exactly what the handlers cut (protocol.h:82-92)
            if (st.parse_error != B2_PARSE_ERROR_NOT_ENOUGH_DATA)   // input_messenger.cpp:227-239
                s->SetFailed(22 /*EINVAL*/, std::string("Close socket: ") + ParseErrorToString((ParseError)st.parse_error));
        }
        return (int)res.n_msgs;
    }

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
    b2_ctx* _ctx = nullptr; uint8_t* _batch = nullptr; size_t _cap; Process _process = nullptr;
    std::unordered_map<uint64_t, std::unique_ptr<Socket>> _sockets;
};

}  // namespace b2
