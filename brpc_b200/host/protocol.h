// b2::Protocol — seam (1) of the drop-in boundary: the table brpc registers per wire protocol, field for field
//   struct Protocol                       src/brpc/protocol.h:77-172
//   RegisterProtocol / FindProtocol       src/brpc/protocol.cpp:77-117 (once-only per type, type < 128: protocol.cpp:62)
//   ParseResult, MakeParseError, MakeMessage   src/brpc/parse_result.h:25-74
// and a `parse` callback served by the GPU path: b2::policy::ParseOnGpu keeps brpc's per-call contract (cut exactly one message
// off `source`, or NOT_ENOUGH_DATA with `source` untouched, or TRY_OTHERS after popping what a handler popped, or a fatal
// error) while the work behind it is a batch: the first call on a socket whose queue is empty sends the socket's pending
// bytes through the C ABI (cut loop over every enabled handler, RpcMeta decode, echo service, reply packing) and every
// call hands out the next cut message.  Types brpc would pass that this path never looks into (Controller,
// google::protobuf::Message, MethodDescriptor, Authenticator, SocketMessage, EndPoint) are opaque forward declarations.
#pragma once
#include <deque>
#include <unordered_map>
#include "input_messenger.h"

namespace google { namespace protobuf { class Message; class MethodDescriptor; } }
namespace b2 {
class Controller; class Authenticator; class SocketMessage; struct EndPoint;

class ParseResult {                                   // parse_result.h:40-74
public:
    explicit ParseResult(ParseError err) : _msg(nullptr), _err(err), _user_desc(nullptr) {}
    explicit ParseResult(InputMessageBase* msg) : _msg(msg), _err(PARSE_OK), _user_desc(nullptr) {}
    ParseResult(ParseError err, const char* user_desc) : _msg(nullptr), _err(err), _user_desc(user_desc) {}
    bool is_ok() const { return error() == PARSE_OK; }
    ParseError error() const { return _err; }
    const char* error_str() const { return _user_desc ? _user_desc : ParseErrorToString(_err); }
    InputMessageBase* message() const { return _msg; }
private:
    InputMessageBase* _msg; ParseError _err; const char* _user_desc;
};
inline ParseResult MakeParseError(ParseError err) { return ParseResult(err); }
inline ParseResult MakeParseError(ParseError err, const char* user_desc) { return ParseResult(err, user_desc); }
inline ParseResult MakeMessage(InputMessageBase* msg) { return ParseResult(msg); }

enum ProtocolType {                                   // options.proto:38-67 (the ones this path frames) + a free slot for the GPU entry
    PROTOCOL_UNKNOWN = 0, PROTOCOL_BAIDU_STD = 1, PROTOCOL_STREAMING_RPC = 2, PROTOCOL_HULU_PBRPC = 3, PROTOCOL_SOFA_PBRPC = 4,
    PROTOCOL_NSHEAD = 12, PROTOCOL_B2_GPU = 100,
};
enum ConnectionType { CONNECTION_TYPE_UNKNOWN = 0, CONNECTION_TYPE_SINGLE = 1, CONNECTION_TYPE_POOLED = 2, CONNECTION_TYPE_SHORT = 4, CONNECTION_TYPE_ALL = 7 };   // options.proto:90-97 + protocol.h:174-180

struct Protocol {                                     // protocol.h:77-172, same fields, same order
    typedef ParseResult (*Parse)(IOBuf* source, Socket* socket, bool read_eof, const void* arg);
    Parse parse;
    typedef void (*SerializeRequest)(IOBuf* request_buf, Controller* cntl, const google::protobuf::Message* request);
    SerializeRequest serialize_request;
    typedef void (*PackRequest)(IOBuf* iobuf_out, SocketMessage** user_message_out, uint64_t correlation_id,
                                const google::protobuf::MethodDescriptor* method, Controller* controller, const IOBuf& request_buf, const Authenticator* auth);
    PackRequest pack_request;
    typedef void (*ProcessRequest)(InputMessageBase* msg);
    ProcessRequest process_request;
    typedef void (*ProcessResponse)(InputMessageBase* msg);
    ProcessResponse process_response;
    typedef bool (*Verify)(const InputMessageBase* msg);
    Verify verify;
    typedef bool (*ParseServerAddress)(EndPoint* out, const char* server_addr_and_port);
    ParseServerAddress parse_server_address;
    typedef const std::string& (*GetMethodName)(const google::protobuf::MethodDescriptor* method, const Controller*);
    GetMethodName get_method_name;
    ConnectionType supported_connection_type;
    const char* name;
    bool support_client() const { return serialize_request && pack_request && process_response; }
    bool support_server() const { return process_request; }
};

// protocol.cpp:62-117: a fixed table of MAX_PROTOCOL_SIZE entries, registration is once-only per type
const int MAX_PROTOCOL_SIZE = 128;
struct ProtocolEntry { bool valid; Protocol protocol; };
inline ProtocolEntry* protocol_map() { static ProtocolEntry m[MAX_PROTOCOL_SIZE]; return m; }
inline int RegisterProtocol(ProtocolType type, const Protocol& protocol) {
    const size_t index = (size_t)type;
    if (index >= (size_t)MAX_PROTOCOL_SIZE) return -1;                  // "ProtocolType=... is out of range"
    if (!protocol.support_client() && !protocol.support_server()) return -1;
    if (protocol_map()[index].valid) return -1;                         // "ProtocolType=... was registered"
    protocol_map()[index].protocol = protocol; protocol_map()[index].valid = true;
    return 0;
}
inline const Protocol* FindProtocol(ProtocolType type) {
    const size_t index = (size_t)type;
    return index < (size_t)MAX_PROTOCOL_SIZE && protocol_map()[index].valid ? &protocol_map()[index].protocol : nullptr;
}

namespace policy {

// the `arg` of the GPU entry (InputMessageHandler::arg, input_messenger.h:59-63): a context plus, per socket, the messages of
// the last batch that parse() has not handed out yet
struct GpuParser {
    b2_ctx* ctx; uint8_t* stage; size_t cap;
    struct Pending { std::deque<b2_msg_desc> descs; std::vector<uint8_t> replies; size_t cursor = 0; uint32_t tail_error = B2_PARSE_ERROR_NOT_ENOUGH_DATA, consumed = 0; bool have = false; };
    std::unordered_map<Socket*, Pending> pending;
    GpuParser(b2_ctx* c, size_t stage_bytes) : ctx(c), stage(static_cast<uint8_t*>(b2_block_alloc(stage_bytes))), cap(stage_bytes) {}
    ~GpuParser() { b2_block_free(stage); }
};
// a cut message of the GPU entry: MostCommonMessage + the reply the device already packed for it (empty for host-handled ones)
struct GpuMessage : public MostCommonMessage { IOBuf reply; };

inline ParseResult ParseOnGpu(IOBuf* source, Socket* socket, bool /*read_eof*/, const void* arg) {
    GpuParser* gp = const_cast<GpuParser*>(static_cast<const GpuParser*>(arg));
    GpuParser::Pending& q = gp->pending[socket];
    if (!q.have) {
        const size_t n = source->length() < gp->cap - 64 ? source->length() : gp->cap - 64;
        if (n == 0) return MakeParseError(PARSE_ERROR_NOT_ENOUGH_DATA);
        source->copy_to(gp->stage, n, 0);
        b2_run run; run.socket_id = socket->id(); run.offset = 0; run.length = (uint32_t)n; run.preferred_proto = socket->preferred_index(); run.flags = 0;
        b2_batch_result res;
        if (b2_process_batch(gp->ctx, gp->stage, (uint32_t)n, &run, 1, &res) != B2_OK) return MakeParseError(PARSE_ERROR_NO_RESOURCE, b2_last_error());
        q.descs.assign(res.msgs, res.msgs + res.n_msgs);
        q.replies.clear();
        for (auto& d : q.descs) {                                       // keep the packed replies: the ABI's buffers live until its next call
            const uint32_t at = (uint32_t)q.replies.size();
            if ((d.status == B2_MSG_ECHOED || d.status == B2_MSG_ERROR_REPLIED) && d.resp_len) q.replies.insert(q.replies.end(), res.resp + d.resp_off, res.resp + d.resp_off + d.resp_len);
            d.resp_off = at;
        }
        q.cursor = 0; q.consumed = res.runs[0].consumed; q.tail_error = res.runs[0].parse_error; q.have = true;
        socket->set_preferred_index(res.runs[0].preferred_proto);
    }
    if (q.descs.empty()) {
        // what ends the cut loop: bytes a handler popped before answering TRY_OTHERS leave `source` too (protocol.h:82-92)
        const ParseError e = (ParseError)q.tail_error;
        if (q.consumed > q.cursor) source->pop_front(q.consumed - q.cursor);
        q.have = false; q.cursor = 0;
        return MakeParseError(e);
    }
    const b2_msg_desc d = q.descs.front(); q.descs.pop_front();
    const uint32_t hdr = d.protocol == B2_PROTOCOL_SOFA_PBRPC ? 24u : d.protocol == B2_PROTOCOL_NSHEAD ? 36u : 12u;
    if (d.frame_off > q.cursor) source->pop_front(d.frame_off - q.cursor);      // garbage a handler popped in front of this message
    GpuMessage* msg = new GpuMessage;                                   // MostCommonMessage::Get()
    msg->socket = socket; msg->desc = d;
    if (d.protocol == B2_PROTOCOL_NSHEAD) source->cutn(&msg->meta, hdr);       // ParseNsheadMessage: meta = the nshead itself (:178-179)
    else { source->pop_front(hdr); source->cutn(&msg->meta, d.meta_size); }
    source->cutn(&msg->payload, d.body_size - (d.protocol == B2_PROTOCOL_NSHEAD ? 0u : d.meta_size));
    if ((d.status == B2_MSG_ECHOED || d.status == B2_MSG_ERROR_REPLIED) && d.resp_len) msg->reply.append(q.replies.data() + d.resp_off, d.resp_len);
    q.cursor = d.frame_off + hdr + d.body_size;
    return MakeMessage(msg);
}
// process_request of the GPU entry: the device already ran the echo service and packed the reply — write it (SendRpcResponse's
// Socket::Write, baidu_rpc_protocol.cpp:441-456); anything else goes to the messenger's host callback
inline void ProcessOnGpu(InputMessageBase* base) {
    GpuMessage* m = static_cast<GpuMessage*>(base);
    if (!m->reply.empty()) m->socket->Write(&m->reply);
    delete m;                                                           // msg->Destroy()
}
inline Protocol GpuProtocol() {
    Protocol p = { ParseOnGpu, nullptr, nullptr, ProcessOnGpu, nullptr, nullptr, nullptr, nullptr, CONNECTION_TYPE_ALL, "b2_gpu" };
    return p;
}
}  // namespace policy
}  // namespace b2
