// b2::IOBuf — the slice of butil::IOBuf's API surface that the message path uses
// (src/butil/iobuf.h:62-369), re-designed for a GPU-staged data path:
//
//   * same observable contract: a non-contiguous, reference-counted queue of
//     BlockRef{offset,length,Block*}; append / cutn / pop_front / pop_back / copy_to /
//     fetch / backing_block / append_user_data / movable() behave as documented there;
//   * same block geometry: DEFAULT_BLOCK_SIZE 8192 with a 32-byte block header, i.e. 8160
//     payload bytes per block (src/butil/iobuf_inl.h:463-493, test/iobuf_unittest.cpp:60-61);
//   * same allocator seam: blocks come from iobuf::blockmem_allocate / blockmem_deallocate
//     (src/butil/iobuf.cpp:168-169) — point them at b2_block_alloc / b2_block_free and every
//     socket read lands in CUDA-pinned memory, exactly how the RDMA pool plugs in
//     (src/brpc/rdma/rdma_helper.cpp:579-582).
//
// Not a copy of butil's implementation: refs live in a std::deque, there is no SmallView /
// BigView split and no TLS block cache; the GPU path batches whole read buffers, so per-ref
// micro-costs are off the hot path.  [NOT thread-safe], like the original.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <sys/uio.h>
#include <unistd.h>
#include <atomic>
#include <deque>
#include <functional>
#include <new>
#include <string>
#include <vector>

namespace b2 {
namespace iobuf {
inline void* default_alloc(size_t n) { return ::operator new(n, std::nothrow); }
inline void default_free(void* p) { ::operator delete(p); }
// the allocator hook (same names as butil::iobuf::)
inline void* (*blockmem_allocate)(size_t) = default_alloc;
inline void (*blockmem_deallocate)(void*) = default_free;
}  // namespace iobuf

class IOBuf {
public:
    static const size_t DEFAULT_BLOCK_SIZE = 8192;
    static const size_t BLOCK_HEADER = 32;
    static const size_t DEFAULT_PAYLOAD = DEFAULT_BLOCK_SIZE - BLOCK_HEADER;   // 8160

    struct Block {
        std::atomic<int> nshared;
        uint32_t size;       // bytes appended so far
        uint32_t cap;        // payload capacity
        char* data;          // == (char*)this + BLOCK_HEADER for pooled blocks, user pointer otherwise
        std::function<void(void*)>* user_deleter;   // non-null for append_user_data blocks
        void inc_ref() { nshared.fetch_add(1, std::memory_order_relaxed); }
        void dec_ref() {
            if (nshared.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                if (user_deleter) { (*user_deleter)(data); delete user_deleter; this->~Block(); ::operator delete(this); }
                else { this->~Block(); iobuf::blockmem_deallocate(this); }
            }
        }
        bool full() const { return size >= cap; }
        size_t left_space() const { return cap - size; }
    };
    struct BlockRef { uint32_t offset; uint32_t length; Block* block; };
    struct Movable { explicit Movable(IOBuf& v) : _v(&v) {} IOBuf& value() const { return *_v; } private: IOBuf* _v; };

    IOBuf() : _nbytes(0) {}
    IOBuf(const IOBuf& rhs) : _refs(rhs._refs), _nbytes(rhs._nbytes) { for (auto& r : _refs) r.block->inc_ref(); }
    IOBuf(const Movable& m) : _nbytes(0) { swap(m.value()); }
    ~IOBuf() { clear(); }
    void operator=(const IOBuf& rhs) { if (this != &rhs) { IOBuf t(rhs); swap(t); } }
    void operator=(const Movable& m) { clear(); swap(m.value()); }
    Movable movable() { return Movable(*this); }
    void swap(IOBuf& o) { _refs.swap(o._refs); std::swap(_nbytes, o._nbytes); }

    bool empty() const { return _nbytes == 0; }
    size_t length() const { return _nbytes; }
    size_t size() const { return _nbytes; }
    void clear() { for (auto& r : _refs) r.block->dec_ref(); _refs.clear(); _nbytes = 0; }

    // Append `count' bytes, copying into the tail block while it has room (and nobody else
    // extended it), then into fresh blocks.  Returns 0 on success, -1 if out of memory.
    int append(const void* data, size_t count) {
        const char* p = static_cast<const char*>(data);
        while (count) {
            Block* b = nullptr;
            if (!_refs.empty()) {
                BlockRef& r = _refs.back();
                // only a block nobody else references is extended in place: a buffer that shares it (a message cut off a read buffer and handed
                // to another thread, a copy) may be appended to as well, and two writers behind the same `size` would collide.  The producer of
                // the block — IOPortal, the only writer of its read blocks — keeps filling it (butil's TLS-block rule, iobuf.cpp:245-330)
                if (!r.block->user_deleter && !r.block->full() && r.offset + r.length == r.block->size && r.block->nshared.load(std::memory_order_acquire) == 1) b = r.block;
            }
            if (!b) {
                b = create_block();
                if (!b) return -1;
                _refs.push_back(BlockRef{0, 0, b});
            }
            const size_t n = count < b->left_space() ? count : b->left_space();
            memcpy(b->data + b->size, p, n);
            b->size += (uint32_t)n; _refs.back().length += (uint32_t)n; _nbytes += n;
            p += n; count -= n;
        }
        return 0;
    }
    int append(const std::string& s) { return append(s.data(), s.size()); }
    int append(const char* s) { return append(s, strlen(s)); }
    int push_back(char c) { return append(&c, 1); }
    // Share the blocks of `other' (no bytes copied).
    void append(const IOBuf& other) { for (auto& r : other._refs) push_ref(r, true); }
    void append(const Movable& m) { IOBuf& o = m.value(); for (auto& r : o._refs) push_ref(r, false); o._refs.clear(); o._nbytes = 0; }
    // Zero-copy reference to caller-owned memory; `deleter(data)' runs when the last ref dies.
    int append_user_data(void* data, size_t size, std::function<void(void*)> deleter) {
        if (size > 0xffffffffull || !size) return -1;
        Block* b = static_cast<Block*>(::operator new(sizeof(Block), std::nothrow));
        if (!b) return -1;
        new (b) Block();
        b->nshared.store(1); b->size = (uint32_t)size; b->cap = (uint32_t)size; b->data = static_cast<char*>(data);
        b->user_deleter = new std::function<void(void*)>(std::move(deleter));
        _refs.push_back(BlockRef{0, (uint32_t)size, b}); _nbytes += size;
        return 0;
    }

    // One refcounted Block over memory this library does not own — the pinned reply area of a GPU batch, a connection's
    // registered read region — and cheap references into it (no allocation per reference: what a zero-copy reply is made of).
    // The creator holds one reference (release_external_block drops it); `deleter(data)' runs when the last one dies.
    static Block* create_external_block(void* data, size_t size, std::function<void(void*)> deleter) {
        if (size > 0xffffffffull) return nullptr;
        Block* b = static_cast<Block*>(::operator new(sizeof(Block), std::nothrow));
        if (!b) return nullptr;
        new (b) Block();
        b->nshared.store(1); b->size = (uint32_t)size; b->cap = (uint32_t)size; b->data = static_cast<char*>(data);
        b->user_deleter = new std::function<void(void*)>(std::move(deleter));
        return b;
    }
    static void release_external_block(Block* b) { if (b) b->dec_ref(); }
    void append_block_range(Block* b, uint32_t offset, uint32_t length) { BlockRef r; r.offset = offset; r.length = length; r.block = b; push_ref(r, true); }

    // (IOBufAsZeroCopyOutputStream::Next) make the free tail of the last block part of the buffer — or start a fresh block — and
    // return where it begins; *len = how many bytes were added
    char* extend_tail(size_t* len) {
        Block* b = nullptr;
        if (!_refs.empty()) { BlockRef& r = _refs.back(); if (!r.block->user_deleter && !r.block->full() && r.offset + r.length == r.block->size && r.block->nshared.load() == 1) b = r.block; }
        if (b) { BlockRef& r = _refs.back(); const size_t n = b->left_space(); char* p = b->data + b->size; b->size += (uint32_t)n; r.length += (uint32_t)n; _nbytes += n; *len = n; return p; }
        b = create_block(); if (!b) return nullptr;
        const size_t n = b->cap; b->size = (uint32_t)n; _refs.push_back(BlockRef{0, (uint32_t)n, b}); _nbytes += n; *len = n;
        return b->data;
    }

    size_t pop_front(size_t n) {
        const size_t saved = n < _nbytes ? n : _nbytes; n = saved;
        while (n) {
            BlockRef& r = _refs.front();
            if (r.length <= n) { n -= r.length; _nbytes -= r.length; r.block->dec_ref(); _refs.pop_front(); }
            else { r.offset += (uint32_t)n; r.length -= (uint32_t)n; _nbytes -= n; n = 0; }
        }
        return saved;
    }
    size_t pop_back(size_t n) {
        const size_t saved = n < _nbytes ? n : _nbytes; n = saved;
        while (n) {
            BlockRef& r = _refs.back();
            if (r.length <= n) { n -= r.length; _nbytes -= r.length; r.block->dec_ref(); _refs.pop_back(); }
            else { r.length -= (uint32_t)n; _nbytes -= n; n = 0; }
        }
        return saved;
    }
    // Cut off n bytes from the front and append them to `out' by reference.
    size_t cutn(IOBuf* out, size_t n) {
        const size_t saved = n < _nbytes ? n : _nbytes; n = saved;
        while (n) {
            BlockRef& r = _refs.front();
            if (r.length <= n) { n -= r.length; _nbytes -= r.length; out->push_ref(r, false); _refs.pop_front(); }
            else {
                BlockRef part{r.offset, (uint32_t)n, r.block};
                out->push_ref(part, true);
                r.offset += (uint32_t)n; r.length -= (uint32_t)n; _nbytes -= n; n = 0;
            }
        }
        return saved;
    }
    size_t cutn(void* out, size_t n) { const size_t c = copy_to(out, n, 0); pop_front(c); return c; }
    size_t cutn(std::string* out, size_t n) { const size_t c = n < _nbytes ? n : _nbytes; const size_t old = out->size(); out->resize(old + c); return cutn(&(*out)[old], c); }
    bool cut1(void* c) { return cutn(c, 1) == 1; }

    size_t copy_to(void* buf, size_t n = (size_t)-1L, size_t pos = 0) const {
        if (pos >= _nbytes) return 0;
        if (n > _nbytes - pos) n = _nbytes - pos;
        char* out = static_cast<char*>(buf); size_t left = n;
        for (auto& r : _refs) {
            if (!left) break;
            if (pos >= r.length) { pos -= r.length; continue; }
            const size_t k = (r.length - pos) < left ? (r.length - pos) : left;
            memcpy(out, r.block->data + r.offset + pos, k);
            out += k; left -= k; pos = 0;
        }
        return n;
    }
    size_t copy_to(std::string* s, size_t n = (size_t)-1L, size_t pos = 0) const {
        if (pos >= _nbytes) { s->clear(); return 0; }
        if (n > _nbytes - pos) n = _nbytes - pos;
        s->resize(n); return copy_to(&(*s)[0], n, pos);
    }
    size_t copy_to(IOBuf* buf, size_t n = (size_t)-1L, size_t pos = 0) const {
        IOBuf t(*this); t.pop_front(pos); IOBuf cut; const size_t c = t.cutn(&cut, n); *buf = cut.movable(); return c;
    }
    std::string to_string() const { std::string s; copy_to(&s); return s; }
    bool equals(const std::string& s) const { return s.size() == _nbytes && to_string() == s; }
    // Contiguous view of the first n bytes: a pointer into the first block when it holds them,
    // else `aux_buffer' filled with a copy.  NULL if fewer than n bytes.
    const void* fetch(void* aux_buffer, size_t n) const {
        if (n > _nbytes) return nullptr;
        if (!n) return aux_buffer;
        const BlockRef& r = _refs.front();
        if (r.length >= n) return r.block->data + r.offset;
        copy_to(aux_buffer, n, 0); return aux_buffer;
    }
    const void* fetch1() const { return _refs.empty() ? nullptr : _refs.front().block->data + _refs.front().offset; }

    size_t backing_block_num() const { return _refs.size(); }
    // (data, size) of the i-th backing block; (NULL, 0) if there is no such block
    std::pair<const char*, size_t> backing_block(size_t i) const {
        if (i >= _refs.size()) return {nullptr, 0};
        return {_refs[i].block->data + _refs[i].offset, _refs[i].length};
    }
    int block_nshared(size_t i) const { return i < _refs.size() ? _refs[i].block->nshared.load() : -1; }

    // Write the front of this buffer to `fd' with ONE writev over its block refs (at most IOV_MAX_REFS of them, stopping once
    // `size_hint' bytes are covered) and pop what was written — IOBuf::cut_into_file_descriptor, src/butil/iobuf.cpp:827-857.
    // Returns writev's result.
    static const size_t IOV_MAX_REFS = 256;
    ssize_t cut_into_file_descriptor(int fd, size_t size_hint = 1024 * 1024) {
        if (empty()) return 0;
        struct iovec vec[IOV_MAX_REFS];
        size_t nvec = 0, covered = 0;
        for (auto& r : _refs) {
            if (nvec >= IOV_MAX_REFS || covered >= size_hint) break;
            vec[nvec].iov_base = r.block->data + r.offset; vec[nvec].iov_len = r.length;
            covered += r.length; nvec++;
        }
        const ssize_t nw = ::writev(fd, vec, (int)nvec);
        if (nw > 0) pop_front((size_t)nw);
        return nw;
    }
    // Several buffers (the replies queued on one socket) in one writev — what KeepWrite/DoWrite do with up to 256 IOBufs
    // (src/brpc/socket.cpp:1780-1889, iobuf.cpp:954-992); the written prefix is popped piece by piece.
    static ssize_t cut_multiple_into_file_descriptor(int fd, IOBuf* const* pieces, size_t count) {
        if (count == 0) return 0;
        struct iovec vec[IOV_MAX_REFS];
        size_t nvec = 0;
        for (size_t i = 0; i < count && nvec < IOV_MAX_REFS; i++)
            for (auto& r : pieces[i]->_refs) {
                if (nvec >= IOV_MAX_REFS) break;
                vec[nvec].iov_base = r.block->data + r.offset; vec[nvec].iov_len = r.length; nvec++;
            }
        if (nvec == 0) return 0;
        const ssize_t nw = ::writev(fd, vec, (int)nvec);
        if (nw <= 0) return nw;
        size_t left = (size_t)nw;
        for (size_t i = 0; i < count && left; i++) left -= pieces[i]->pop_front(left);
        return nw;
    }

protected:
    static Block* create_block() {
        void* mem = iobuf::blockmem_allocate(DEFAULT_BLOCK_SIZE);
        if (!mem) return nullptr;
        static_assert(sizeof(Block) <= BLOCK_HEADER, "block header is 32 bytes");
        Block* b = new (mem) Block();
        b->nshared.store(1); b->size = 0; b->cap = (uint32_t)DEFAULT_PAYLOAD; b->data = static_cast<char*>(mem) + BLOCK_HEADER; b->user_deleter = nullptr;
        return b;
    }
    // Append a ref, merging with the tail when it continues the same block (as IOBuf::_push_back_ref does).
    void push_ref(const BlockRef& r, bool add_ref) {
        if (!r.length) { if (!add_ref) r.block->dec_ref(); return; }
        if (!_refs.empty()) {
            BlockRef& t = _refs.back();
            if (t.block == r.block && t.offset + t.length == r.offset) {
                t.length += r.length; _nbytes += r.length;
                if (!add_ref) r.block->dec_ref();
                return;
            }
        }
        if (add_ref) r.block->inc_ref();
        _refs.push_back(r); _nbytes += r.length;
    }
    std::deque<BlockRef> _refs;
    size_t _nbytes;
};

inline void swap(IOBuf& a, IOBuf& b) { a.swap(b); }

// IOPortal — an IOBuf that reads from a file descriptor straight into blocks (src/butil/iobuf.h:445-492,
// IOPortal::pappend_from_file_descriptor iobuf.cpp:1481-1541): one readv over the free tail of the last block plus fresh
// blocks, at most MAX_APPEND_IOVEC of them or `max_count' bytes; blocks that stayed empty are kept for the next read.
// With blockmem_allocate pointed at b2_block_alloc the bytes land in CUDA-pinned memory (Socket::DoRead, socket.cpp:2042-2122).
class IOPortal : public IOBuf {
public:
    static const int MAX_APPEND_IOVEC = 64;
    IOPortal() {}
    ~IOPortal() { return_cached_blocks(); }
    IOPortal(const IOPortal&) = delete;
    void operator=(const IOPortal&) = delete;

    ssize_t append_from_file_descriptor(int fd, size_t max_count) {
        struct iovec vec[MAX_APPEND_IOVEC];
        Block* blk[MAX_APPEND_IOVEC];
        int nvec = 0; size_t space = 0;
        // the tail block of the buffer, when our last ref ends where the block's data ends
        if (!_refs.empty()) {
            BlockRef& r = _refs.back();
            if (!r.block->user_deleter && !r.block->full() && r.offset + r.length == r.block->size) {   // (others may share earlier bytes of the block)
                blk[nvec] = r.block; vec[nvec].iov_base = r.block->data + r.block->size;
                vec[nvec].iov_len = r.block->left_space() < max_count ? r.block->left_space() : max_count;
                space += vec[nvec].iov_len; nvec++;
            }
        }
        const int first_fresh = nvec;
        while (space < max_count && nvec < MAX_APPEND_IOVEC) {
            Block* b;
            if (!_cached.empty()) { b = _cached.back(); _cached.pop_back(); }
            else { b = create_block(); if (!b) { if (nvec == 0) { errno = ENOMEM; return -1; } break; } }
            blk[nvec] = b; vec[nvec].iov_base = b->data;
            vec[nvec].iov_len = b->cap < max_count - space ? b->cap : max_count - space;
            space += vec[nvec].iov_len; nvec++;
        }
        const ssize_t nr = ::readv(fd, vec, nvec);
        size_t left = nr > 0 ? (size_t)nr : 0;
        for (int i = 0; i < nvec; i++) {
            const size_t got = left < vec[i].iov_len ? left : vec[i].iov_len;
            left -= got;
            if (i < first_fresh) {                                  // extended the existing tail ref
                if (got) { blk[i]->size += (uint32_t)got; _refs.back().length += (uint32_t)got; _nbytes += got; }
            } else if (got) {
                blk[i]->size = (uint32_t)got;
                _refs.push_back(BlockRef{0, (uint32_t)got, blk[i]}); _nbytes += got;   // the block's creation reference becomes the buffer's
            } else _cached.push_back(blk[i]);                       // untouched: keep it for the next read
        }
        return nr;
    }
    void return_cached_blocks() { for (Block* b : _cached) b->dec_ref(); _cached.clear(); }
    size_t cached_block_num() const { return _cached.size(); }

private:
    std::vector<Block*> _cached;
};


// ---- the adapters every function of the path reads / writes IOBufs through --------------------------------
// IOBufAsZeroCopyInputStream / IOBufAsZeroCopyOutputStream (src/butil/iobuf.h:559-607, iobuf.cpp:1826-1875,1877-2011): the
// google::protobuf::io::ZeroCopy{Input,Output}Stream contract (Next / BackUp / Skip / ByteCount) over block references — what
// ParsePbFromIOBuf (protocol.cpp:202-239) and SerializeRpcHeaderAndMeta (baidu_rpc_protocol.cpp:83-103) hand to libprotobuf.
// Here they are plain classes with the same members (protobuf is not linked in this repo).
class IOBufAsZeroCopyInputStream {
public:
    explicit IOBufAsZeroCopyInputStream(const IOBuf& buf) : _ref_index(0), _add_offset(0), _byte_count(0), _buf(&buf) {}
    bool Next(const void** data, int* size) {
        const std::pair<const char*, size_t> b = _buf->backing_block((size_t)_ref_index);
        if (b.first == nullptr) return false;
        *data = b.first + _add_offset; *size = (int)(b.second - (size_t)_add_offset);
        _byte_count += (int64_t)(b.second - (size_t)_add_offset);
        ++_ref_index; _add_offset = 0;
        return true;
    }
    void BackUp(int count) {                           // only right after a Next: `count` bytes of its block were not used
        if (_ref_index > 0) {
            const std::pair<const char*, size_t> b = _buf->backing_block((size_t)--_ref_index);
            _add_offset = (int)b.second - count; _byte_count -= count;
        }
    }
    bool Skip(int count) {
        for (;;) {
            const std::pair<const char*, size_t> b = _buf->backing_block((size_t)_ref_index);
            if (b.first == nullptr) return count == 0;
            const int left = (int)b.second - _add_offset;
            if (count < left) { _add_offset += count; _byte_count += count; return true; }
            count -= left; _byte_count += left; ++_ref_index; _add_offset = 0;
            if (count == 0) return true;
        }
    }
    int64_t ByteCount() const { return _byte_count; }
private:
    int _ref_index, _add_offset; int64_t _byte_count; const IOBuf* _buf;
};

class IOBufAsZeroCopyOutputStream {
public:
    explicit IOBufAsZeroCopyOutputStream(IOBuf* buf) : _buf(buf), _byte_count(0), _last(0) {}
    // hands out the free tail of the buffer's last block (a fresh block when it is full); the bytes count as written until BackUp
    bool Next(void** data, int* size) {
        const size_t before = _buf->length();
        char* p = _buf->extend_tail(&_last);
        if (!p) return false;
        *data = p; *size = (int)_last; _byte_count += (int64_t)(_buf->length() - before);
        return true;
    }
    void BackUp(int count) {                           // `count' can be as long as ByteCount() (iobuf.h:597)
        _buf->pop_back((size_t)count); _byte_count -= count;
    }
    int64_t ByteCount() const { return _byte_count; }
private:
    IOBuf* _buf; int64_t _byte_count; size_t _last;
};

// IOBufCutter (src/butil/iobuf.h:509-556, iobuf_inl.h:463-578): cuts from the front with the current block's span cached, the
// fast way to parse many small fields off one buffer.
class IOBufCutter {
public:
    explicit IOBufCutter(IOBuf* buf) : _buf(buf) {}
    size_t cutn(IOBuf* out, size_t n) { return _buf->cutn(out, n); }
    size_t cutn(std::string* out, size_t n) { return _buf->cutn(out, n); }
    size_t cutn(void* out, size_t n) { return _buf->cutn(out, n); }
    bool cut1(void* data) { return _buf->cutn(data, 1) == 1; }
    size_t copy_to(void* data, size_t n) { return _buf->copy_to(data, n, 0); }
    const void* fetch1() { const std::pair<const char*, size_t> b = _buf->backing_block(0); return b.first && b.second ? b.first : nullptr; }
    size_t pop_front(size_t n) { return _buf->pop_front(n); }
    size_t remaining_bytes() const { return _buf->length(); }
private:
    IOBuf* _buf;
};

// IOBufAppender (src/butil/iobuf.h:684-724): appends through a cached span of the tail block; buf() / move_to() give the bytes back.
class IOBufAppender {
public:
    IOBufAppender() : _data(nullptr), _data_end(nullptr), _zc_stream(&_buf) {}
    int append(const void* src, size_t n) {
        const char* s = static_cast<const char*>(src);
        while (n) {
            if (_data == _data_end && add_block() != 0) return -1;
            const size_t k = n < (size_t)(_data_end - _data) ? n : (size_t)(_data_end - _data);
            memcpy(_data, s, k); _data += k; s += k; n -= k;
        }
        return 0;
    }
    int append(const std::string& s) { return append(s.data(), s.size()); }
    int append_decimal(long d) { char tmp[24]; const int n = snprintf(tmp, sizeof tmp, "%ld", d); return append(tmp, (size_t)n); }
    int push_back(char c) { if (_data == _data_end && add_block() != 0) return -1; *_data++ = c; return 0; }
    IOBuf& buf() { shrink(); return _buf; }
    void move_to(IOBuf& target) { IOBuf& b = buf(); target.clear(); target.swap(b); }
private:
    void shrink() { const size_t unused = (size_t)(_data_end - _data); if (unused) { _zc_stream.BackUp((int)unused); _data = _data_end = nullptr; } }
    int add_block() { void* p = nullptr; int n = 0; if (!_zc_stream.Next(&p, &n)) return -1; _data = static_cast<char*>(p); _data_end = _data + n; return 0; }
    char* _data; char* _data_end; IOBuf _buf; IOBufAsZeroCopyOutputStream _zc_stream;
};

// Forward byte iterator over an IOBuf that does not modify it (butil::IOBufBytesIterator, src/butil/iobuf.h:688-727) —
// what the reference's h2 parser walks frames with.
class IOBufBytesIterator {
public:
    explicit IOBufBytesIterator(const IOBuf& buf) : _buf(&buf), _block(0), _off(0), _left(buf.length()) { settle(); }
    uint8_t operator*() const { return (uint8_t)_buf->backing_block(_block).first[_off]; }
    operator const void*() const { return _left ? this : nullptr; }
    void operator++() { if (_left) { _left--; _off++; settle(); } }
    size_t bytes_left() const { return _left; }
    size_t forward(size_t n) { const size_t k = n < _left ? n : _left; size_t m = k; while (m) { const size_t in = _buf->backing_block(_block).second - _off; const size_t s = in < m ? in : m; _off += s; _left -= s; m -= s; settle(); } return k; }
    size_t copy_and_forward(void* out, size_t n) {
        uint8_t* o = static_cast<uint8_t*>(out); const size_t k = n < _left ? n : _left; size_t m = k;
        while (m) { auto b = _buf->backing_block(_block); const size_t s = (b.second - _off) < m ? (b.second - _off) : m; memcpy(o, b.first + _off, s); o += s; _off += s; _left -= s; m -= s; settle(); }
        return k;
    }
private:
    void settle() { while (_left && _off >= _buf->backing_block(_block).second) { _block++; _off = 0; } }
    const IOBuf* _buf; size_t _block, _off, _left;
};

}  // namespace b2
