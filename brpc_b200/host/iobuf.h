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
#include <atomic>
#include <deque>
#include <functional>
#include <new>
#include <string>

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
                if (!r.block->user_deleter && !r.block->full() && r.offset + r.length == r.block->size) b = r.block;
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

private:
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

}  // namespace b2
