"""Writes tests/cpp/emul_lib.cc (+ emul_kernels.cuh, emul_h2.cuh): the product library's own sources — brpc_b200/csrc/b2_api.cu, b2_kernels.cuh,
b2_h2.cuh — rewritten just enough for g++ on top of tests/cpp/cuda_emul.h:
  * kernel<<<grid, block, smem, stream>>>(args);  ->  be_launch(grid, block, smem, stream, [&] { kernel(args); });
  * `extern __shared__ T name[];`                 ->  a pointer to the launch's dynamic shared memory block
  * the inline PTX: the 16-byte cp.async and the TMA bulk copies become memcpy at issue time, mbarrier / commit / wait / fence statements
    disappear, the snappy ring's ld/st.shared become array accesses, the system-scope load / store of the ring doorbell become atomics,
    the L2 prefetch becomes a READ of the range (so that a prefetch beyond the buffer is an ASan report)
Nothing else is touched: the rest is the product source as it stands.  Test infrastructure only."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "brpc_b200", "csrc")


def replace_function(src, signature_regex, new_text):
    """the top-level (or namespace-level) function whose definition starts at `signature_regex`, up to its closing brace at column 0 or the
    end of a one-line body"""
    m = re.search(signature_regex, src, re.M)
    assert m, signature_regex
    line_end = src.index("\n", m.start())
    if src[m.start():line_end].rstrip().endswith("}"):
        end = line_end
    else:
        end = src.index("\n}", m.start()) + 2
    return src[:m.start()] + new_text + src[end:]


def device_source(name):
    src = open(os.path.join(CSRC, name)).read()
    if name == "b2_kernels.cuh":
        n0 = src.count('asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"')
        src = src.replace('asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");', 'be_cp16((void*)dst, (const void*)(src));')
        assert src.count("be_cp16((void*)dst, (const void*)(src));") == n0 == 2
        src = src.replace("const uint32_t dst = (uint32_t)__cvta_generic_to_shared(", "const size_t dst = (size_t)__cvta_generic_to_shared(")
        n1 = src.count('asm volatile("cp.async.wait_group 0;" ::: "memory");')
        src = src.replace('asm volatile("cp.async.wait_group 0;" ::: "memory");', "be_cp16_wait();")
        assert n1 == 2
        # the snappy decoder's window in shared memory: a 32-bit shared address on the device, a pointer here
        src = replace_function(src, r"^__device__ __forceinline__ uint32_t ring_ld\(", "__device__ __forceinline__ uint32_t ring_ld(const uint8_t* ring_s, uint32_t p) { return ring_s[p & (kSnapRing - 1)]; }")
        src = replace_function(src, r"^__device__ __forceinline__ void ring_st\(", "__device__ __forceinline__ void ring_st(uint8_t* ring_s, uint32_t p, uint32_t v) { ring_s[p & (kSnapRing - 1)] = (uint8_t)v; }")
        old = "const uint32_t ring_s = use_ring ? (uint32_t)__cvta_generic_to_shared(ring) : 0u;"
        assert src.count(old) == 1
        src = src.replace(old, "uint8_t* const ring_s = ring;")
        # TMA + mbarrier: the copy happens when it is issued, the wait finds it done
        src = replace_function(src, r"^__device__ __forceinline__ void mbar_init\(", "__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t) { *bar = 0; }")
        src = replace_function(src, r"^__device__ __forceinline__ void mbar_arrive_expect_tx\(", "__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long*, uint32_t) {}")
        src = replace_function(src, r"^__device__ __forceinline__ void mbar_wait\(", "__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t) { be_mbar_wait(bar); }")
        src = replace_function(src, r"^__device__ __forceinline__ void bulk_g2s\(", "__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) { be_bulk_g2s(sdst, gsrc, bytes, bar); }")
        src = replace_function(src, r"^__device__ __forceinline__ void bulk_s2g\(", "__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) { be_bulk_s2g(gdst, ssrc, bytes); }")
        src = replace_function(src, r"^__device__ __forceinline__ void bulk_prefetch_l2\(", "__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) { be_prefetch(gsrc, bytes); }")
        src = replace_function(src, r"^__device__ __forceinline__ void bulk_commit\(", "__device__ __forceinline__ void bulk_commit() { be_bulk_commit(); }")
        src = replace_function(src, r"^template <int N> __device__ __forceinline__ void bulk_wait_read\(", "template <int N> __device__ __forceinline__ void bulk_wait_read() { be_bulk_retire(N); }")
        src = replace_function(src, r"^template <int N> __device__ __forceinline__ void bulk_wait\(", "template <int N> __device__ __forceinline__ void bulk_wait() { be_bulk_retire(N); }")
        # the ring's doorbell words in mapped host memory, the device clock
        src = replace_function(src, r"^__device__ __forceinline__ uint32_t ld_sys_u32\(", "__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t* p) { return __atomic_load_n((const uint32_t*)p, __ATOMIC_ACQUIRE); }")
        src = replace_function(src, r"^__device__ __forceinline__ void st_sys_u32\(", "__device__ __forceinline__ void st_sys_u32(volatile uint32_t* p, uint32_t v) { __atomic_store_n((uint32_t*)p, v, __ATOMIC_RELEASE); }")
        src = replace_function(src, r"^__device__ __forceinline__ unsigned long long globaltimer_ns\(", "__device__ __forceinline__ unsigned long long globaltimer_ns() { return be_now_ns(); }")
    src = re.sub(r"asm volatile\((?:.|\n)*?\);", ";", src)
    assert "asm" not in re.sub(r"//.*", "", src), "inline PTX left in " + name
    # dynamic shared memory: the block the launch allocated
    src = re.sub(r"extern __shared__ (?:__align__\(\d+\) )?(\w+) (\w+)\[\];", r"\1* const \2 = reinterpret_cast<\1*>(be_cur->dyn);", src)
    assert "extern __shared__" not in src
    src = src.replace("#include <cuda_runtime.h>", "")
    for inc in ("b2_core.cuh", "b2_inflate.cuh", "b2_hpack_tables.cuh"):
        src = src.replace('#include "%s"' % inc, '#include "../../brpc_b200/csrc/%s"' % inc)
    src = src.replace('#include "b2_kernels.cuh"', '#include "emul_kernels.cuh"')
    return "// GENERATED by tests/cpp/gen_emul_lib.py from brpc_b200/csrc/%s - do not edit\n" % name + src


def split_top(s):
    """split at commas outside parentheses / brackets / template arguments"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    out, pos, n = "", 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            break
        m = re.search(r"([A-Za-z_]\w*(?:<[^<>;]*>)?)\s*$", src[pos:k])        # the kernel (with its template arguments) right before <<<
        assert m, src[k - 80:k + 40]
        name_at = pos + m.start(1)
        e = src.index(">>>", k)
        cfg = split_top(src[k + 3:e])
        assert len(cfg) == 4, cfg
        a = src.index("(", e)
        assert src[e + 3:a].strip() == ""
        depth, j = 0, a
        while True:
            if src[j] == "(":
                depth += 1
            elif src[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        args = src[a + 1:j]
        assert src[j + 1] == ";", src[j - 40:j + 10]
        persistent = "ring_stream" in cfg[3]
        out += src[pos:name_at]
        out += "%s(\"%s\", (unsigned)(%s), (unsigned)(%s), (size_t)(%s), %s, [%s] { %s(%s); });" % (
            "be_launch_async" if persistent else "be_launch", m.group(1), cfg[0], cfg[1], cfg[2], cfg[3], "=" if persistent else "&", m.group(1), args)
        pos = j + 2
        n += 1
    return out + src[pos:], n


def main():
    open(os.path.join(HERE, "emul_kernels.cuh"), "w").write(device_source("b2_kernels.cuh"))
    open(os.path.join(HERE, "emul_h2.cuh"), "w").write(device_source("b2_h2.cuh"))
    api = open(os.path.join(CSRC, "b2_api.cu")).read()
    api, n = rewrite_launches(api)
    assert n >= 29 and "<<<" not in api
    api = api.replace("#include <cuda_runtime.h>", '#include "cuda_emul.h"')
    api = api.replace('#include "b2_kernels.cuh"', '#include "emul_kernels.cuh"').replace('#include "b2_h2.cuh"', '#include "emul_h2.cuh"')
    open(os.path.join(HERE, "emul_lib.cc"), "w").write("// GENERATED by tests/cpp/gen_emul_lib.py from brpc_b200/csrc/b2_api.cu - do not edit\n" + api)


if __name__ == "__main__":
    main()
