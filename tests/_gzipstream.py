"""The system zlib driven the way google::protobuf::io::GzipInputStream drives it (tests only).

brpc's gzip / zlib decompress handlers (src/brpc/policy/gzip_compress.cpp:75-89) wrap the body in a GzipInputStream and let
the protobuf parser pull from it.  protobuf's gzip_stream.cc is not in the image, zlib is: this module restates
GzipInputStream::Next / Inflate / DoNextOutput call for call on top of the real inflate() (ctypes), with the sub-stream being
ONE block holding the whole body.  It is the pin of oracle/b2_oracle_gzip.c.
"""
import ctypes as C
import ctypes.util

_z = C.CDLL(ctypes.util.find_library("z") or "libz.so.1")

Z_OK, Z_STREAM_END, Z_BUF_ERROR, Z_NO_FLUSH = 0, 1, -5, 0
GZIP, ZLIB = 2, 3                     # == B2_COMPRESS_TYPE_GZIP / _ZLIB
K_BUFFER = 65536                      # kDefaultBufferSize


class ZStream(C.Structure):
    _fields_ = [("next_in", C.c_void_p), ("avail_in", C.c_uint), ("total_in", C.c_ulong),
                ("next_out", C.c_void_p), ("avail_out", C.c_uint), ("total_out", C.c_ulong),
                ("msg", C.c_char_p), ("state", C.c_void_p), ("zalloc", C.c_void_p), ("zfree", C.c_void_p), ("opaque", C.c_void_p),
                ("data_type", C.c_int), ("adler", C.c_ulong), ("reserved", C.c_ulong)]


_z.zlibVersion.restype = C.c_char_p
_z.inflateInit2_.argtypes = [C.POINTER(ZStream), C.c_int, C.c_char_p, C.c_int]
_z.inflate.argtypes = [C.POINTER(ZStream), C.c_int]
_z.inflateEnd.argtypes = [C.POINTER(ZStream)]


def _init(zs, fmt):
    return _z.inflateInit2_(C.byref(zs), 15 | (16 if fmt == GZIP else 0), _z.zlibVersion(), C.sizeof(ZStream))


def gzip_input_stream(data: bytes, fmt: int) -> bytes:
    """All bytes Next() hands out before it returns false."""
    zs = ZStream()
    src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
    outbuf = C.create_string_buffer(K_BUFFER)
    out_base = C.addressof(outbuf)
    sub_used = [False]                 # the sub-stream yields its single block once (an empty IOBuf yields nothing)
    st = {"zerror": Z_OK, "output_position": out_base}
    zs.next_out = out_base; zs.avail_out = K_BUFFER
    got = bytearray()

    def inflate_call():
        if st["zerror"] == Z_OK and zs.avail_out == 0:
            pass                        # previous inflate filled the output buffer: input parameters stay
        elif zs.avail_in == 0:
            first = not zs.next_in
            if sub_used[0] or len(data) == 0:
                zs.next_out = None; zs.avail_out = 0
                return Z_STREAM_END
            sub_used[0] = True
            zs.next_in = C.addressof(src); zs.avail_in = len(data)
            if first:
                e = _init(zs, fmt)
                if e != Z_OK:
                    return e
        zs.next_out = out_base; zs.avail_out = K_BUFFER
        st["output_position"] = out_base
        return _z.inflate(C.byref(zs), Z_NO_FLUSH)

    def do_next_output():
        n = (zs.next_out or 0) - st["output_position"]
        got.extend(C.string_at(st["output_position"], n))
        st["output_position"] = zs.next_out

    try:
        while True:                     # one iteration = one Next()
            ok = st["zerror"] in (Z_OK, Z_STREAM_END, Z_BUF_ERROR)
            if not ok or not zs.next_out:
                break
            if zs.next_out != st["output_position"]:
                do_next_output(); continue
            if st["zerror"] == Z_STREAM_END:
                st["zerror"] = _z.inflateEnd(C.byref(zs))      # (next_out != NULL here) the sub-stream may hold further members
                if st["zerror"] != Z_OK:
                    break
                st["zerror"] = _init(zs, fmt)
                if st["zerror"] != Z_OK:
                    break
            st["zerror"] = inflate_call()
            if st["zerror"] == Z_STREAM_END and not zs.next_out:
                break                   # the sub-stream's Next returned false inside Inflate
            if st["zerror"] not in (Z_OK, Z_STREAM_END, Z_BUF_ERROR):
                break
            do_next_output()
    finally:
        if zs.state:
            _z.inflateEnd(C.byref(zs))
    return bytes(got)
