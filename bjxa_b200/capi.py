"""ctypes binding of the libbjxa C ABI (include/bjxa.h).

`BjxaLib(path)` binds the 19 public symbols of any libbjxa-compatible shared
object: the product library (bjxa_b200/lib/libbjxa_b200.so), or -- from the
tests and the CPU-baseline leg of the benchmark only -- a build of the
unmodified reference library.  Names, argument order and error behaviour are the
reference's (/root/reference/src/bjxa.h:36-65): every call returns what the C
function returns and `BjxaLib.errno()` reads the C errno.
"""
from __future__ import annotations

import ctypes as C
import os

HEADER_SIZE_XA = 32     # /root/reference/src/bjxa.h:18
HEADER_SIZE_RIFF = 44   # /root/reference/src/bjxa.h:19


class Format(C.Structure):
    """bjxa_format_t (/root/reference/src/bjxa.h:24-32); field order is ABI."""
    _fields_ = [
        ("data_len_pcm", C.c_uint32),
        ("blocks", C.c_uint32),
        ("block_size_pcm", C.c_uint8),
        ("block_size_xa", C.c_uint8),
        ("samples_rate", C.c_uint16),
        ("sample_bits", C.c_uint8),
        ("channels", C.c_uint8),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_VP = C.c_void_p
_SZ = C.c_size_t
_SSZ = C.c_ssize_t

# name -> (restype, argtypes); the 19 symbols of src/libbjxa.map:16-47
SYMBOLS = {
    "bjxa_decoder": (_VP, []),
    "bjxa_free_decoder": (C.c_int, [C.POINTER(_VP)]),
    "bjxa_parse_header": (_SSZ, [_VP, _VP, _SZ]),
    "bjxa_fread_header": (_SSZ, [_VP, _VP]),
    "bjxa_decode_format": (C.c_int, [_VP, C.POINTER(Format)]),
    "bjxa_decode": (C.c_int, [_VP, _VP, _SZ, _VP, _SZ]),
    "bjxa_dump_riff_header": (_SSZ, [_VP, _VP, _SZ]),
    "bjxa_fwrite_riff_header": (_SSZ, [_VP, _VP]),
    "bjxa_dump_pcm": (C.c_int, [_VP, _VP, _SZ]),
    "bjxa_fwrite_pcm": (C.c_int, [_VP, _SZ, _VP]),
    "bjxa_encoder": (_VP, []),
    "bjxa_free_encoder": (C.c_int, [C.POINTER(_VP)]),
    "bjxa_encode_init": (C.c_int, [_VP, C.POINTER(Format), C.c_uint8]),
    "bjxa_parse_riff_header": (_SSZ, [C.POINTER(Format), _VP, _SZ]),
    "bjxa_fread_riff_header": (_SSZ, [C.POINTER(Format), _VP]),
    "bjxa_encode_format": (C.c_int, [_VP, C.POINTER(Format)]),
    "bjxa_encode": (C.c_int, [_VP, _VP, _SZ, _VP, _SZ]),
    "bjxa_dump_header": (_SSZ, [_VP, _VP, _SZ]),
    "bjxa_fwrite_header": (_SSZ, [_VP, _VP]),
}


def _addr(buf):
    """Address of a bytes / bytearray / numpy / ctypes buffer, or None."""
    if buf is None:
        return None
    if isinstance(buf, int):
        return buf
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), _VP).value
    if hasattr(buf, "ctypes"):          # numpy
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):        # torch (CPU tensor)
        return buf.data_ptr()
    return C.addressof((C.c_char * len(buf)).from_buffer(buf))


class BjxaLib:
    """One loaded libbjxa-compatible shared object."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it first (python -c 'import "
                f"__graft_entry__ as g; g.build()'); there is no fallback path")
        self.path = path
        self.dll = C.CDLL(path, use_errno=True)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.dll, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, "_" + name, fn)

    # -- errno ---------------------------------------------------------------
    @staticmethod
    def errno() -> int:
        return C.get_errno()

    @staticmethod
    def clear_errno():
        C.set_errno(0)

    # -- decoder -------------------------------------------------------------
    def decoder(self):
        return self._bjxa_decoder()

    def free_decoder(self, dec) -> int:
        p = _VP(dec)
        return self._bjxa_free_decoder(C.byref(p))

    def parse_header(self, dec, src, length=None) -> int:
        n = len(src) if length is None and src is not None else (length or 0)
        self._keep = src
        return self._bjxa_parse_header(dec, _addr(src), n)

    def decode_format(self, dec):
        fmt = Format()
        rc = self._bjxa_decode_format(dec, C.byref(fmt))
        return rc, fmt

    def decode(self, dec, dst, dst_len, src, src_len) -> int:
        self._keep = (dst, src)
        return self._bjxa_decode(dec, _addr(dst), dst_len, _addr(src), src_len)

    def dump_riff_header(self, dec, dst, length=None) -> int:
        n = len(dst) if length is None and dst is not None else (length or 0)
        return self._bjxa_dump_riff_header(dec, _addr(dst), n)

    def dump_pcm(self, dst, src, length) -> int:
        return self._bjxa_dump_pcm(_addr(dst), _addr(src), length)

    # -- encoder -------------------------------------------------------------
    def encoder(self):
        return self._bjxa_encoder()

    def free_encoder(self, enc) -> int:
        p = _VP(enc)
        return self._bjxa_free_encoder(C.byref(p))

    def parse_riff_header(self, src, length=None):
        fmt = Format()
        n = len(src) if length is None and src is not None else (length or 0)
        self._keep = src
        rc = self._bjxa_parse_riff_header(C.byref(fmt), _addr(src), n)
        return rc, fmt

    def encode_init(self, enc, fmt: Format, bits: int) -> int:
        return self._bjxa_encode_init(enc, C.byref(fmt), bits)

    def encode_format(self, enc):
        fmt = Format()
        rc = self._bjxa_encode_format(enc, C.byref(fmt))
        return rc, fmt

    def encode(self, enc, dst, dst_len, src, src_len) -> int:
        self._keep = (dst, src)
        return self._bjxa_encode(enc, _addr(dst), dst_len, _addr(src), src_len)

    def dump_header(self, enc, dst, length=None) -> int:
        n = len(dst) if length is None and dst is not None else (length or 0)
        return self._bjxa_dump_header(enc, _addr(dst), n)

    # -- whole-file helpers, written the way src/bjxa_decode.c:57-100 and
    #    src/bjxa_encode.c:63-106 (single-pass mode) drive the API ------------
    def xa_to_wav(self, xa: bytes) -> bytes:
        dec = self.decoder()
        try:
            if self.parse_header(dec, xa[:HEADER_SIZE_XA], HEADER_SIZE_XA) < 0:
                raise OSError(self.errno(), "bjxa_parse_header")
            rc, fmt = self.decode_format(dec)
            if rc < 0:
                raise OSError(self.errno(), "bjxa_decode_format")
            hdr = bytearray(HEADER_SIZE_RIFF)
            if self.dump_riff_header(dec, hdr) < 0:
                raise OSError(self.errno(), "bjxa_dump_riff_header")
            xa_len = fmt.block_size_xa * fmt.blocks
            pcm = bytearray(max(fmt.data_len_pcm, fmt.block_size_pcm))
            pay = xa[HEADER_SIZE_XA:HEADER_SIZE_XA + xa_len]
            got = self.decode(dec, pcm, len(pcm), pay, len(pay))
            if got != fmt.blocks:
                raise OSError(self.errno(), f"bjxa_decode returned {got}")
            return bytes(hdr) + bytes(pcm[:fmt.data_len_pcm])
        finally:
            self.free_decoder(dec)

    def wav_to_xa(self, wav: bytes, bits: int) -> bytes:
        enc = self.encoder()
        try:
            rc, fmt = self.parse_riff_header(wav[:HEADER_SIZE_RIFF], HEADER_SIZE_RIFF)
            if rc < 0:
                raise OSError(self.errno(), "bjxa_parse_riff_header")
            if self.encode_init(enc, fmt, bits) < 0:
                raise OSError(self.errno(), "bjxa_encode_init")
            rc, fmt = self.encode_format(enc)
            if rc < 0:
                raise OSError(self.errno(), "bjxa_encode_format")
            hdr = bytearray(HEADER_SIZE_XA)
            if self.dump_header(enc, hdr) < 0:
                raise OSError(self.errno(), "bjxa_dump_header")
            xa_len = fmt.block_size_xa * fmt.blocks
            out = bytearray(xa_len)
            pcm = wav[HEADER_SIZE_RIFF:HEADER_SIZE_RIFF + fmt.data_len_pcm]
            # the reference insists on a full PCM block of input even when the
            # stream is shorter (src/libbjxa.c:778)
            if len(pcm) < fmt.block_size_pcm:
                pcm = pcm + bytes(fmt.block_size_pcm - len(pcm))
            got = self.encode(enc, out, len(out), pcm, len(pcm))
            if got != fmt.blocks:
                raise OSError(self.errno(), f"bjxa_encode returned {got}")
            return bytes(hdr) + bytes(out)
        finally:
            self.free_encoder(enc)
