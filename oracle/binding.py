"""ctypes binding of the oracle (oracle/_build/libbjxa_oracle.so) and, when it
has been built, of the unmodified reference (oracle/_ref/libbjxa_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and
by bench.py's cpu_baseline / --impl reference legs -- never by bjxa_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.environ.get("BJXA_ORACLE_SO") or os.path.join(HERE, "_build", "libbjxa_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbjxa_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "bjxa_ref")


def build(ref: bool = True) -> None:
    """Compile the oracle (and the reference when /root/reference exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


class _Stream(C.Structure):
    _fields_ = [("data_len", C.c_uint32), ("samples", C.c_uint32),
                ("rate", C.c_uint16), ("bits", C.c_uint8),
                ("channels", C.c_uint8), ("prev", (C.c_int16 * 2) * 2)]


class Oracle:
    """The plain-C restatement (oracle/bjxa_oracle.c)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.dll = d = C.CDLL(path)
        d.xao_decode_blocks.restype = C.c_long
        d.xao_decode_blocks.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_void_p,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        d.xao_encode_blocks.restype = C.c_long
        d.xao_encode_blocks.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_uint32,
                                        C.c_void_p]
        d.xao_xa_to_wav.restype = C.c_long
        d.xao_xa_to_wav.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        d.xao_wav_to_xa.restype = C.c_long
        d.xao_wav_to_xa.argtypes = [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p,
                                    C.c_size_t]
        d.xao_parse_xa_header.restype = C.c_int
        d.xao_parse_xa_header.argtypes = [C.POINTER(_Stream), C.c_void_p]

    def decode_blocks(self, bits, channels, prev, xa: np.ndarray, blocks: int,
                      pcm_bytes: int):
        """Returns (blocks_done, bad, pcm int16 array, final prev[2][2]).

        `prev` is [[p0,p1],[p0,p1]]; `pcm_bytes` is the PCM still owed by the
        stream (caps the last block like fmt->data_len_pcm does)."""
        xa = np.ascontiguousarray(xa, dtype=np.uint8)
        st = np.array(prev, dtype=np.int16).reshape(2, 2).copy()
        out = np.zeros(blocks * 32 * channels, dtype=np.int16)
        left = C.c_uint32(pcm_bytes)
        bad = C.c_int(0)
        done = self.dll.xao_decode_blocks(bits, channels, st.ctypes.data,
                                          xa.ctypes.data, blocks, out.ctypes.data,
                                          C.byref(left), C.byref(bad))
        written = pcm_bytes - left.value
        return done, bool(bad.value), out[:written // 2], st

    def encode_blocks(self, bits, channels, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        frames = pcm.size // channels
        blocks = (frames + 31) // 32
        out = np.zeros(blocks * channels * (4 * bits + 1), dtype=np.uint8)
        done = self.dll.xao_encode_blocks(bits, channels, pcm.ctypes.data,
                                          pcm.size * 2, out.ctypes.data)
        assert done == blocks
        return out

    def encode_search_blocks(self, bits, channels, prev, pcm: np.ndarray):
        """The searching encoder (an extension, see bjxa_oracle.h).
        -> (xa bytes, decoder state after the last block [[n-1, n-2], ...])"""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        frames = pcm.size // channels
        blocks = (frames + 31) // 32
        out = np.zeros(blocks * channels * (4 * bits + 1), dtype=np.uint8)
        st = ((C.c_int16 * 2) * 2)()
        for c in range(2):
            st[c][0], st[c][1] = int(prev[c][0]), int(prev[c][1])
        self.dll.xao_encode_search_blocks.restype = C.c_long
        done = self.dll.xao_encode_search_blocks(
            C.c_uint(bits), C.c_uint(channels), st, C.c_void_p(pcm.ctypes.data),
            C.c_uint32(pcm.size * 2), C.c_void_p(out.ctypes.data))
        assert done == blocks, done
        return out, [[st[c][0], st[c][1]] for c in range(2)]

    def parse_xa_header(self, hdr: bytes):
        st = _Stream()
        buf = (C.c_char * 32).from_buffer_copy(hdr[:32])
        rc = self.dll.xao_parse_xa_header(C.byref(st), buf)
        if rc < 0:
            return None
        return {"data_len": st.data_len, "samples": st.samples, "rate": st.rate,
                "bits": st.bits, "channels": st.channels,
                "prev": [[st.prev[0][0], st.prev[0][1]],
                         [st.prev[1][0], st.prev[1][1]]]}

    def xa_to_wav(self, xa: bytes):
        src = np.frombuffer(xa, dtype=np.uint8)
        cap = 44 + 64 * (len(xa) // 17 + 2) * 2
        dst = np.zeros(cap, dtype=np.uint8)
        n = self.dll.xao_xa_to_wav(src.ctypes.data, src.size, dst.ctypes.data, cap)
        return None if n < 0 else dst[:n].tobytes()

    def wav_to_xa(self, wav: bytes, bits: int):
        src = np.frombuffer(wav, dtype=np.uint8)
        cap = 32 + 66 * (len(wav) // 64 + 2)
        dst = np.zeros(cap, dtype=np.uint8)
        n = self.dll.xao_wav_to_xa(src.ctypes.data, src.size, bits,
                                   dst.ctypes.data, cap)
        return None if n < 0 else dst[:n].tobytes()


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def reference_lib():
    """The unmodified reference, bound through the same ctypes class as the
    product (bjxa_b200.capi.BjxaLib) so both are driven identically."""
    from bjxa_b200.capi import BjxaLib
    return BjxaLib(REF_SO)
