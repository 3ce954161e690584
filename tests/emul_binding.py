"""ctypes binding of tests/_build/libxa_emul.so -- the CPU single-stepper of the
tile code (tests/emul/xa_emul.cc).  Test harness only."""
import ctypes as C
import os
import subprocess

import numpy as np

from bjxa_b200.api import DESC_DTYPE, make_descs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_build", "libxa_emul.so")


class Emul:
    def __init__(self):
        subprocess.run(["make", "-s", "-C", ROOT, "emul"], check=True)
        self.dll = d = C.CDLL(SO)
        d.xa_emul_decode.restype = C.c_int
        d.xa_emul_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        d.xa_emul_encode.restype = C.c_int
        d.xa_emul_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_uint64, C.c_int]
        d.xa_emul_plan.restype = C.c_int
        d.xa_emul_plan.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                   C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        self.dec_tile_blocks = d.xa_emul_dec_tile_blocks
        self.enc_tile_blocks = d.xa_emul_enc_tile_blocks

    def decode(self, descs, src, dst_bytes, order=0):
        assert descs.dtype == DESC_DTYPE
        src = np.ascontiguousarray(src, dtype=np.uint8)
        dst = np.full(dst_bytes, 0xCD, dtype=np.uint8)
        prev = np.zeros((descs.size, 2, 2), dtype=np.int16)
        bad = np.zeros(descs.size, dtype=np.uint32)
        rc = self.dll.xa_emul_decode(descs.ctypes.data, descs.size, src.ctypes.data,
                                     src.size, dst.ctypes.data, prev.ctypes.data,
                                     bad.ctypes.data, order)
        return rc, dst, prev, bad

    def encode(self, descs, src, dst_bytes, order=0):
        src = np.ascontiguousarray(src).view(np.uint8)
        dst = np.full(dst_bytes, 0xCD, dtype=np.uint8)
        rc = self.dll.xa_emul_encode(descs.ctypes.data, descs.size, src.ctypes.data,
                                     src.size, dst.ctypes.data, dst.size, order)
        return rc, dst

    def plan(self, kind, descs, cap=1 << 20):
        ts = np.zeros(cap, dtype=np.uint32)
        tf = np.zeros(cap, dtype=np.uint32)
        tb = np.zeros(7, dtype=np.uint32)
        ns = C.c_uint32(0)
        n = self.dll.xa_emul_plan(kind, descs.ctypes.data, descs.size, ts.ctypes.data,
                                  tf.ctypes.data, cap, tb.ctypes.data, C.byref(ns))
        if n < 0:
            return n, None, None, None, None
        return n, ts[:n], tf[:n], tb, ns.value
