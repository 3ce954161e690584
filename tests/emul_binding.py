"""ctypes binding of tests/_build/libxa_emul.so -- the CPU single-stepper of the
tile code (tests/emul/xa_emul.cc).  Test harness only."""
import ctypes as C
import os
import subprocess

import numpy as np

from bjxa_b200.api import DESC_DTYPE, make_descs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.environ.get("XA_EMUL_SO") or os.path.join(ROOT, "tests", "_build", "libxa_emul.so")


class Emul:
    def __init__(self):
        if not os.environ.get("XA_EMUL_SO"):
            subprocess.run(["make", "-s", "-C", ROOT, "emul"], check=True)
        self.dll = d = C.CDLL(SO)
        d.xa_emul_decode.restype = C.c_int
        d.xa_emul_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        d.xa_emul_encode.restype = C.c_int
        d.xa_emul_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_uint64, C.c_int]
        d.xa_emul_plan.restype = C.c_int
        d.xa_emul_plan.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.strip_blocks = d.xa_emul_strip_blocks      # (ns, ch) -> effective blocks
        self.wide = d.xa_emul_wide()
        self.enc_tile_blocks = d.xa_emul_enc_tile_blocks
        self.stereo_direct = d.xa_emul_stereo_direct    # (0|1): which stereo form to step
        self.use_alt = d.xa_emul_use_alt                # (0|1): step the alternative tile lists
        self.pool = d.xa_emul_pool                      # (0|1): direct forms walk as the pooled kernel does
        self.relay = d.xa_emul_relay                    # (0|1): long-strip tiles walk their chains but hand stragglers on (relay form)
        self.seg = d.xa_emul_seg                        # (0|1): classes of enough streams go through the segment form (xa_walk.h)
        self.seg_items = d.xa_emul_seg_items()
        self.seg_back = d.xa_emul_seg_back()
        self.split = d.xa_emul_split                    # (0|1): long-strip tiles go through the split form (xa_walk.h)

    def dec_tile_blocks(self, ch):
        """effective blocks of one long (NS = 1) strip"""
        return self.strip_blocks(1, ch)

    def decode(self, descs, src, dst_bytes, order=0, strips=0):
        assert descs.dtype == DESC_DTYPE
        src = np.ascontiguousarray(src, dtype=np.uint8)
        dst = np.full(dst_bytes, 0xCD, dtype=np.uint8)
        prev = np.zeros((descs.size, 2, 2), dtype=np.int16)
        bad = np.zeros(descs.size, dtype=np.uint32)
        rc = self.dll.xa_emul_decode(descs.ctypes.data, descs.size, src.ctypes.data,
                                     src.size, dst.ctypes.data, prev.ctypes.data,
                                     bad.ctypes.data, order, strips)
        return rc, dst, prev, bad

    def encode(self, descs, src, dst_bytes, order=0):
        src = np.ascontiguousarray(src).view(np.uint8)
        dst = np.full(dst_bytes, 0xCD, dtype=np.uint8)
        rc = self.dll.xa_emul_encode(descs.ctypes.data, descs.size, src.ctypes.data,
                                     src.size, dst.ctypes.data, dst.size, order)
        return rc, dst

    def search(self, descs, src, dst_bytes):
        """The searching encoder's arithmetic (xa_core.h) in plain loops.
        -> (rc, xa arena, decoder state after the last block [n][2][2])"""
        src = np.ascontiguousarray(src).view(np.uint8)
        dst = np.full(dst_bytes, 0xCD, dtype=np.uint8)
        prev = np.zeros((descs.size, 2, 2), dtype=np.int16)
        self.dll.xa_emul_search.restype = C.c_int
        rc = self.dll.xa_emul_search(C.c_void_p(descs.ctypes.data), C.c_size_t(descs.size),
                                     C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data),
                                     C.c_void_p(prev.ctypes.data))
        return rc, dst, prev

    def plan(self, kind, descs, strips=0, cap=1 << 20):
        """-> (n_tiles, first stream of each tile, strip counts, j, tile_begin,
        n_slots, strips per tile of each bucket); the lists in the alternative
        tile shape, if any, are left in self.alt = (alt_begin, alt_ns)"""
        ts = np.zeros(cap, dtype=np.uint32)
        tc = np.zeros(cap, dtype=np.uint32)
        tj = np.zeros(cap, dtype=np.uint32)
        tb = np.zeros(7, dtype=np.uint32)
        nsb = np.zeros(6, dtype=np.int32)
        ns = C.c_uint32(0)
        ab = np.zeros(7, dtype=np.uint32)
        an = np.zeros(6, dtype=np.int32)
        n = self.dll.xa_emul_plan(kind, descs.ctypes.data, descs.size, strips,
                                  ts.ctypes.data, tc.ctypes.data, tj.ctypes.data, cap,
                                  tb.ctypes.data, C.byref(ns), nsb.ctypes.data,
                                  ab.ctypes.data, an.ctypes.data)
        self.alt = (ab, an)
        if n < 0:
            return n, None, None, None, None, None, None
        return n, ts[:n], tc[:n], tj[:n], tb, ns.value, nsb
