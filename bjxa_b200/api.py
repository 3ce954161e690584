"""Python mirror of the product's C ABI (include/bjxa.h + include/bjxa_batch.h).

`Bjxa` extends the plain libbjxa binding (capi.BjxaLib -- the reference's 19
functions) with the additive batched / device-resident calls.  It is a thin
ctypes layer: every method forwards to the shared library, and raises if the
library is missing.  No computation happens in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import BjxaLib, _addr

PLAN_DECODE = 0
PLAN_ENCODE = 1
PLAN_ENCODE_SEARCH = 2     # extension: per-block filter/range search (include/bjxa_batch.h)


class StreamDesc(C.Structure):
    """bjxa_stream_desc_t (include/bjxa_batch.h)."""
    _fields_ = [
        ("xa_off", C.c_uint64),
        ("pcm_off", C.c_uint64),
        ("blocks", C.c_uint32),
        ("pcm_len", C.c_uint32),
        ("prev", (C.c_int16 * 2) * 2),
        ("bits", C.c_uint8),
        ("channels", C.c_uint8),
        ("reserved", C.c_uint16),
        ("done", C.c_uint32),
        ("result", C.c_int32),
        ("error", C.c_int32),
    ]


DESC_DTYPE = np.dtype({
    "names": ["xa_off", "pcm_off", "blocks", "pcm_len", "prev", "bits", "channels",
              "reserved", "done", "result", "error"],
    "formats": ["<u8", "<u8", "<u4", "<u4", ("<i2", (2, 2)), "u1", "u1", "<u2",
                "<u4", "<i4", "<i4"],
    "offsets": [0, 8, 16, 20, 24, 32, 33, 34, 36, 40, 44],
    "itemsize": 48,
})
assert C.sizeof(StreamDesc) == DESC_DTYPE.itemsize == 48

FILE_DTYPE = np.dtype({
    "names": ["in_off", "in_len", "out_off", "out_len", "error", "blocks", "bits",
              "channels", "rate", "reserved"],
    "formats": ["<u8", "<u8", "<u8", "<u8", "<i4", "<u4", "u1", "u1", "<u2", "<u4"],
    "offsets": [0, 8, 16, 24, 32, 36, 40, 41, 42, 44],
    "itemsize": 48,
})      # bjxa_file_desc_t
CORPUS_XA_TO_WAV = 0
CORPUS_WAV_TO_XA = 1

_VP, _SZ = C.c_void_p, C.c_size_t

BATCH_SYMBOLS = {
    "bjxa_batch_decode": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "bjxa_batch_encode": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ]),
    "bjxa_plan_create": (_VP, [C.c_int, _VP, _SZ]),
    "bjxa_plan_reset": (C.c_int, [_VP, C.c_int, _VP, _SZ]),
    "bjxa_plan_run": (C.c_int, [_VP, _VP, _SZ, _VP, _SZ, _VP]),
    "bjxa_plan_fetch": (C.c_int, [_VP, _VP, _SZ]),
    "bjxa_plan_launches": (C.c_int, [_VP]),
    "bjxa_plan_launched": (C.c_uint64, [_VP]),
    "bjxa_plan_checksum": (C.c_int, [_VP, _VP, _SZ]),
    "bjxa_thread_release": (None, []),
    "bjxa_plan_extent": (C.c_int, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bjxa_plan_free": (C.c_int, [C.POINTER(_VP)]),
    "bjxa_decoder_describe": (C.c_int, [_VP, C.POINTER(StreamDesc)]),
    "bjxa_decoder_commit": (C.c_int, [_VP, C.POINTER(StreamDesc)]),
    "bjxa_encoder_describe": (C.c_int, [_VP, C.POINTER(StreamDesc)]),
    "bjxa_encoder_commit": (C.c_int, [_VP, C.POINTER(StreamDesc)]),
    "bjxa_gpu_count": (C.c_int, []),
    "bjxa_gpu_select": (C.c_int, [C.c_int]),
    "bjxa_gpu_alloc": (_VP, [_SZ]),
    "bjxa_gpu_free": (C.c_int, [_VP]),
    "bjxa_host_alloc": (_VP, [_SZ]),
    "bjxa_host_free": (C.c_int, [_VP]),
    "bjxa_gpu_upload": (C.c_int, [_VP, _VP, _SZ]),
    "bjxa_gpu_download": (C.c_int, [_VP, _VP, _SZ]),
    "bjxa_gpu_sync": (C.c_int, [_VP]),
    "bjxa_gpu_stream_create": (_VP, []),
    "bjxa_gpu_stream_destroy": (C.c_int, [_VP]),
    "bjxa_gpu_upload_async": (C.c_int, [_VP, _VP, _SZ, _VP]),
    "bjxa_gpu_download_async": (C.c_int, [_VP, _VP, _SZ, _VP]),
    "bjxa_shard_range": (C.c_int, [_VP, _SZ, C.c_int, C.c_int,
                                   C.POINTER(_SZ), C.POINTER(_SZ)]),
    "bjxa_gpu_scatter_async": (C.c_int, [_VP, _VP, C.c_uint32, _SZ, _VP]),
    "bjxa_corpus_extent": (C.c_int, [C.c_int, _VP, _SZ, _VP, _SZ, C.POINTER(C.c_uint64)]),
    "bjxa_corpus_run": (C.c_int, [C.c_int, _VP, _SZ, _VP, _SZ, _VP, _SZ]),
}


def make_descs(n: int) -> np.ndarray:
    return np.zeros(n, dtype=DESC_DTYPE)


class Bjxa(BjxaLib):
    """The product library: reference API + batched/device API."""

    def __init__(self, path: str):
        super().__init__(path)
        for name, (res, args) in BATCH_SYMBOLS.items():
            fn = getattr(self.dll, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, "_" + name, fn)

    def _check(self, rc, what):
        if rc is None or (isinstance(rc, int) and rc < 0):
            raise OSError(self.errno(), f"{what} failed")
        return rc

    # -- device helpers --------------------------------------------------------
    def gpu_count(self) -> int:
        return self._bjxa_gpu_count()

    def gpu_alloc(self, nbytes: int) -> int:
        p = self._bjxa_gpu_alloc(nbytes)
        if not p:
            raise OSError(self.errno(), "bjxa_gpu_alloc failed")
        return p

    def gpu_free(self, p: int):
        self._check(self._bjxa_gpu_free(p), "bjxa_gpu_free")

    def upload(self, dptr: int, host) -> None:
        host = np.ascontiguousarray(host)
        self._check(self._bjxa_gpu_upload(dptr, host.ctypes.data, host.nbytes),
                    "bjxa_gpu_upload")

    def download(self, dptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        self._check(self._bjxa_gpu_download(out.ctypes.data, dptr, nbytes),
                    "bjxa_gpu_download")
        return out

    def sync(self, stream: int = 0):
        self._check(self._bjxa_gpu_sync(stream), "bjxa_gpu_sync")

    # -- plans -----------------------------------------------------------------
    def plan_create(self, kind: int, descs: np.ndarray) -> int:
        assert descs.dtype == DESC_DTYPE
        p = self._bjxa_plan_create(kind, descs.ctypes.data, descs.size)
        if not p:
            raise OSError(self.errno(), "bjxa_plan_create failed")
        return p

    def plan_reset(self, plan: int, kind: int, descs: np.ndarray):
        self._check(self._bjxa_plan_reset(plan, kind, descs.ctypes.data, descs.size),
                    "bjxa_plan_reset")

    def plan_run(self, plan: int, dst: int, dst_bytes: int, src: int, src_bytes: int,
                 stream: int = 0):
        self._check(self._bjxa_plan_run(plan, dst, dst_bytes, src, src_bytes, stream),
                    "bjxa_plan_run")

    def plan_fetch(self, plan: int, n: int) -> np.ndarray:
        out = make_descs(n)
        self._check(self._bjxa_plan_fetch(plan, out.ctypes.data, n), "bjxa_plan_fetch")
        return out

    def plan_launches(self, plan: int) -> int:
        return self._bjxa_plan_launches(plan)

    def plan_launched(self, plan: int) -> int:
        """kernels bjxa_plan_run has launched for this plan so far"""
        return int(self._bjxa_plan_launched(plan))

    def plan_checksum(self, plan: int, n: int) -> np.ndarray:
        """Per-stream checksums of the last run's output, computed on the device
        (bjxa_plan_checksum); synth.stream_checksum is the same sum in numpy."""
        sums = np.zeros(n, dtype=np.uint64)
        self._check(self._bjxa_plan_checksum(plan, sums.ctypes.data, n), "bjxa_plan_checksum")
        return sums

    def plan_extent(self, plan: int):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self._bjxa_plan_extent(plan, C.byref(a), C.byref(b)),
                    "bjxa_plan_extent")
        return a.value, b.value

    def plan_free(self, plan: int):
        p = _VP(plan)
        self._check(self._bjxa_plan_free(C.byref(p)), "bjxa_plan_free")

    # -- codec <-> descriptor ----------------------------------------------------
    def decoder_describe(self, dec) -> StreamDesc:
        d = StreamDesc()
        self._check(self._bjxa_decoder_describe(dec, C.byref(d)), "bjxa_decoder_describe")
        return d

    def decoder_commit(self, dec, d: StreamDesc):
        self._check(self._bjxa_decoder_commit(dec, C.byref(d)), "bjxa_decoder_commit")

    # -- host-buffer batches -----------------------------------------------------
    def _batch(self, fn, codecs, dsts, srcs):
        n = len(codecs)
        arr_c = (_VP * n)(*codecs)
        arr_d = (_VP * n)(*[_addr(d) for d in dsts])
        arr_s = (_VP * n)(*[_addr(s) for s in srcs])
        len_d = (_SZ * n)(*[len(d) if not isinstance(d, np.ndarray) else d.nbytes
                            for d in dsts])
        len_s = (_SZ * n)(*[len(s) if not isinstance(s, np.ndarray) else s.nbytes
                            for s in srcs])
        res = (C.c_int * n)()
        errs = (C.c_int * n)()
        self._keep = (dsts, srcs)
        rc = fn(arr_c, arr_d, len_d, arr_s, len_s, res, errs, n)
        if rc < 0:
            raise OSError(self.errno(), "batch call failed")
        return list(res), list(errs)

    def batch_decode(self, decs, dsts, srcs):
        return self._batch(self._bjxa_batch_decode, decs, dsts, srcs)

    def batch_encode(self, encs, dsts, srcs):
        return self._batch(self._bjxa_batch_encode, encs, dsts, srcs)

    # -- whole files ----------------------------------------------------------------
    def corpus(self, kind: int, files, bits: int = 0, align: bool = True):
        """files: list of bytes (whole .xa or .wav files).  Lays them out in one
        pinned arena, runs bjxa_corpus_run, returns (table, list of produced
        files as bytes -- empty where the table's error is set and nothing was
        produced)."""
        table = np.zeros(len(files), dtype=FILE_DTYPE)
        off = 0
        for i, f in enumerate(files):
            if kind == CORPUS_WAV_TO_XA and align:
                off += (-(off + 44)) % 16          # PCM data 16-byte aligned
            table[i]["in_off"], table[i]["in_len"] = off, len(f)
            off += len(f)
        table["bits"] = bits
        in_bytes = max(off, 16)
        h_in = self._bjxa_host_alloc(in_bytes)
        if not h_in:
            raise MemoryError("bjxa_host_alloc")
        h_out = None
        try:
            arena = np.ctypeslib.as_array((C.c_uint8 * in_bytes).from_address(h_in))
            arena[:] = 0xEE
            for i, f in enumerate(files):
                o = int(table[i]["in_off"])
                arena[o:o + len(f)] = np.frombuffer(f, dtype=np.uint8)
            need = C.c_uint64(0)
            self._check(self._bjxa_corpus_extent(kind, h_in, in_bytes, table.ctypes.data,
                                                 len(files), C.byref(need)),
                        "bjxa_corpus_extent")
            out_bytes = int(need.value)
            h_out = self._bjxa_host_alloc(out_bytes)
            if not h_out:
                raise MemoryError("bjxa_host_alloc")
            out = np.ctypeslib.as_array((C.c_uint8 * out_bytes).from_address(h_out))
            out[:] = 0xCD
            self._check(self._bjxa_corpus_run(kind, h_in, in_bytes, h_out, out_bytes,
                                              table.ctypes.data, len(files)),
                        "bjxa_corpus_run")
            made = [out[int(t["out_off"]):int(t["out_off"] + t["out_len"])].tobytes()
                    for t in table]
            self.last_corpus_out = out.copy()
            return table, made
        finally:
            self._bjxa_host_free(h_in)
            if h_out:
                self._bjxa_host_free(h_out)

    # -- sharding ------------------------------------------------------------------
    def shard_range(self, n: int, rank: int, world: int, nbytes=None):
        first, count = _SZ(0), _SZ(0)
        ptr = None
        if nbytes is not None:
            nbytes = np.ascontiguousarray(nbytes, dtype=np.uint64)
            ptr = nbytes.ctypes.data
        self._check(self._bjxa_shard_range(ptr, n, rank, world, C.byref(first),
                                           C.byref(count)), "bjxa_shard_range")
        return first.value, count.value
