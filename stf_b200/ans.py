"""Drop-in for the reference's `compressai.ans` pybind11 module, backed by the host rANS codec in
libstf_b200.so (stf_b200/csrc/rans_host.cpp).  Same class and method names, same argument order
and the same bytes out (reference compressai/cpp_exts/rans/rans_interface.cpp:352-372):

    BufferedRansEncoder().encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets); .flush()
    RansEncoder().encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes
    RansDecoder().set_stream(b); .decode_stream(indexes, cdfs, cdfs_sizes, offsets) -> list[int]
    RansDecoder().decode_with_indexes(b, indexes, cdfs, cdfs_sizes, offsets) -> list[int]

Beyond the reference: every sequence argument may be an int32 numpy array / CPU torch tensor
(zero-copy), `cdfs` may be a prepared `RansTable` (then sizes / offsets are ignored), decoders
have `decode_stream_array` returning numpy, and `encode_batch` / `decode_batch` code one stream
per image on several host threads.
"""
import ctypes
import os

import numpy as np

from . import _C


def _as_i32(a):
    """int32, C-contiguous numpy view of a list / numpy array / CPU torch tensor."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32)


class RansTable:
    """Prepared CDF table set (encoder reciprocals + decoder LUTs), reusable across calls."""

    def __init__(self, cdfs, cdfs_sizes, offsets):
        cdf = _as_i32(cdfs)
        if cdf.ndim != 2:
            raise ValueError(f"Invalid CDF size {cdf.shape}")
        sizes, offs = _as_i32(cdfs_sizes).reshape(-1), _as_i32(offsets).reshape(-1)
        if sizes.size != cdf.shape[0] or offs.size != cdf.shape[0]:
            raise ValueError("cdfs, cdfs_sizes and offsets disagree on the number of rows")
        L = _C.lib()
        self._h = L.stf_rans_table_create(cdf.ctypes.data_as(_C._i32p), cdf.shape[0], cdf.shape[1],
                                          sizes.ctypes.data_as(_C._i32p), offs.ctypes.data_as(_C._i32p))
        if not self._h:
            raise ValueError("malformed CDF table (rows must start at 0, end at 65536 and increase strictly)")
        self._destroy = L.stf_rans_table_destroy      # bound now: module globals may be gone at interpreter exit
        self.rows = cdf.shape[0]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._destroy(h)

    def device_image(self, device):
        """The table packed for the device decoder (stf_rans_device_table_pack) as a CUDA uint8 tensor, cached per device."""
        import torch
        cache = self.__dict__.setdefault("_dev_images", {})
        key = str(device)
        if key not in cache:
            L = _C.lib()
            nbytes = int(L.stf_rans_device_table_bytes(self._h))
            if nbytes <= 0:
                _C.check(nbytes, "stf_rans_device_table_bytes")
            host = np.zeros(nbytes, dtype=np.uint8)
            _C.check(L.stf_rans_device_table_pack(self._h, host.ctypes.data), "stf_rans_device_table_pack")
            cache[key] = torch.from_numpy(host).to(device)
        return cache[key]


def _table(cdfs, sizes, offsets):
    return cdfs if isinstance(cdfs, RansTable) else RansTable(cdfs, sizes, offsets)


def encode_array(table: RansTable, symbols, indexes) -> bytes:
    s, ix = _as_i32(symbols).reshape(-1), _as_i32(indexes).reshape(-1)
    if s.size != ix.size:
        raise ValueError("symbols and indexes differ in length")
    L = _C.lib()
    cap = L.stf_rans_encode_bound(s.size)
    out = np.empty(cap, dtype=np.uint8)
    n = L.stf_rans_encode(table._h, s.ctypes.data, ix.ctypes.data, s.size, out.ctypes.data, cap)
    if n < 0:
        _C.check(int(n), "stf_rans_encode")
    return out[:n].tobytes()


def default_threads():
    """Host threads of this process for the per-image rANS streams: the cores this process may run on, shared evenly
    between the ranks of the node (torchrun's LOCAL_WORLD_SIZE) so that N processes do not oversubscribe the host."""
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(32, cores // ranks))


def encode_batch(table: RansTable, symbols, indexes, threads=None):
    """symbols / indexes: sequences (or 2-D arrays) of per-image int32 arrays -> list[bytes]."""
    syms = [_as_i32(s).reshape(-1) for s in symbols]
    idxs = [_as_i32(i).reshape(-1) for i in indexes]
    count = len(syms)
    if count == 0:
        return []
    L = _C.lib()
    caps = [int(L.stf_rans_encode_bound(s.size)) for s in syms]
    outs = [np.empty(c, dtype=np.uint8) for c in caps]
    VP, I64 = ctypes.c_void_p * count, ctypes.c_int64 * count
    lens = I64()
    rc = L.stf_rans_encode_batch(table._h, count, VP(*[s.ctypes.data for s in syms]), VP(*[i.ctypes.data for i in idxs]),
                                 I64(*[s.size for s in syms]), VP(*[o.ctypes.data for o in outs]), I64(*caps), lens,
                                 threads or default_threads())
    _C.check(rc, "stf_rans_encode_batch")
    return [outs[i][: lens[i]].tobytes() for i in range(count)]


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None) -> bytes:
        return encode_array(_table(cdfs, cdfs_sizes, offsets), symbols, indexes)


class BufferedRansEncoder:
    """Accumulates (symbols, indexes) chunks; flush() codes them as ONE stream in push order."""

    def __init__(self):
        self._sym, self._idx, self._tab = [], [], None

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None):
        self._tab = _table(cdfs, cdfs_sizes, offsets)
        self._sym.append(_as_i32(symbols).reshape(-1))
        self._idx.append(_as_i32(indexes).reshape(-1))

    def flush(self) -> bytes:
        if self._tab is None:
            raise ValueError("flush() before encode_with_indexes()")
        out = encode_array(self._tab, np.concatenate(self._sym), np.concatenate(self._idx))
        self._sym, self._idx = [], []
        return out


class RansDecoder:
    def __init__(self):
        self._h = None
        self._destroy = _C.lib().stf_rans_decoder_destroy   # bound now: module globals may be gone at interpreter exit

    def __del__(self):
        self._close()

    def _close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._destroy(h)

    def set_stream(self, encoded: bytes):
        self._close()
        buf = np.frombuffer(encoded if isinstance(encoded, bytes) else bytes(encoded), dtype=np.uint8)
        self._buf = buf                      # zero-copy: the decoder reads the caller's bytes object in place
        self._h = _C.lib().stf_rans_decoder_create_view(buf.ctypes.data, buf.size)
        if not self._h:
            raise ValueError("invalid rANS stream (need a multiple of 4 bytes, at least 8)")

    def decode_stream_array(self, indexes, cdfs, cdfs_sizes=None, offsets=None) -> np.ndarray:
        if not self._h:
            raise ValueError("set_stream() first")
        tab = _table(cdfs, cdfs_sizes, offsets)
        ix = _as_i32(indexes).reshape(-1)
        out = np.empty(ix.size, dtype=np.int32)
        _C.check(_C.lib().stf_rans_decode(self._h, tab._h, ix.ctypes.data, ix.size, out.ctypes.data), "stf_rans_decode")
        return out

    def decode_stream(self, indexes, cdfs, cdfs_sizes=None, offsets=None):
        return self.decode_stream_array(indexes, cdfs, cdfs_sizes, offsets).tolist()

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes=None, offsets=None):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)


def _row_ptrs(a2d):
    """ctypes array of the row addresses of a C-contiguous 2-D array (no per-row numpy views / .ctypes objects); the caller
    has checked the element type against the entry point it calls."""
    if a2d.ndim != 2 or a2d.strides[1] != a2d.itemsize:
        raise ValueError("expected a 2-D array with contiguous rows")
    base, stride = a2d.ctypes.data, a2d.strides[0]
    return (ctypes.c_void_p * a2d.shape[0])(*[base + i * stride for i in range(a2d.shape[0])])


class DecodePlan:
    """Repeated decode_batch calls on FIXED buffers (the slice loop of decompress(): 12 calls per sub-batch on the same pinned
    index / symbol buffers): the ctypes argument arrays are built once, each run() is one C call."""

    def __init__(self, decoders, table: RansTable, idx2d, out2d, threads=None):
        count = len(decoders)
        if idx2d.shape != out2d.shape or idx2d.shape[0] != count:
            raise ValueError("index / output buffers must be (len(decoders), n)")
        if out2d.dtype != np.int32 or idx2d.dtype not in (np.int32, np.uint8):
            raise TypeError("indexes must be int32 or uint8, symbols int32")
        self._keep = (decoders, table, idx2d, out2d)
        self._args = ((ctypes.c_void_p * count)(*[d._h for d in decoders]), table._h, count, _row_ptrs(idx2d),
                      (ctypes.c_int64 * count)(*([idx2d.shape[1]] * count)), _row_ptrs(out2d), threads or default_threads())
        self._fn = _C.lib().stf_rans_decode_batch_u8 if idx2d.dtype == np.uint8 else _C.lib().stf_rans_decode_batch

    def run(self):
        _C.check(self._fn(*self._args), "stf_rans_decode_batch")


def encode_rows(table: RansTable, sym2d, idx2d, scratch=None, threads=None):
    """encode_batch for the rows of two (B, n) arrays -- int32 / int32, or the narrow transfer format int16 symbols / uint8
    indexes; `scratch` (a dict kept by the caller) recycles the output buffers between calls (8 n + 64 bytes per image
    otherwise freshly mapped every time).  -> list[bytes]."""
    count, n = sym2d.shape
    if idx2d.shape != sym2d.shape:
        raise ValueError("symbols and indexes differ in shape")
    L = _C.lib()
    if sym2d.dtype == np.int16 and idx2d.dtype == np.uint8:
        fn = L.stf_rans_encode_batch_narrow
    elif sym2d.dtype == np.int32 and idx2d.dtype == np.int32:
        fn = L.stf_rans_encode_batch
    else:
        raise TypeError(f"symbols / indexes must be int32 / int32 or int16 / uint8, got {sym2d.dtype} / {idx2d.dtype}")
    cap = int(L.stf_rans_encode_bound(n))
    key = (count, cap)
    if scratch is None:
        scratch = {}
    if scratch.get("key") != key:
        out = np.empty((count, cap), dtype=np.uint8)
        scratch.update(key=key, out=out, out_ptrs=(ctypes.c_void_p * count)(*[out.ctypes.data + i * cap for i in range(count)]),
                       caps=(ctypes.c_int64 * count)(*([cap] * count)), ns=(ctypes.c_int64 * count)(*([n] * count)),
                       lens=(ctypes.c_int64 * count)())
    rc = fn(table._h, count, _row_ptrs(sym2d), _row_ptrs(idx2d), scratch["ns"], scratch["out_ptrs"],
            scratch["caps"], scratch["lens"], threads or default_threads())
    _C.check(rc, "stf_rans_encode_batch")
    out, lens = scratch["out"], scratch["lens"]
    return [out[i, : lens[i]].tobytes() for i in range(count)]


def decode_batch(decoders, table: RansTable, indexes, outs=None, threads=None):
    """Advance each decoder by len(indexes[i]) symbols in parallel; returns list of int32 arrays."""
    count = len(decoders)
    idxs = [_as_i32(i).reshape(-1) for i in indexes]
    if outs is None:
        outs = [np.empty(i.size, dtype=np.int32) for i in idxs]
    VP, I64 = ctypes.c_void_p * count, ctypes.c_int64 * count
    rc = _C.lib().stf_rans_decode_batch(VP(*[d._h for d in decoders]), table._h, count, VP(*[i.ctypes.data for i in idxs]),
                                        I64(*[i.size for i in idxs]), VP(*[o.ctypes.data for o in outs]),
                                        threads or default_threads())
    _C.check(rc, "stf_rans_decode_batch")
    return outs


def pmf_to_quantized_cdf(pmf, precision: int = 16):
    """compressai._CXX.pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:24-81): list[float] -> list[int]."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = _C.lib().stf_pmf_to_quantized_cdf(p.ctypes.data_as(_C._f32p), p.size, precision,
                                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    _C.check(rc, "stf_pmf_to_quantized_cdf")
    return out.tolist()


class DeviceStreams:
    """The y-strings of one sub-batch staged for the device decoder: all streams back to back in ONE pinned host buffer of
    fixed capacity (so that a CUDA graph can hold its device copy), word offsets / lengths, and the per-stream decoder state.
    `load(strings)` refills the staging buffers (host side only); the H2D copies are issued by the caller's captured graph
    or eagerly through `upload()`."""

    def __init__(self, count, capacity_words, device):
        import torch
        self.count, self.capacity = int(count), int(capacity_words)
        self.words_h = torch.empty(self.capacity, dtype=torch.int32, pin_memory=True)
        self.meta_h = torch.empty((2, self.count), dtype=torch.int64, pin_memory=True)   # [0] offsets, [1] lengths
        self.words = torch.empty(self.capacity, dtype=torch.int32, device=device)
        self.meta = torch.empty((2, self.count), dtype=torch.int64, device=device)
        self.lengths32 = torch.empty(self.count, dtype=torch.int32, device=device)
        self.state_x = torch.zeros(self.count, dtype=torch.int64, device=device)
        self.state_pos = torch.zeros(self.count, dtype=torch.int32, device=device)
        self.status = torch.zeros(self.count, dtype=torch.int32, device=device)

    def fits(self, strings):
        return len(strings) == self.count and sum(len(s) for s in strings) // 4 + 2 * len(strings) <= self.capacity and \
            all(len(s) % 4 == 0 and len(s) >= 8 for s in strings)

    def load(self, strings):
        wh, off = self.words_h.numpy(), 0
        for b, s in enumerate(strings):
            nw = len(s) // 4
            wh[off:off + nw] = np.frombuffer(s, dtype=np.int32)
            self.meta_h[0, b], self.meta_h[1, b] = off, nw
            off += nw
        self.used_words = off

    def upload(self):
        """H2D of the staged streams (stream-ordered on the current stream; capturable)."""
        self.words.copy_(self.words_h, non_blocking=True)
        self.meta.copy_(self.meta_h, non_blocking=True)
        self.lengths32.copy_(self.meta[1])


def decode_device(table: RansTable, ds: DeviceStreams, indexes, symbols_out, first):
    """Advance the sub-batch's decoders by one slice on the device: indexes / symbols_out are (count, n) int32 CUDA tensors
    (row b = stream b).  No host synchronisation; check ds.status after the last slice."""
    from . import ops
    img = table.device_image(indexes.device)
    count, n = indexes.shape
    ops._launch("rans_decode_kernel", 8 * count * n, _C.lib().stf_rans_decode_device, img.data_ptr(), img.numel(),
                ds.words.data_ptr(), ds.meta[0].data_ptr(), ds.lengths32.data_ptr(), ds.state_x.data_ptr(),
                ds.state_pos.data_ptr(), ds.status.data_ptr(), int(bool(first)), indexes.data_ptr(), indexes.stride(0),
                symbols_out.data_ptr(), symbols_out.stride(0), count, n, _C.stream())
