"""FlatIPIndex — the faiss.IndexFlatIP surface the reference uses, backed by libb2r (sm_100a).

Reference call sites this object serves (paths relative to the reference checkout):
  faiss.IndexFlatIP(d) / index.add            src/ingestion.py:82-83
  faiss.read_index / faiss.write_index        src/retrieval.py:524, src/ingestion.py:102
  index.search(x=..., k=...)                  src/retrieval.py:622,730,1140,1165,1261,1423
  index.reconstruct / reconstruct_batch       src/retrieval.py:573, generate_similarity_matrix.py:61
  index.ntotal / index.d                      generate_similarity_matrix.py:49-51

Semantics follow faiss-cpu 1.9.0.post1: rows are stored verbatim as fp32, ids are insertion order,
search returns (D float32[nq,k], I int64[nq,k]) sorted by score descending (ties: ascending id),
k > ntotal pads with I=-1, D=-3.4028235e38.  All arithmetic runs on the GPU; there is no CPU path.
"""
import ctypes as C
import struct
import threading

import numpy as np

from . import _lib

METRIC_INNER_PRODUCT = 0


class FlatIPIndex:
    """Exact inner-product index resident in HBM (one handle == one row shard on one GPU)."""

    def __init__(self, d, device=0):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        self.d = int(d)
        self.device = int(device)
        self.is_trained = True
        self.metric_type = METRIC_INNER_PRODUCT
        self._borrowed = None  # keeps an adopted torch tensor alive
        self._lock = threading.Lock()
        _lib.check(self._lib.b2r_flat_create(self.d, self.device, C.byref(self._h)))

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_device_tensor(cls, rows):
        """Zero-copy: adopt a contiguous float32 CUDA tensor [n, d] (the tensor is kept alive)."""
        if not (rows.is_cuda and rows.is_contiguous() and rows.dim() == 2 and str(rows.dtype) == "torch.float32"):
            raise ValueError("from_device_tensor needs a contiguous float32 CUDA tensor [n, d]")
        idx = cls(rows.shape[1], device=rows.device.index or 0)
        _lib.check(idx._lib.b2r_flat_adopt(idx._h, C.c_void_p(rows.data_ptr()), int(rows.shape[0])))
        idx._borrowed = rows
        return idx

    def add(self, x):
        """index.add(x): append float32 rows [n, d] (numpy array, or a CUDA tensor on this device)."""
        if hasattr(x, "is_cuda"):
            if not x.is_cuda:
                x = x.numpy()
            else:
                if x.dim() != 2 or x.shape[1] != self.d:
                    raise AssertionError(f"add: expected [n, {self.d}] rows")
                x = x.contiguous().float()
                with self._lock:
                    _lib.check(self._lib.b2r_flat_add(self._h, C.c_void_p(x.data_ptr()), int(x.shape[0]), 1))
                return
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError(f"add: expected [n, {self.d}] rows, got {x.shape}")
        with self._lock:
            _lib.check(self._lib.b2r_flat_add(self._h, _lib.ptr(x), int(x.shape[0]), 0))

    # ---- properties ---------------------------------------------------------------------------
    @property
    def ntotal(self):
        v = C.c_int64(0)
        _lib.check(self._lib.b2r_flat_ntotal(self._h, C.byref(v)))
        return int(v.value)

    @property
    def rows_ptr(self):
        p = C.c_void_p()
        _lib.check(self._lib.b2r_flat_rows_dev(self._h, C.byref(p)))
        return p.value or 0

    # ---- search -------------------------------------------------------------------------------
    def search(self, x, k):
        """D, I = index.search(x, k) with host (numpy) buffers; thread-safe like faiss."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError(f"search: query width {x.shape} does not match index d={self.d}")
        k = int(k)
        if k < 1:
            raise AssertionError("search: k must be >= 1")
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        _lib.check(self._lib.b2r_flat_search_host(self._h, _lib.ptr(x), nq, k, _lib.ptr(D), _lib.ptr(I)))
        return D, I

    def search_dev(self, q, k, D=None, I=None, id_offset=0, stream=None, overlap_previous=False):
        """Resident search: q is a float32 CUDA tensor [nq, d]; returns CUDA tensors (D, I).

        Enqueued on `stream` (default: torch's current stream); nothing is synchronised.
        overlap_previous: B2R_LAUNCH_OVERLAP_PREVIOUS — the scan (one or two queries) may start while the previous
        kernel of the stream is finishing; q must have been complete when that kernel started (see include/b2r.h).
        """
        import torch

        if q.dim() == 1:
            q = q.reshape(1, -1)
        if not (q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.shape[1] == self.d):
            raise AssertionError(f"search_dev: need a contiguous float32 CUDA tensor [nq, {self.d}]")
        nq, k = int(q.shape[0]), int(k)
        if D is None:
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        if I is None:
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        if stream is None:
            stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(
            self._lib.b2r_flat_search_dev_ex(
                self._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                int(id_offset), C.c_void_p(stream), 1 if overlap_previous else 0,
            )
        )
        return D, I

    def enable_bf16_candidates(self, enable=True):
        """Keep a bf16 copy of the rows for the candidate pass of batches of more than 128 queries (exact results:
        fp32 rescoring + completeness proof, see b2r_flat_set_candidate_bf16)."""
        _lib.check(self._lib.b2r_flat_set_candidate_bf16(self._h, 1 if enable else 0))

    def last_launch_count(self):
        return int(self._lib.b2r_last_launch_count())

    # ---- reconstruct --------------------------------------------------------------------------
    def reconstruct_batch(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        out = np.empty((ids.shape[0], self.d), dtype=np.float32)
        _lib.check(self._lib.b2r_flat_reconstruct(self._h, _lib.ptr(ids), int(ids.shape[0]), _lib.ptr(out)))
        return out

    def reconstruct(self, i):
        return self.reconstruct_batch(np.array([int(i)], dtype=np.int64))[0]

    def reconstruct_n(self, i0=0, n=-1):
        n = self.ntotal - i0 if n < 0 else n
        return self.reconstruct_batch(np.arange(i0, i0 + n, dtype=np.int64))

    # ---- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b2r_flat_destroy(self._h)
            self._h = C.c_void_p()
            self._borrowed = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# On-disk format of faiss.write_index for IndexFlatIP ("IxFI").  Layout as written by faiss
# (index_write.cpp: write_index_header + WRITEXBVECTOR), restated from the published source of
# faiss 1.9.0 — no .faiss file ships with the reference, so this is unverified against a real file:
#   char[4] "IxFI" | i32 d | i64 ntotal | i64 dummy(1<<20) | i64 dummy(1<<20) | u8 is_trained |
#   i32 metric_type (0 = inner product) | u64 n_floats (= ntotal*d) | float32[n_floats]
# ------------------------------------------------------------------------------------------------
_HDR = struct.Struct("<4siqqqBi")


def write_index(index, path):
    """faiss.write_index(index, path)                                   src/ingestion.py:102"""
    rows = index.reconstruct_n(0, index.ntotal) if index.ntotal else np.zeros((0, index.d), np.float32)
    with open(str(path), "wb") as f:
        f.write(_HDR.pack(b"IxFI", index.d, index.ntotal, 1 << 20, 1 << 20, 1, METRIC_INNER_PRODUCT))
        f.write(struct.pack("<Q", rows.size))
        f.write(rows.astype("<f4", copy=False).tobytes())


def read_index_rows(path):
    """Parse an IxFI file into (d, float32[n, d]) without touching the GPU."""
    with open(str(path), "rb") as f:
        hdr = f.read(_HDR.size)
        if len(hdr) != _HDR.size:
            raise RuntimeError(f"{path}: truncated faiss index header")
        fourcc, d, ntotal, _, _, _, metric = _HDR.unpack(hdr)
        if fourcc not in (b"IxFI", b"IxFl"):
            raise RuntimeError(f"{path}: unsupported faiss index type {fourcc!r} (only IndexFlatIP)")
        if metric != METRIC_INNER_PRODUCT:
            raise RuntimeError(f"{path}: metric_type {metric} is not inner product")
        (n_floats,) = struct.unpack("<Q", f.read(8))
        if n_floats != ntotal * d:
            raise RuntimeError(f"{path}: vector count {n_floats} != ntotal*d = {ntotal * d}")
        rows = np.fromfile(f, dtype="<f4", count=n_floats)
        if rows.size != n_floats:
            raise RuntimeError(f"{path}: truncated faiss index payload")
    return d, rows.reshape(ntotal, d)


def read_index(path, device=0):
    """faiss.read_index(path)                                           src/retrieval.py:524"""
    d, rows = read_index_rows(path)
    index = FlatIPIndex(d, device=device)
    if rows.shape[0]:
        index.add(rows)
    return index


IndexFlatIP = FlatIPIndex
