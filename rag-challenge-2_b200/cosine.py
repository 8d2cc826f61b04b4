"""N x N cosine-similarity analysis on the GPU — the arithmetic of
semantic_similarity_analysis/generate_similarity_matrix.py (compute_cosine_similarity_matrix :66-100
and the off-diagonal statistics of generate_heatmap :162-171) without materialising the matrix.

The reference normalises the rows (X / (|X| + 1e-8)), forms S = Xn Xn^T with numpy (160 GB at
N = 200k), forces the diagonal to 1, clips to [-1, 1] and takes mean / median / std / max / min of
the off-diagonal entries.  Here the tiles of S are produced by the tensor-core kernel (3xTF32,
csrc/flat_gemm.cu, STATS epilogue) and reduced on the fly; the median comes from a histogram.
Per-row nearest neighbours (the heat-map's structure) are the batched exact search on Xn.
"""
import ctypes as C

import numpy as np

from . import _lib

N_BINS = 8192


def cosine_stats(rows, col_begin=0, col_end=None, n_bins=N_BINS, finalize=True):
    """rows: float32 CUDA tensor [n, d] (or a FlatIPIndex's adopted tensor).  Returns the statistics
    dictionary of the reference (avg / median / std / max / min similarity) for the pairs (i, j),
    i != j; with a column range, the partial sums of that shard (finalize=False returns raw
    accumulators + histogram for an all-reduce across GPUs)."""
    import torch

    if not (rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and rows.dim() == 2):
        raise ValueError("cosine_stats needs a contiguous float32 CUDA tensor [n, d]")
    n, d = int(rows.shape[0]), int(rows.shape[1])
    col_end = n if col_end is None else int(col_end)
    source = rows
    if d % 32:  # the kernels take widths that are multiples of 32: zero columns change neither norms nor products
        rows = torch.nn.functional.pad(rows, (0, 32 - d % 32)).contiguous()
        d = int(rows.shape[1])
    stats = torch.zeros(6, dtype=torch.float64, device=rows.device)
    hist = torch.zeros(n_bins, dtype=torch.int64, device=rows.device)
    stream = torch.cuda.current_stream(rows.device).cuda_stream
    lib = _lib.load()
    _lib.check(lib.b2r_cosine_stats_dev(C.c_void_p(rows.data_ptr()), n, d, int(col_begin), col_end,
                                        C.c_void_p(stats.data_ptr()), C.c_void_p(hist.data_ptr()), int(n_bins), 0,
                                        None, None, rows.device.index or 0, C.c_void_p(stream)))
    if not finalize:
        return stats, hist
    out = finalize_stats(stats.cpu().numpy(), hist.cpu().numpy())
    if 2 <= n <= EXACT_MEDIAN_MAX_ROWS and int(col_begin) == 0 and col_end == n:
        out["median_similarity"] = exact_median_small(source)
    return out


EXACT_MEDIAN_MAX_ROWS = 1024


def exact_median_small(rows):
    """Report-sized inputs (n <= 1024, the reference's real case: a report has 1e3-1e4 chunks at most and the
    miniature fixtures tens): with so few entries the two middle order statistics np.median averages
    (generate_similarity_matrix.py:166) can lie further apart than a histogram bin, so the median is taken from the
    values themselves — every row's n inner products with the normalised rows through the exact scan (k = n),
    self entries dropped by id, clipped, sorted on the device."""
    import torch

    from .flat_index import FlatIPIndex

    n = int(rows.shape[0])
    xn = (rows / (rows.norm(dim=1, keepdim=True) + 1e-8)).contiguous()
    index = FlatIPIndex.from_device_tensor(xn)
    D, J = index.search_dev(xn, n)
    off = D[J != torch.arange(n, device=rows.device).unsqueeze(1)].clamp_(-1.0, 1.0)
    if off.numel() != n * (n - 1):
        raise RuntimeError("exact_median_small: the k = n search did not return every row once")
    s = off.sort().values
    m = s.numel()  # n (n - 1): always even
    return float((s[m // 2 - 1] + s[m // 2]) / 2)


def finalize_stats(stats, hist):
    """[sum, sumsq, min, max, count] + histogram -> the reference's statistics (generate_similarity_matrix.py:165-171)."""
    total, sumsq, mn, mx, count = (float(stats[i]) for i in range(5))
    if count <= 0:
        raise ValueError("no off-diagonal entries")
    mean = total / count
    var = max(0.0, sumsq / count - mean * mean)  # np.std: population standard deviation
    return {
        "avg_similarity": mean,
        "median_similarity": histogram_median(np.asarray(hist, dtype=np.float64), count),
        "std_similarity": var ** 0.5,
        "max_similarity": mx,
        "min_similarity": mn,
        "count": int(round(count)),
    }


def histogram_median(hist, count):
    """Median of `count` values from equal-width bins over [-1, 1], linear inside the median bin.
    np.median averages the two middle order statistics for an even count; both lie in (or next to)
    the same bin, so the interpolated midpoint is within one bin width (2 / n_bins) of it."""
    n_bins = hist.shape[0]
    width = 2.0 / n_bins
    cum = np.cumsum(hist)
    target = 0.5 * count
    b = int(np.searchsorted(cum, target, side="left"))
    b = min(b, n_bins - 1)
    before = cum[b - 1] if b > 0 else 0.0
    inside = hist[b] if hist[b] > 0 else 1.0
    return -1.0 + width * (b + (target - before) / inside)


def cosine_topk(rows, k, batch=4096):
    """Per-row k nearest neighbours by cosine similarity (self excluded): exact batched search on the
    normalised rows.  Returns CUDA tensors (S float32 [n, k], I int64 [n, k])."""
    import torch

    from .flat_index import FlatIPIndex

    n = rows.shape[0]
    norms = rows.norm(dim=1, keepdim=True)
    xn = (rows / (norms + 1e-8)).contiguous()
    index = FlatIPIndex.from_device_tensor(xn)
    kk = min(k + 1, n)
    S = torch.empty((n, k), dtype=torch.float32, device=rows.device)
    I = torch.empty((n, k), dtype=torch.int64, device=rows.device)
    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        D, J = index.search_dev(xn[lo:hi].contiguous(), kk)
        # drop the row itself (an exact duplicate with a lower id may rank above it: drop by id)
        self_col = (J == torch.arange(lo, hi, device=rows.device).unsqueeze(1))
        has_self = self_col.any(dim=1, keepdim=True)
        drop = torch.where(has_self, self_col.float().argmax(dim=1, keepdim=True),
                           torch.full((hi - lo, 1), kk - 1, device=rows.device, dtype=torch.int64))
        keep = torch.ones_like(J, dtype=torch.bool).scatter_(1, drop, False)
        S[lo:hi, : kk - 1] = D[keep].view(hi - lo, kk - 1)
        I[lo:hi, : kk - 1] = J[keep].view(hi - lo, kk - 1)
        if kk - 1 < k:
            S[lo:hi, kk - 1:] = -3.4028235e38
            I[lo:hi, kk - 1:] = -1
    return S, I


def heatmap_tiles(rows, pixels=512):
    """Down-sampled heat map of the N x N cosine matrix (next-row f4): the reference plots the full matrix with
    Plotly (generate_similarity_matrix.py:124-160), which at N = 200k would be a 160 GB array.  Pixel (a, b) is the
    mean of S over rows of group a and columns of group b (groups = `pixels` contiguous row ranges).  The mean of
    dot products is the dot product of the means, so the map is the Gram matrix of the pooled normalised rows —
    O(N d + pixels^2 d) — corrected for the diagonal the reference forces to 1.  (The clip to [-1, 1] only touches
    values that rounding pushed past +-1; it is applied to the pixel means.)  Returns a float32 CUDA tensor
    [p, p] with p = min(pixels, N), and the group boundaries."""
    import torch

    n, d = int(rows.shape[0]), int(rows.shape[1])
    p = max(1, min(int(pixels), n))
    bounds = torch.div(torch.arange(p + 1, device=rows.device, dtype=torch.int64) * n, p, rounding_mode="floor").contiguous()
    if not (rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous()):
        raise ValueError("heatmap_tiles needs a contiguous float32 CUDA tensor [n, d]")
    # three kernels of this library (csrc/index_build.cu): divisors |x| + 1e-8, float64 pooled sums per group, Gram
    div = torch.empty(n, dtype=torch.float32, device=rows.device)
    pooled = torch.empty((p, d), dtype=torch.float64, device=rows.device)
    self_dot = torch.empty(p, dtype=torch.float64, device=rows.device)
    tiles = torch.empty((p, p), dtype=torch.float32, device=rows.device)
    stream = torch.cuda.current_stream(rows.device).cuda_stream
    with torch.cuda.device(rows.device):
        _lib.check(_lib.load().b2r_cosine_heatmap_dev(C.c_void_p(rows.data_ptr()), n, d, C.c_void_p(bounds.data_ptr()), p,
                                                      C.c_void_p(div.data_ptr()), C.c_void_p(pooled.data_ptr()),
                                                      C.c_void_p(self_dot.data_ptr()), C.c_void_p(tiles.data_ptr()),
                                                      C.c_void_p(stream)))
    return tiles, bounds


class SimilarityMatrixGenerator:
    """Script-level mirror of semantic_similarity_analysis/generate_similarity_matrix.py:19-331
    (`SimilarityMatrixGenerator(vector_db_dir, documents_dir, output_dir)`, `process_document`,
    `process_all_documents` -> `self.stats`, written to `<output_dir>/statistics.json`): the same per-report
    statistics dictionary, computed on the GPU without the N x N matrix.  The Plotly page (:124-238, out of
    scope) is replaced by the down-sampled tile map `<name>_similarity_tiles.npy` (`heatmap_tiles`)."""

    def __init__(self, vector_db_dir, documents_dir, output_dir, device=0, pixels=512):
        from pathlib import Path

        self.vector_db_dir = Path(vector_db_dir)
        self.documents_dir = Path(documents_dir)
        self.output_dir = Path(output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.device = int(device)
        self.pixels = int(pixels)
        self.stats = {}

    def load_vectors_from_faiss(self, faiss_path):
        """:38-64 — all rows of the index, as a float32 CUDA tensor (the reference returns a numpy array)."""
        import torch

        from .flat_index import read_index_rows

        _, rows = read_index_rows(faiss_path)
        return torch.from_numpy(np.ascontiguousarray(rows)).to(torch.device("cuda", self.device))

    @staticmethod
    def load_document_info(doc_path):
        """:102-122"""
        import json

        with open(doc_path, "r", encoding="utf-8") as f:
            doc = json.load(f)
        return {
            "name": doc["metainfo"]["sha1_name"],
            "company": doc["metainfo"].get("company_name", "Unknown"),
            "chunks_count": len(doc["content"]["chunks"]),
            "pages": doc["metainfo"].get("pages_amount", 0),
            "text_blocks": doc["metainfo"].get("text_blocks_amount", 0),
            "tables": doc["metainfo"].get("tables_amount", 0),
        }

    def process_document(self, doc_name):
        """:258-301 — missing files are reported and skipped."""
        faiss_path = self.vector_db_dir / f"{doc_name}.faiss"
        doc_path = self.documents_dir / f"{doc_name}.json"
        if not faiss_path.exists():
            print(f"❌ 未找到 FAISS 文件: {faiss_path}")
            return
        if not doc_path.exists():
            print(f"❌ 未找到文档文件: {doc_path}")
            return
        doc_info = self.load_document_info(doc_path)
        rows = self.load_vectors_from_faiss(faiss_path)
        n = int(rows.shape[0])
        s = cosine_stats(rows)  # raises on n < 2 as numpy's max() of an empty off-diagonal does (:170)
        self.stats[doc_info["name"]] = {
            "chunks_count": n,
            "avg_similarity": s["avg_similarity"], "median_similarity": s["median_similarity"],
            "std_similarity": s["std_similarity"], "max_similarity": s["max_similarity"],
            "min_similarity": s["min_similarity"],
            "company": doc_info["company"], "pages": doc_info["pages"], "text_blocks": doc_info["text_blocks"],
            "tables": doc_info["tables"],
        }
        tiles, _ = heatmap_tiles(rows, self.pixels)
        np.save(self.output_dir / f"{doc_name}_similarity_tiles.npy", tiles.cpu().numpy())

    def process_all_documents(self):
        """:303-349 — every `*.faiss` in sorted order; a failing document is reported and skipped."""
        import json

        for doc_name in [f.stem for f in sorted(self.vector_db_dir.glob("*.faiss"))]:
            try:
                self.process_document(doc_name)
            except Exception as e:  # noqa: BLE001 - the reference catches, prints and continues (:326-331)
                print(f"\n❌ 处理 {doc_name} 时出错: {e}")
        with open(self.output_dir / "statistics.json", "w", encoding="utf-8") as f:
            json.dump(self.stats, f, ensure_ascii=False, indent=2)
        return self.stats
