"""Deterministic synthetic corpora for the benchmark and the parity tests.

The embedding and LLM services the reference calls are offline, and the reference ships no
indexes, so throughput is measured on:
  * dense rows: x_i = g_i / |g_i|, g_i ~ N(0, I_d), float32 (text-embedding-v4 width d = 1024);
    queries: half random unit vectors, half normalize(x_j + 0.1 * noise);
  * token corpus: V = 50,000 terms with Zipf(1.0) ranks, document length
    clip(round(lognormal(ln 120, 0.5)), 8, 512) whitespace tokens "t<rank>";
    queries of 3..12 tokens from the same law, half of them avoiding the 20 most frequent terms.
Generation is chunked by fixed-size blocks of rows / documents with a per-block seed, so the global
corpus is the same no matter how many GPUs it is sharded over.
"""
import math

import numpy as np

ROW_BLOCK = 1 << 14   # rows per seeded block (dense)
DOC_BLOCK = 1 << 14   # documents per seeded block (tokens)


# ------------------------------------------------------------------------------------------------
# dense rows
# ------------------------------------------------------------------------------------------------
def dense_rows_torch(row_lo, row_hi, d, device, seed=1234):
    """float32 CUDA tensor with global rows [row_lo, row_hi).  Blocks are seeded whole, so a shard may start or end
    anywhere: a boundary block is generated in full and sliced (the corpus does not depend on the sharding)."""
    import torch

    out = torch.empty((row_hi - row_lo, d), dtype=torch.float32, device=device)
    for blk_lo in range(row_lo // ROW_BLOCK * ROW_BLOCK, row_hi, ROW_BLOCK):
        lo, hi = max(row_lo, blk_lo), min(row_hi, blk_lo + ROW_BLOCK)
        g = torch.Generator(device=device).manual_seed(seed + blk_lo // ROW_BLOCK)
        blk = torch.randn((ROW_BLOCK, d), generator=g, device=device, dtype=torch.float32)[lo - blk_lo: hi - blk_lo]
        blk /= blk.norm(dim=1, keepdim=True)
        out[lo - row_lo: hi - row_lo] = blk
    return out


def dense_queries_torch(nq, d, device, rows=None, seed=99):
    """Half random unit vectors, half noisy copies of stored rows (when `rows` is given)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn((nq, d), generator=g, device=device, dtype=torch.float32)
    if rows is not None and rows.shape[0] > 0:
        pick = torch.randint(0, rows.shape[0], (nq,), generator=g, device=device)
        noise = torch.randn((nq, d), generator=g, device=device, dtype=torch.float32) * (0.1 / math.sqrt(d))
        near = rows[pick] + noise
        mask = (torch.arange(nq, device=device) % 2 == 0).unsqueeze(1)
        q = torch.where(mask, near, q)
    q /= q.norm(dim=1, keepdim=True)
    return q.contiguous()


def dense_rows_np(n, d, seed=1234):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


# ------------------------------------------------------------------------------------------------
# token corpus
# ------------------------------------------------------------------------------------------------
def zipf_cdf(vocab, s=1.0):
    w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64) ** s
    return np.cumsum(w / w.sum())


def token_corpus_np(n_docs, vocab=50_000, seed=4242, mean_len=120.0, sigma=0.5):
    """Small corpora for the CPU oracle: returns a list of token lists (strings "t<rank>")."""
    rng = np.random.default_rng(seed)
    cdf = zipf_cdf(vocab)
    lens = np.clip(np.rint(rng.lognormal(math.log(mean_len), sigma, n_docs)), 8, 512).astype(np.int64)
    docs = []
    for L in lens:
        ranks = np.searchsorted(cdf, rng.random(int(L)), side="left")
        docs.append([f"t{int(r)}" for r in np.minimum(ranks, vocab - 1)])
    return docs


def token_queries_np(nq, vocab=50_000, seed=4243):
    rng = np.random.default_rng(seed)
    cdf = zipf_cdf(vocab)
    out = []
    for i in range(nq):
        L = int(rng.integers(3, 13))
        toks = []
        while len(toks) < L:
            r = int(min(np.searchsorted(cdf, rng.random(), side="left"), vocab - 1))
            if i % 2 == 1 and r < 20:
                continue
            toks.append(r)
        out.append(toks)
    return out


def token_postings_torch(doc_lo, doc_hi, vocab, device, seed=4242, mean_len=120.0, sigma=0.5):
    """Term-major CSR postings of global documents [doc_lo, doc_hi), built in HBM.

    Returns dict(indptr int64[V+1], doc_ids int32[P] (LOCAL ids, ascending per term), tf int32[P],
    doc_len int32[n], df int64[V], total_len int).  Blocks of DOC_BLOCK documents are seeded whole; a shard that
    starts or ends inside a block generates that block in full and keeps its own documents, so the corpus does not
    depend on the sharding (the corpus end, `doc_hi` of the last shard, truncates the last block as before).
    """
    import torch

    n = doc_hi - doc_lo
    cdf = torch.from_numpy(zipf_cdf(vocab)).to(device)
    keys = []
    lens_all = []
    for blk_lo in range(doc_lo // DOC_BLOCK * DOC_BLOCK, doc_hi, DOC_BLOCK):
        lo, hi = max(doc_lo, blk_lo), min(doc_hi, blk_lo + DOC_BLOCK)
        g = torch.Generator(device=device).manual_seed(seed + blk_lo // DOC_BLOCK)
        z = torch.randn((DOC_BLOCK,), generator=g, device=device, dtype=torch.float64)[: hi - blk_lo]
        lens = torch.clamp(torch.round(torch.exp(math.log(mean_len) + sigma * z)), 8, 512).to(torch.int64)
        total = int(lens.sum().item())
        skip = int(lens[: lo - blk_lo].sum().item()) if lo > blk_lo else 0   # tokens of the block's documents before `lo`
        u = torch.rand((int(DOC_BLOCK) * 512,), generator=g, device=device, dtype=torch.float64)[skip:total]
        lens = lens[lo - blk_lo:]
        ranks = torch.clamp(torch.searchsorted(cdf, u), max=vocab - 1)
        doc_of = torch.repeat_interleave(torch.arange(lo - doc_lo, hi - doc_lo, device=device), lens)
        keys.append(ranks * n + doc_of)
        lens_all.append(lens)
        del u, ranks, doc_of, z
    key = torch.cat(keys)
    del keys
    key, _ = torch.sort(key)
    uniq, counts = torch.unique_consecutive(key, return_counts=True)
    del key
    term = uniq // n
    doc_ids = (uniq - term * n).to(torch.int32)
    del uniq
    df = torch.bincount(term, minlength=vocab)
    indptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(df, 0)
    doc_len = torch.cat(lens_all).to(torch.int32)
    return {
        "indptr": indptr,
        "doc_ids": doc_ids.contiguous(),
        "tf": counts.to(torch.int32).contiguous(),
        "doc_len": doc_len.contiguous(),
        "df": df,
        "total_len": int(doc_len.sum().item()),
    }


def bm25_idf_from_df(df, corpus_size, epsilon=0.25):
    """rank_bm25's idf (epsilon floor included) from document frequencies; terms with df == 0 are not
    in rank_bm25's vocabulary: they get idf 0 here and do not enter the average.  numpy float64 —
    used for the synthetic corpora where the vocabulary is an integer range."""
    df = np.asarray(df, dtype=np.float64)
    present = df > 0
    idf = np.zeros_like(df)
    idf[present] = np.log(corpus_size - df[present] + 0.5) - np.log(df[present] + 0.5)
    average_idf = idf[present].sum() / max(1, int(present.sum()))
    eps = epsilon * average_idf
    idf[present & (idf < 0)] = eps
    return idf, average_idf


# ------------------------------------------------------------------------------------------------
# synthetic year databases (BASELINE config 2: J2020..J2025, one per GPU)
# ------------------------------------------------------------------------------------------------
class _LazySeq:
    """Sequence of n records produced on demand (a 2M-chunk report as 2M dicts would cost GBs of host RAM)."""

    def __init__(self, n, make):
        self._n, self._make = int(n), make

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._make(i)

    def __iter__(self):
        return (self._make(i) for i in range(self._n))


class _LazyPages(dict):
    """page number -> page record, materialised on first use (what retrieval._page_text_lookup returns)."""

    def __init__(self, n_pages, make):
        super().__init__()
        self._n, self._make = int(n_pages), make

    def __missing__(self, page_no):
        if not 1 <= int(page_no) <= self._n:
            raise KeyError(page_no)
        rec = self[page_no] = self._make(int(page_no))
        return rec

    def __contains__(self, page_no):
        try:
            return 1 <= page_no.__index__() <= self._n   # Python and numpy integers alike
        except AttributeError:
            return False


def synthetic_year_report(name, company, n_chunks, vector_db, chunks_per_page=8, materialise=False):
    """A report dict {"name", "vector_db", "document"} in the chunked-report schema whose
    chunk / page records are generated on demand: chunk i sits on page i // chunks_per_page + 1.
    `materialise=True` builds the real lists a report loaded from JSON holds (~0.4 KB of host RAM per chunk) —
    what a host-path timing should run on: the on-demand records cost a Python call + an f-string per hit."""
    n_pages = (n_chunks + chunks_per_page - 1) // chunks_per_page

    def chunk(i):
        return {"page": i // chunks_per_page + 1, "length_tokens": 300, "text": f"{name} chunk {i}", "id": i,
                "type": "content"}

    def page(p):
        return {"page": p, "text": f"{name} page {p}"}

    if materialise:
        import gc
        was = gc.isenabled()
        gc.disable()   # millions of acyclic dicts: the collector's generation passes would dominate the build
        try:
            document = {"metainfo": {"sha1_name": name, "company_name": company, "pages_amount": n_pages},
                        "content": {"pages": [page(j + 1) for j in range(n_pages)],
                                    "chunks": [chunk(i) for i in range(n_chunks)]}}
        finally:
            if was:
                gc.enable()
        return {"name": name, "vector_db": vector_db, "document": document}
    document = {"metainfo": {"sha1_name": name, "company_name": company, "pages_amount": n_pages},
                "content": {"pages": _LazySeq(n_pages, lambda j: page(j + 1)), "chunks": _LazySeq(n_chunks, chunk)}}
    return {"name": name, "vector_db": vector_db, "document": document,
            "_page_lookup": (id(document["content"]["pages"]), _LazyPages(n_pages, page))}
