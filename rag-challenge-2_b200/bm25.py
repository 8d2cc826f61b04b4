"""BM25Okapi — the rank_bm25.BM25Okapi surface the reference uses, scored on the GPU by libb2r.

Reference call sites (paths relative to the reference checkout):
  BM25Okapi([chunk.split() for chunk in chunks])          src/ingestion.py:19-22
  pickle.dump / pickle.load of the index object           src/ingestion.py:46-47, src/retrieval.py:253-254
  bm25_index.get_scores(query.split())                    src/retrieval.py:261-262
  sorted(range(n), key=scores.__getitem__, reverse=True)  src/retrieval.py:265   (-> topk())

Host side (this file): tokenised corpus -> vocabulary, term-major CSR postings, and the corpus
statistics exactly as rank-bm25 0.2.2 computes them (idf with the epsilon floor, avgdl; Python
float arithmetic in the same order).  Device side (csrc/bm25.cu): per-posting weights, scoring and
the fused top-k.  There is no CPU scoring path.
"""
import ctypes as C
import math
import pickle

import numpy as np

from . import _lib


def corpus_statistics(nd, corpus_size, epsilon=0.25):
    """rank_bm25.BM25Okapi._calc_idf: nd maps word -> number of documents containing it (dict order
    is preserved because the average is a running Python float sum)."""
    idf = {}
    idf_sum = 0
    negative = []
    for word, freq in nd.items():
        v = math.log(corpus_size - freq + 0.5) - math.log(freq + 0.5)
        idf[word] = v
        idf_sum += v
        if v < 0:
            negative.append(word)
    average_idf = idf_sum / len(idf)
    eps = epsilon * average_idf
    for word in negative:
        idf[word] = eps
    return idf, average_idf


def _postings_from_tokens_cuda(tok, offs, n_terms, pos_offset):
    """CSR postings of a tokenised corpus with this library's kernels (csrc/index_build.cu): per-document warp sort ->
    unique (term, tf) pairs, df / first positions / per-block counts by atomics, block offsets by a scan, placement in
    document order.  tok int32 CUDA [total], offs int64 CUDA [n_docs + 1].  Returns None when a document is longer than
    the warp sort takes (1024 tokens): the caller then uses the torch sort path."""
    import torch

    lib = _lib.load()
    dev = tok.device
    n_docs = int(offs.shape[0]) - 1
    total = int(tok.shape[0])
    n_blocks = (n_docs + 4095) // 4096
    pair_term = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    pair_tf = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    n_unique = torch.empty(n_docs, dtype=torch.int32, device=dev)
    doc_len = torch.empty(n_docs, dtype=torch.int32, device=dev)
    df = torch.zeros(n_terms, dtype=torch.int64, device=dev)
    first = torch.full((n_terms,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
    cell = torch.zeros(n_terms * (n_blocks + 1), dtype=torch.int32, device=dev)
    indptr = torch.empty(n_terms + 1, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(lib.b2r_bm25_build_count_dev(p(tok), p(offs), n_docs, int(n_terms), int(pos_offset), p(pair_term), p(pair_tf),
                                                p(n_unique), p(doc_len), p(df), p(first), p(cell), p(indptr), p(err), stream))
        code = int(err.item())
        if code == 1:
            return None
        if code == 2:
            raise ValueError("token id outside [0, n_terms)")
        n_post = int(indptr[-1].item())
        doc_ids = torch.empty(max(n_post, 1), dtype=torch.int32, device=dev)
        tf = torch.empty(max(n_post, 1), dtype=torch.int32, device=dev)
        cursor = torch.zeros_like(cell)
        _lib.check(lib.b2r_bm25_build_place_dev(p(offs), n_docs, int(n_terms), p(pair_term), p(pair_tf), p(n_unique), p(indptr),
                                                p(cell), p(cursor), p(doc_ids), p(tf), stream))
    return {"indptr": indptr, "doc_ids": doc_ids[:n_post], "tf": tf[:n_post], "doc_len": doc_len, "df": df, "first": first}


def token_stream_statistics(token_ids, doc_offsets, n_terms, epsilon=0.25, device=None, group=None, use_kernels=True):
    """Tokenised corpus -> CSR postings + rank_bm25 statistics on `device`.  On CUDA the postings come from this library's
    build kernels (_postings_from_tokens_cuda); the torch sort / unique formulation below is the CPU path of the gloo
    plumbing test and the fallback for documents longer than 1024 tokens (use_kernels=False forces it: the tests compare
    the two).  Returns dict(indptr, doc_ids, tf, doc_len tensors; idf float64 numpy [n_terms] (0 for terms that never
    occur); avgdl; average_idf; corpus_size)."""
    import torch

    tok = torch.as_tensor(token_ids).to(device=device).reshape(-1)
    offs = torch.as_tensor(doc_offsets).to(device=device, dtype=torch.int64).reshape(-1).contiguous()
    n_docs = int(offs.shape[0]) - 1
    totals = torch.tensor([int(tok.shape[0]), n_docs], dtype=torch.int64, device=tok.device)
    pos_offset = 0
    if group is not None:
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        sizes = [torch.zeros(2, dtype=torch.int64, device=tok.device) for _ in range(world)]
        dist.all_gather(sizes, totals, group=group)
        pos_offset = sum(int(sz[0]) for sz in sizes[:rank])   # documents are sharded in rank order
        totals = torch.stack(sizes).sum(0)
    built = None
    if use_kernels and tok.is_cuda and n_docs > 0:
        built = _postings_from_tokens_cuda(tok.to(torch.int32).contiguous(), offs, n_terms, pos_offset)
    if built is not None:
        indptr, doc_ids, counts, lens, df, first = (built[k] for k in ("indptr", "doc_ids", "tf", "doc_len", "df", "first"))
    else:
        tok = tok.to(torch.int64)
        lens = offs[1:] - offs[:-1]
        doc_of = torch.repeat_interleave(torch.arange(n_docs, device=tok.device), lens)
        key, _ = torch.sort(tok * max(n_docs, 1) + doc_of)
        uniq, counts = torch.unique_consecutive(key, return_counts=True)
        term = uniq // max(n_docs, 1)
        doc_ids = (uniq - term * max(n_docs, 1)).to(torch.int32)
        df = torch.bincount(term, minlength=n_terms)
        indptr = torch.zeros(n_terms + 1, dtype=torch.int64, device=tok.device)
        indptr[1:] = torch.cumsum(df, 0)
        # rank_bm25 sums the idf values in first-seen (dict insertion) order: global position of each term's first token
        first = torch.full((n_terms,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=tok.device)
        pos = torch.arange(tok.shape[0], dtype=torch.int64, device=tok.device) + pos_offset
        first.scatter_reduce_(0, tok, pos, reduce="amin", include_self=True)
    if group is not None:
        dist.all_reduce(df, group=group)
        dist.all_reduce(first, op=dist.ReduceOp.MIN, group=group)
    corpus_size = int(totals[1])
    df_host = df.cpu().numpy()
    order = np.argsort(first.cpu().numpy(), kind="stable")
    nd = {int(t): int(df_host[t]) for t in order if df_host[t] > 0}
    idf_map, average_idf = corpus_statistics(nd, corpus_size, epsilon)
    idf = np.zeros(n_terms, dtype=np.float64)
    for t, v in idf_map.items():
        idf[t] = v
    return {"indptr": indptr, "doc_ids": doc_ids.contiguous(), "tf": counts.to(torch.int32).contiguous(),
            "doc_len": lens.to(torch.int32).contiguous(), "idf": idf, "avgdl": int(totals[0]) / corpus_size,
            "average_idf": average_idf, "corpus_size": corpus_size}


class BM25Okapi:
    """Drop-in for rank_bm25.BM25Okapi; `device` selects the GPU that holds the postings."""

    def __init__(self, corpus, tokenizer=None, k1=1.5, b=0.75, epsilon=0.25, device=0, weight_bits=64):
        # weight_bits: 64 (default) keeps rank_bm25's float64 arithmetic bit for bit; 32 is the explicit opt-in that
        # stores the per-posting weights in float32 (8 B / posting, scores within 1e-6 relative: the 10M-chunk bench)
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.tokenizer = tokenizer
        self.device = int(device)
        self.weight_bits = int(weight_bits)
        self._h = None
        self._vocab = None
        if tokenizer is not None:
            corpus = [tokenizer(doc) for doc in corpus]
        self.corpus_size = 0
        self.doc_len = []
        self.doc_freqs = []
        nd = {}
        num_tokens = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_tokens += len(document)
            freqs = {}
            for word in document:
                freqs[word] = freqs.get(word, 0) + 1
            self.doc_freqs.append(freqs)
            for word in freqs:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = num_tokens / self.corpus_size  # ZeroDivisionError on an empty corpus, like rank_bm25
        self.idf, self.average_idf = corpus_statistics(nd, self.corpus_size, epsilon)
        self._build_from_doc_freqs()

    # ---- alternative constructors ---------------------------------------------------------------
    @classmethod
    def from_postings(cls, indptr, doc_ids, tf, doc_len, idf, avgdl, k1=1.5, b=0.75, device=0, weight_bits=64,
                      vocab=None):
        """Build from term-major CSR postings (numpy arrays, or CUDA tensors for corpora generated in
        HBM).  `idf` / `avgdl` are the corpus-global statistics, so a document shard scores exactly
        like the unsharded corpus.  vocab: optional dict token -> term id for string queries."""
        self = cls.__new__(cls)
        self.k1, self.b, self.epsilon = k1, b, None
        self.tokenizer = None
        self.device = int(device)
        self.weight_bits = int(weight_bits)
        self.doc_freqs = None
        self.idf = None
        self.average_idf = None
        self.avgdl = float(avgdl)
        self._vocab = vocab
        self._h = None
        on_dev = hasattr(indptr, "is_cuda")
        if on_dev:
            import torch

            args = [indptr.to(torch.int64).contiguous(), doc_ids.to(torch.int32).contiguous(),
                    tf.to(torch.int32).contiguous(), doc_len.to(torch.int32).contiguous(),
                    idf.to(torch.float64).contiguous()]
            self.corpus_size = int(doc_len.shape[0])
            n_terms = int(indptr.shape[0]) - 1
            ptrs = [C.c_void_p(a.data_ptr()) for a in args]
            torch.cuda.synchronize(indptr.device)
        else:
            args = [np.ascontiguousarray(indptr, np.int64), np.ascontiguousarray(doc_ids, np.int32),
                    np.ascontiguousarray(tf, np.int32), np.ascontiguousarray(doc_len, np.int32),
                    np.ascontiguousarray(idf, np.float64)]
            self.corpus_size = int(args[3].shape[0])
            n_terms = int(args[0].shape[0]) - 1
            ptrs = [_lib.ptr(a) for a in args]
        self.doc_len = None
        self._n_terms = n_terms
        self._create(n_terms, ptrs, 1 if on_dev else 0)
        return self

    @classmethod
    def from_token_ids(cls, token_ids, doc_offsets, n_terms, k1=1.5, b=0.75, epsilon=0.25, device=0, weight_bits=64,
                       vocab=None, group=None):
        """Index build on the GPU from a tokenised corpus (next-row f4): `token_ids` int32[total] are the term ids
        of all documents back to back, `doc_offsets` int64[n_docs + 1] their boundaries (numpy arrays or tensors).
        The (term, doc) pairs are sorted and run-length counted on the device, which yields the term-major CSR
        postings, df and doc_len; idf / avgdl follow rank_bm25 (src/ingestion.py:19-22 builds
        BM25Okapi([chunk.split() ...]) with Python dictionaries).  With `group` (torch.distributed) the documents
        are this rank's shard: df, the token count, the document count and the terms' first-seen order are
        all-reduced so every shard scores exactly like the unsharded corpus."""
        import torch

        dev = torch.device("cuda", int(device))
        stats = token_stream_statistics(token_ids, doc_offsets, n_terms, epsilon=epsilon, device=dev, group=group)
        self = cls.from_postings(stats["indptr"], stats["doc_ids"], stats["tf"], stats["doc_len"],
                                 torch.from_numpy(stats["idf"]).to(dev), stats["avgdl"], k1=k1, b=b, device=int(device),
                                 weight_bits=weight_bits, vocab=vocab)
        self.epsilon = epsilon
        self.average_idf = stats["average_idf"]
        return self

    def _create(self, n_terms, ptrs, on_dev):
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.b2r_bm25_create_ex(self.corpus_size, n_terms, ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                                          float(self.avgdl), ptrs[4], float(self.k1), float(self.b), self.device,
                                          on_dev, self.weight_bits, C.byref(h)))
        self._h = h
        self._lib = lib

    def _build_from_doc_freqs(self):
        """doc_freqs (list of {token: tf}) -> vocabulary + CSR postings -> device index."""
        vocab = {w: i for i, w in enumerate(self.idf)}  # first-seen order, as rank_bm25's nd dict
        n_post = sum(len(f) for f in self.doc_freqs)
        terms = np.empty(n_post, dtype=np.int32)
        docs = np.empty(n_post, dtype=np.int32)
        tfs = np.empty(n_post, dtype=np.int32)
        p = 0
        for doc_id, freqs in enumerate(self.doc_freqs):
            for w, f in freqs.items():
                terms[p] = vocab[w]
                docs[p] = doc_id
                tfs[p] = f
                p += 1
        order = np.argsort(terms, kind="stable")  # stable: doc ids stay ascending inside a term
        n_terms = len(vocab)
        indptr = np.zeros(n_terms + 1, dtype=np.int64)
        np.cumsum(np.bincount(terms, minlength=n_terms), out=indptr[1:])
        idf_arr = np.array([self.idf[w] for w in vocab], dtype=np.float64)
        args = [indptr, np.ascontiguousarray(docs[order]), np.ascontiguousarray(tfs[order]),
                np.asarray(self.doc_len, dtype=np.int32), idf_arr]
        self._vocab = vocab
        self._n_terms = n_terms
        self._create(n_terms, [_lib.ptr(a) for a in args], 0)

    # ---- queries ----------------------------------------------------------------------------------
    def _term_ids(self, query):
        """Tokens -> term ids; unknown tokens contribute 0 in rank_bm25 and are dropped here."""
        if len(query) and not isinstance(query[0], str):
            return np.ascontiguousarray(query, dtype=np.int32)
        if self._vocab is None:
            raise ValueError("this index was built from raw postings without a vocabulary; pass term ids")
        v = self._vocab
        return np.array([v[t] for t in query if t in v], dtype=np.int32)

    def get_scores(self, query):
        """float64[corpus_size] — rank_bm25.BM25Okapi.get_scores(query tokens)."""
        self._ensure_built()
        ids = self._term_ids(query)
        out = np.empty(self.corpus_size, dtype=np.float64)
        _lib.check(self._lib.b2r_bm25_scores_host(self._h, _lib.ptr(ids), int(ids.shape[0]), _lib.ptr(out)))
        return out

    def topk(self, queries, k):
        """Fused scoring + top-k for a batch: queries is a list of token lists (or term-id arrays).
        Returns (S float64[nq,k], I int64[nq,k]) ordered like the reference's stable descending sort
        (src/retrieval.py:265); k > corpus_size pads with id -1."""
        self._ensure_built()
        ids = [self._term_ids(q) for q in queries]
        offs = np.zeros(len(ids) + 1, dtype=np.int32)
        np.cumsum([len(a) for a in ids], out=offs[1:])
        flat = np.ascontiguousarray(np.concatenate(ids) if ids else np.zeros(0), dtype=np.int32)
        nq = len(ids)
        S = np.empty((nq, int(k)), dtype=np.float64)
        I = np.empty((nq, int(k)), dtype=np.int64)
        _lib.check(self._lib.b2r_bm25_topk_host(self._h, _lib.ptr(flat), _lib.ptr(offs), nq, int(k), _lib.ptr(S),
                                                _lib.ptr(I)))
        return S, I

    def topk_dev(self, term_ids, term_offsets, k, S=None, I=None, id_offset=0, stream=None, overlap_previous=False):
        """Resident batch: term_ids / term_offsets are int32 CUDA tensors; returns CUDA tensors.
        overlap_previous: launch beside the previous kernel of the stream (B2R_LAUNCH_OVERLAP_PREVIOUS)."""
        import torch

        nq = int(term_offsets.shape[0]) - 1
        if S is None:
            S = torch.empty((nq, k), dtype=torch.float64, device=term_ids.device)
        if I is None:
            I = torch.empty((nq, k), dtype=torch.int64, device=term_ids.device)
        if stream is None:
            stream = torch.cuda.current_stream(term_ids.device).cuda_stream
        _lib.check(self._lib.b2r_bm25_topk_dev_ex(self._h, C.c_void_p(term_ids.data_ptr()),
                                                  C.c_void_p(term_offsets.data_ptr()), None, nq, int(k),
                                                  C.c_void_p(S.data_ptr()), C.c_void_p(I.data_ptr()), int(id_offset),
                                                  C.c_void_p(stream), 1 if overlap_previous else 0))
        return S, I

    def get_top_n(self, query, documents, n=5):
        """rank_bm25.BM25.get_top_n: the n best documents (numpy argsort order upstream; here the
        deterministic score-descending / id-ascending order of topk)."""
        assert self.corpus_size == len(documents), "The documents given don't match the index corpus!"
        _, I = self.topk([query], min(n, self.corpus_size))
        return [documents[i] for i in I[0] if i >= 0]

    def n_postings(self):
        self._ensure_built()
        n, bits = C.c_int64(0), C.c_int(0)
        _lib.check(self._lib.b2r_bm25_postings(self._h, C.byref(n), C.byref(bits)))
        return int(n.value), int(bits.value)

    def last_launch_count(self):
        return int(self._lib.b2r_last_launch_count())

    # ---- pickling (src/ingestion.py:46-47, src/retrieval.py:253-254) -----------------------------
    _STATE_KEYS = ("corpus_size", "avgdl", "doc_freqs", "idf", "doc_len", "tokenizer", "k1", "b", "epsilon",
                   "average_idf")

    def __getstate__(self):
        if self.doc_freqs is None:
            raise pickle.PicklingError("an index built from raw postings cannot be pickled")
        state = {k: getattr(self, k) for k in self._STATE_KEYS}
        state["weight_bits"] = self.weight_bits  # ours only; rank_bm25's own pickles simply lack it
        return state

    def __setstate__(self, state):
        # accepts the attribute dict of a genuine rank_bm25.BM25Okapi pickle as well as our own
        self.__dict__.update(state)
        self.device = getattr(self, "device", 0)
        self.weight_bits = getattr(self, "weight_bits", 64)
        self._h = None
        self._vocab = None

    def _ensure_built(self):
        if self._h is None:
            self._build_from_doc_freqs()

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.b2r_bm25_destroy(h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Rank25Unpickler(pickle.Unpickler):
    """Loads bm25_dbs/<sha1>.pkl files written by the reference (class rank_bm25.BM25Okapi) into
    this module's BM25Okapi without rank_bm25 being installed."""

    def find_class(self, module, name):
        if module.split(".")[0] == "rank_bm25" and name in ("BM25Okapi", "BM25"):
            return BM25Okapi
        return super().find_class(module, name)


def load_bm25_pickle(fileobj_or_path, device=0):
    if hasattr(fileobj_or_path, "read"):
        obj = _Rank25Unpickler(fileobj_or_path).load()
    else:
        with open(str(fileobj_or_path), "rb") as f:
            obj = _Rank25Unpickler(f).load()
    if not isinstance(obj, BM25Okapi):
        raise TypeError(f"pickle holds {type(obj)!r}, not a BM25Okapi index")
    obj.device = device
    return obj
