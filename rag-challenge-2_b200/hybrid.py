"""Row-sharded exact search, top-k merge and the dense + BM25 hybrid top-k (BASELINE configs 3/4).

One process per GPU (torch.distributed, NCCL over NVLink): rank r holds rows / documents
[bounds[r], bounds[r+1]) of the corpus.  A search is: local scan with global ids -> all-gather of
the per-rank [nq, k] lists (360 B per query per rank at k = 30) -> merge kernel on every rank, so
all ranks end with the same answer.  With one rank the gather and merge are skipped.
The union of per-shard top-k contains the global top-k, so the result is exact.

Replaces, for a corpus that does not fit the reference's one-faiss-index-per-report design, the
Python merge of per-report hits (src/retrieval.py:1108-1113, 1719, 1754; BM25 :293-295).
"""
import ctypes as C

from . import _lib


def shard_bounds(n, world_size, align=1):
    """Contiguous row ranges, one per rank, starting on multiples of `align` (the synthetic
    generators are seeded per block), balanced to within one block: the step ends with the slowest
    rank, so the largest shard sets the pace.  The ceil(n / align) blocks are dealt out evenly; the
    ranks that get the extra block come first and the (possibly partial) last block goes to the last
    rank, which therefore never holds more rows than the first."""
    blocks = -(-n // align)
    base, extra = divmod(blocks, world_size)
    bounds = [0]
    for r in range(world_size):
        bounds.append(min(n, bounds[-1] + (base + (1 if r < extra else 0)) * align))
    bounds[-1] = n
    return bounds


def merge_topk(D_lists, I_lists, k_out, stream=None):
    """K3 on device tensors laid out [n_lists, nq, k_in] (float32 or float64 scores)."""
    import torch

    lib = _lib.load()
    n_lists, nq, k_in = D_lists.shape
    D = torch.empty((nq, k_out), dtype=D_lists.dtype, device=D_lists.device)
    I = torch.empty((nq, k_out), dtype=torch.int64, device=D_lists.device)
    if stream is None:
        stream = torch.cuda.current_stream(D_lists.device).cuda_stream
    fn = lib.b2r_topk_merge_dev if D_lists.dtype == torch.float32 else lib.b2r_topk_merge_f64_dev
    _lib.check(fn(C.c_void_p(D_lists.data_ptr()), C.c_void_p(I_lists.data_ptr()), int(n_lists), int(nq), int(k_in),
                  int(k_out), C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), C.c_void_p(stream)))
    return D, I


def hybrid_fuse(D, I, S, J, k_out, want_methods=False, stream=None):
    """Dense + BM25 fusion kernel (csrc/hybrid_fuse.cu) on device tensors."""
    import torch

    lib = _lib.load()
    nq = D.shape[0]
    F = torch.empty((nq, k_out), dtype=torch.float32, device=D.device)
    O = torch.empty((nq, k_out), dtype=torch.int64, device=D.device)
    M = torch.empty((nq, k_out), dtype=torch.uint8, device=D.device) if want_methods else None
    if stream is None:
        stream = torch.cuda.current_stream(D.device).cuda_stream
    _lib.check(lib.b2r_hybrid_fuse_dev(C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()), C.c_void_p(S.data_ptr()),
                                       C.c_void_p(J.data_ptr()), int(nq), int(D.shape[1]), int(S.shape[1]), int(k_out),
                                       C.c_void_p(F.data_ptr()), C.c_void_p(O.data_ptr()),
                                       C.c_void_p(M.data_ptr()) if M is not None else None, C.c_void_p(stream)))
    return (F, O, M) if want_methods else (F, O)


def exchange_fuse(D, I, S, J, peer_ptrs, rank, world, nq_max, k_max, step, stream=None, status=None):
    """One-kernel exchange + merge + fusion over peer memory (csrc/hybrid_fuse.cu).  status: int32 CUDA tensor [1]
    (zero = ok) that receives 1 + a missing rank if the bounded wait for the peers expires; None: the kernel traps."""
    import torch

    lib = _lib.load()
    nq, k = int(D.shape[0]), int(D.shape[1])
    F = torch.empty((nq, k), dtype=torch.float32, device=D.device)
    O = torch.empty((nq, k), dtype=torch.int64, device=D.device)
    if stream is None:
        stream = torch.cuda.current_stream(D.device).cuda_stream
    arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
    _lib.check(lib.b2r_hybrid_exchange_fuse_dev(C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                                C.c_void_p(S.data_ptr()), C.c_void_p(J.data_ptr()), nq, k, arr,
                                                int(rank), int(world), int(nq_max), int(k_max), int(step),
                                                C.c_void_p(F.data_ptr()), C.c_void_p(O.data_ptr()),
                                                C.c_void_p(status.data_ptr()) if status is not None else None,
                                                C.c_void_p(stream)))
    return F, O


class PeerExchange:
    """Symmetric buffers of all ranks (torch symmetric memory: CUDA IPC / fabric handles over NVLink)
    for the one-kernel exchange.  Collective constructor: every rank of `group` must call it."""

    def __init__(self, device, group=None, nq_max=64, k_max=128):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nq_max, self.k_max = nq_max, k_max
        nbytes = int(_lib.load().b2r_hybrid_exchange_bytes(nq_max, k_max, self.world))
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.buf, group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.step = 0
        self.status = torch.zeros(1, dtype=torch.int32, device=device)

    def fuse(self, D, I, S, J):
        # the step counter advances only once the launch has been accepted: a call rejected by the argument checks
        # must not leave this rank one step ahead of its peers
        out = exchange_fuse(D, I, S, J, self.ptrs, self.rank, self.world, self.nq_max, self.k_max, self.step + 1,
                            status=self.status)
        self.step += 1
        return out

    def check(self):
        """Synchronises and raises if any exchange kernel since the last check gave up waiting for a peer."""
        v = int(self.status.item())
        if v:
            self.status.zero_()
            raise RuntimeError(f"hybrid exchange timed out waiting for rank {v - 1}: a peer failed or the ranks are out of lockstep")


class ShardedSearcher:
    """SPMD search over a row-sharded corpus.

    local_dense(q, k)            -> (D[nq,k] float32, I[nq,k] int64 GLOBAL ids) on this rank's shard
    local_bm25(tids, toffs, k)   -> (S[nq,k] float64, J[nq,k] int64 GLOBAL ids)  (optional)
    gather(t)                    -> tensor [world, *t.shape] (all ranks' t, rank order)
    merge(D_lists, I_lists, k)   -> merged (D, I)
    fuse(D, I, S, J, k)          -> fused (F, O)
    The defaults bind the CUDA kernels and torch.distributed; tests inject CPU stand-ins to cover the
    partition / gather / merge plumbing with the gloo backend.
    """

    def __init__(self, local_dense=None, local_bm25=None, group=None, world_size=None, gather=None, merge=None,
                 fuse=None):
        self.local_dense = local_dense
        self.local_bm25 = local_bm25
        self.group = group
        if world_size is None:
            import torch.distributed as dist

            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.world_size = world_size
        self._gather = gather or self._dist_gather
        self._merge = merge or merge_topk
        self._fuse = fuse or hybrid_fuse
        self.overlap = None      # "pdl": BM25 beside the scan by programmatic dependent launch (one stream);
                                 # "streams": side stream + high-priority scan stream; None: one after the other
        self.side_stream = None  # "streams" mode: a torch.cuda.Stream for BM25
        self.scan_stream = None  # high-priority stream for the scan while BM25 runs beside it (see search_hybrid)
        self.peer_exchange = None  # set to a PeerExchange to replace all-gather + merges + fusion by one kernel
        self.pipeline = False      # True: consecutive search_hybrid steps are chained by programmatic dependent launch —
                                   # the scan of step s + 1 starts while the fusion / exchange kernel of step s finishes
                                   # (its queries must be resident before step s is enqueued; see b2r_flat_search_dev_ex)
        self.launches = 0      # kernels of this library enqueued by the last call
        self.collectives = 0   # collectives enqueued by the last call

    def _dist_gather(self, t):
        import torch
        import torch.distributed as dist

        flat = t.contiguous().view(-1)
        out = torch.empty(self.world_size * flat.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat, group=self.group)
        self.collectives += 1
        return out.view((self.world_size,) + tuple(t.shape))

    def _gather_pair(self, scores, ids):
        """One collective for a (scores, ids) pair: both are packed into an int64 buffer."""
        import torch

        if scores.dtype == torch.float64:
            packed = torch.stack([scores.view(torch.int64), ids])
        else:
            packed = torch.stack([scores.to(torch.float64).view(torch.int64), ids])
        g = self._gather(packed)  # [world, 2, nq, k]
        s = g[:, 0].contiguous().view(torch.float64)
        if scores.dtype != torch.float64:
            s = s.to(scores.dtype)  # float32 -> float64 -> float32 is exact
        return s.contiguous(), g[:, 1].contiguous()

    def search_dense(self, q, k):
        self.launches = self.collectives = 0
        D, I = self.local_dense(q, k)
        self.launches += self._last_dense_launches()
        if self.world_size == 1:
            return D, I
        Dl, Il = self._gather_pair(D, I)
        self.launches += 1
        return self._merge(Dl, Il, k)

    def search_bm25(self, term_ids, term_offsets, k):
        self.launches = self.collectives = 0
        S, J = self.local_bm25(term_ids, term_offsets, k)
        self.launches += 1
        if self.world_size == 1:
            return S, J
        Sl, Jl = self._gather_pair(S, J)
        self.launches += 1
        return self._merge(Sl, Jl, k)

    def search_hybrid(self, q, term_ids, term_offsets, k):
        """Dense top-k + BM25 top-k + fusion; with several ranks both lists travel in ONE all-gather."""
        import torch

        self.launches = self.collectives = 0
        side = self.side_stream if self.overlap == "streams" else None
        if self.overlap == "pdl":
            # BM25 is latency-bound and small, the scan HBM-bound: BM25 is launched right behind the scan with
            # programmatic stream serialization, so it runs BESIDE the scan — but only once every persistent
            # scan CTA (one per SM, statically assigned tiles) holds its SM.  With two streams both kernels
            # become runnable at the same instant, two BM25 CTAs can take an SM before its scan CTA arrives,
            # and that CTA — hence the whole scan — is delayed by a BM25 lifetime (measured: +70 us per step).
            D, I = self.local_dense(q, k, overlap_previous=True) if self.pipeline else self.local_dense(q, k)
            self.launches += self._last_dense_launches()
            S, J = self.local_bm25(term_ids, term_offsets, k, overlap_previous=True)
        elif side is not None:
            # two-stream variant: BM25 on a side stream, the scan on a high-priority stream so that the block
            # scheduler places its CTAs first (helps, but less reliably than the dependent launch)
            main = torch.cuda.current_stream(q.device)
            ready = main.record_event()
            scan = self.scan_stream
            if scan is not None:
                scan.wait_event(ready)
                with torch.cuda.stream(scan):
                    D, I = self.local_dense(q, k)
            else:
                D, I = self.local_dense(q, k)
            self.launches += self._last_dense_launches()
            side.wait_event(ready)
            with torch.cuda.stream(side):
                S, J = self.local_bm25(term_ids, term_offsets, k)
            if scan is not None:
                main.wait_stream(scan)
                for t in (D, I):
                    t.record_stream(main)
            main.wait_stream(side)
            for t in (S, J):
                t.record_stream(main)
        else:
            D, I = self.local_dense(q, k)
            self.launches += self._last_dense_launches()
            S, J = self.local_bm25(term_ids, term_offsets, k)
        self.launches += 1
        if self.world_size > 1 and self.peer_exchange is not None:
            self.launches += 1
            return self.peer_exchange.fuse(D, I, S, J)  # exchange + merge + fusion in one kernel over NVLink
        if self.world_size > 1:
            packed = torch.stack([D.to(torch.float64).view(torch.int64), I, S.view(torch.int64), J])
            g = self._gather(packed)  # [world, 4, nq, k]
            Dl = g[:, 0].contiguous().view(torch.float64).to(torch.float32)
            Sl = g[:, 2].contiguous().view(torch.float64)
            D, I = self._merge(Dl, g[:, 1].contiguous(), k)
            S, J = self._merge(Sl, g[:, 3].contiguous(), k)
            self.launches += 2
        out = self._fuse(D, I, S, J, k)
        self.launches += 1
        return out

    def _last_dense_launches(self):
        n = getattr(self.local_dense, "last_launches", None)
        return n() if callable(n) else 1


class HybridIndex:
    """Resident dense rows + BM25 postings of one rank's shard, searched together.

    `flat` is a FlatIPIndex over this rank's rows, `bm25` a BM25Okapi over the same chunks (built
    with corpus-global idf / avgdl), `offset` the global id of local row 0."""

    def __init__(self, flat, bm25=None, offset=0, group=None, world_size=None, overlap="pdl", exchange="nccl"):
        self.flat, self.bm25, self.offset = flat, bm25, int(offset)

        def local_dense(q, k, overlap_previous=False):
            return flat.search_dev(q, k, id_offset=self.offset, overlap_previous=overlap_previous)

        local_dense.last_launches = flat.last_launch_count

        def local_bm25(tids, toffs, k, overlap_previous=False):
            return bm25.topk_dev(tids, toffs, k, id_offset=self.offset, overlap_previous=overlap_previous)

        self.searcher = ShardedSearcher(local_dense, local_bm25 if bm25 is not None else None, group=group,
                                        world_size=world_size)
        if overlap is True:
            overlap = "pdl"
        if bm25 is not None and overlap in ("pdl", "streams"):
            self.searcher.overlap = overlap
        if bm25 is not None and overlap == "streams":
            import torch

            dev = torch.device("cuda", flat.device)
            self.searcher.side_stream = torch.cuda.Stream(device=dev)
            self.searcher.scan_stream = torch.cuda.Stream(device=dev, priority=-1)
        if exchange == "peer" and self.searcher.world_size > 1 and bm25 is not None:
            import torch

            self.searcher.peer_exchange = PeerExchange(torch.device("cuda", flat.device), group)
        self._pinned = {}
        self._ctx = None   # b2r_hybrid_ctx of the one-call host step
        self._own_callables = (self.searcher.local_dense, self.searcher.local_bm25)

    # resident API -----------------------------------------------------------------------------------
    def search_dev(self, q, k):
        return self.searcher.search_dense(q, k)

    def search_hybrid_dev(self, q, term_ids, term_offsets, k):
        """Resident hybrid step.  With the dependent-launch overlap (and, on several ranks, the peer exchange) the step is
        ONE C call (b2r_hybrid_step_dev): scan + BM25 beside it, whose last CTA per query fuses / exchanges in place —
        two launches for up to 2 queries; searcher.pipeline chains consecutive steps.  Anything else (NCCL exchange,
        two-stream overlap, injected callables) runs the step from Python (ShardedSearcher.search_hybrid)."""
        s = self.searcher
        native = (self.bm25 is not None and s.overlap == "pdl" and (s.world_size == 1 or s.peer_exchange is not None)
                  and (s.local_dense, s.local_bm25) == self._own_callables and int(k) <= 256)
        if not native:
            return s.search_hybrid(q, term_ids, term_offsets, k)
        import torch

        lib = _lib.load()
        nq, k = int(q.shape[0]), int(k)
        nk = nq * k
        buf = torch.empty(40 * nk, dtype=torch.uint8, device=q.device)   # one allocation: I, S, J, O (8 B) then D, F (4 B)
        I = buf[: 8 * nk].view(torch.int64).view(nq, k)
        S = buf[8 * nk: 16 * nk].view(torch.float64).view(nq, k)
        J = buf[16 * nk: 24 * nk].view(torch.int64).view(nq, k)
        O = buf[24 * nk: 32 * nk].view(torch.int64).view(nq, k)
        D = buf[32 * nk: 36 * nk].view(torch.float32).view(nq, k)
        F = buf[36 * nk: 40 * nk].view(torch.float32).view(nq, k)
        px = s.peer_exchange
        if px is not None:
            arr = (C.c_void_p * len(px.ptrs))(*[C.c_void_p(int(p)) for p in px.ptrs])
            ex = (arr, px.rank, px.world, px.nq_max, px.k_max, px.step + 1, C.c_void_p(px.status.data_ptr()))
        else:
            ex = (None, 0, 1, 0, 0, 0, None)
        self.bm25._ensure_built()
        n_launch = C.c_int(0)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(lib.b2r_hybrid_step_dev(self.flat._h, self.bm25._h, C.c_void_p(q.data_ptr()), nq,
                                           C.c_void_p(term_ids.data_ptr()), C.c_void_p(term_offsets.data_ptr()), k,
                                           self.offset, *ex, C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                           C.c_void_p(S.data_ptr()), C.c_void_p(J.data_ptr()), C.c_void_p(F.data_ptr()),
                                           C.c_void_p(O.data_ptr()), C.c_void_p(stream), 1 if s.pipeline else 0,
                                           C.byref(n_launch)))
        if px is not None:
            px.step += 1
        s.launches, s.collectives = int(n_launch.value), 0
        self.last_parts = (D, I, S, J)
        return F, O

    # host API: the call a user makes (numpy in, numpy out; copies inside) ------------------------------
    def _staging(self, nbytes, nk):
        """Pinned host + device staging buffers (grown on demand), reused across calls."""
        import torch

        dev = torch.device("cuda", self.flat.device)
        st = self._pinned
        if st.get("in_cap", 0) < nbytes:
            cap = max(nbytes, 1 << 16)
            st["in_host"] = torch.empty(cap, dtype=torch.uint8).pin_memory()
            st["in_dev"] = torch.empty(cap, dtype=torch.uint8, device=dev)
            st["in_cap"] = cap
        if st.get("out_cap", 0) < nk:
            cap = max(nk, 4096)
            st["F_host"] = torch.empty(cap, dtype=torch.float32).pin_memory()
            st["O_host"] = torch.empty(cap, dtype=torch.int64).pin_memory()
            st["out_cap"] = cap
        return st

    def search_hybrid(self, x, queries_term_ids, k):
        """x float32[nq,d] (host), queries_term_ids: list of int term-id lists -> (F, O) numpy.
        One packed host->device copy from pinned memory, the resident search, pinned device->host copies of the
        result, one stream synchronisation.  Records the bytes moved each way.  With one rank, or with the
        peer-memory exchange, the whole step is ONE C call (b2r_hybrid_search_host); the NCCL exchange keeps the
        step in Python (torch.distributed owns the collective)."""
        import numpy as np

        s = self.searcher
        if self.bm25 is not None and (s.world_size == 1 or s.peer_exchange is not None):
            return self._search_hybrid_c(x, queries_term_ids, k)
        return self._search_hybrid_py(x, queries_term_ids, k)

    def _search_hybrid_c(self, x, queries_term_ids, k):
        return self.search_hybrid_collect(self.search_hybrid_submit(x, queries_term_ids, k))

    def search_hybrid_submit(self, x, queries_term_ids, k):
        """First half of search_hybrid: stages the inputs, enqueues the H2D copy and the step, returns a ticket at once.
        Up to two submissions may be outstanding (the next question is submitted while the previous one is still on the
        GPU, so the device goes from step to step without waiting for the host); collect them in submission order.
        With several ranks every rank must submit the same sequence.  (One rank, or the peer-memory exchange.)"""
        import numpy as np

        s = self.searcher
        if self.bm25 is None or not (s.world_size == 1 or s.peer_exchange is not None):
            raise RuntimeError("search_hybrid_submit needs a BM25 index and, on several ranks, the peer-memory exchange")
        lib = _lib.load()
        if self._ctx is None:
            h = C.c_void_p()
            _lib.check(lib.b2r_hybrid_ctx_create(int(self.flat.device), C.byref(h)))
            self._ctx = h
        x = np.ascontiguousarray(x, dtype=np.float32)
        nq, d = x.shape
        offs = np.zeros(nq + 1, dtype=np.int32)
        np.cumsum([len(t) for t in queries_term_ids], out=offs[1:])
        tids = (np.concatenate([np.asarray(t, np.int32) for t in queries_term_ids]) if offs[-1]
                else np.zeros(1, np.int32)).astype(np.int32, copy=False)
        px = s.peer_exchange
        if px is not None:
            import torch

            torch.cuda.current_stream(torch.device("cuda", self.flat.device)).synchronize()  # order after resident steps
            arr = (C.c_void_p * len(px.ptrs))(*[C.c_void_p(int(p)) for p in px.ptrs])
            args = (arr, px.rank, px.world, px.nq_max, px.k_max, px.step + 1)
        else:
            args = (None, 0, 1, 0, 0, 0)
        self.bm25._ensure_built()
        ticket = C.c_int(-1)
        _lib.check(lib.b2r_hybrid_submit_host(self._ctx, self.flat._h, self.bm25._h, _lib.ptr(x), nq, d, _lib.ptr(tids),
                                              _lib.ptr(offs), int(k), self.offset, *args, C.byref(ticket)))
        if px is not None:
            px.step += 1   # only after the call succeeded: a rejected call enqueued nothing (the checks come first)
        self.last_h2d_bytes = 4 * nq * d + 4 * (nq + 1) + 4 * int(offs[-1])
        self.last_d2h_bytes = 12 * nq * int(k)
        return (int(ticket.value), nq, int(k))

    def search_hybrid_collect(self, ticket):
        """Second half: waits for the submitted step and returns (F float32[nq,k], O int64[nq,k])."""
        import numpy as np

        t, nq, k = ticket
        F = np.empty((nq, k), dtype=np.float32)
        O = np.empty((nq, k), dtype=np.int64)
        _lib.check(_lib.load().b2r_hybrid_collect_host(self._ctx, t, _lib.ptr(F), _lib.ptr(O)))
        return F, O

    def close(self):
        if getattr(self, "_ctx", None) is not None:
            _lib.load().b2r_hybrid_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _search_hybrid_py(self, x, queries_term_ids, k):
        import numpy as np
        import torch

        x = np.ascontiguousarray(x, dtype=np.float32)
        nq, d = x.shape
        offs = np.zeros(nq + 1, dtype=np.int32)
        np.cumsum([len(t) for t in queries_term_ids], out=offs[1:])
        n_terms = int(offs[-1])
        qb, ob, tb = x.nbytes, offs.nbytes, 4 * n_terms
        st = self._staging(qb + ob + tb, nq * k)
        host = st["in_host"].numpy()
        host[:qb].view(np.float32)[:] = x.reshape(-1)
        host[qb: qb + ob].view(np.int32)[:] = offs
        if n_terms:
            host[qb + ob: qb + ob + tb].view(np.int32)[:] = np.concatenate([np.asarray(t, np.int32) for t in queries_term_ids])
        dev_buf = st["in_dev"]
        dev_buf[: qb + ob + tb].copy_(st["in_host"][: qb + ob + tb], non_blocking=True)
        q_dev = dev_buf[:qb].view(torch.float32).view(nq, d)
        o_dev = dev_buf[qb: qb + ob].view(torch.int32)
        t_dev = dev_buf[qb + ob: qb + ob + max(tb, 4)].view(torch.int32)
        F, O = self.searcher.search_hybrid(q_dev, t_dev, o_dev, k)
        Fh, Oh = st["F_host"][: nq * k], st["O_host"][: nq * k]
        Fh.copy_(F.reshape(-1), non_blocking=True)
        Oh.copy_(O.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(F.device).synchronize()
        self.last_h2d_bytes, self.last_d2h_bytes = qb + ob + tb, nq * k * 12
        return Fh.numpy().reshape(nq, k).copy(), Oh.numpy().reshape(nq, k).copy()
