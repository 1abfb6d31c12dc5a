"""Sample-time ("ensemble") parallelism across GPUs.

The reference gives every averaging sample s_k its own MPI ensemble member
(Ensemble(COMM_WORLD, space_parallel), /root/reference/averaged_sw_explicit.py:37,233-235)
and joins them with one `ensemble.allreduce` per averaged evaluation
(timestepping_methods.py:10-11).  Here the members are the columns of a device batch; the
Mbar columns are split into contiguous blocks, one block per GPU/process, the weighted
sum over the local block is a kernel, and the sum over GPUs is one all-reduce of n doubles
(NCCL over NVLink; gloo in the CPU tests).
"""
import numpy as np

from .samples import sample_times_weights


def block_partition(Mbar, world):
    """Contiguous blocks k in [offsets[g], offsets[g+1]) -- as even as possible."""
    base, extra = divmod(int(Mbar), int(world))
    sizes = [base + (1 if g < extra else 0) for g in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(int)


def _panel_widths(k, cap=32):
    """Panel widths the library gives a batch of k columns: full panels of `cap`, the rest padded to a power of two >= 2."""
    out = [cap] * (k // cap)
    r = k % cap
    if r:
        w = 2
        while w < r:
            w *= 2
        out.append(w)
    return out


def balanced_partition(terms, world, latency_columns=6.4, forward_columns=6.8, cap=32):
    """Unequal, cost-balanced shares of the averaging samples for `world` GPUs.

    terms[k]: Chebyshev terms of sample k (cheby_exp.series_terms).  The samples are sorted by series length (= by |t_k|) and cut
    into `world` contiguous groups; a group costs  max(terms) * forward_columns  for the shared-input forward run plus, per
    backward panel (sorted, at most `cap` columns),  max(terms of the panel) * (latency_columns + panel width):  a panel's time
    per term is a latency floor plus a part proportional to its width (measured at ref 5: 70 + 11 W us per term, forward run
    75 us per term; profiles/r02_summary.md section 2).  The cut minimising the most expensive group is found by dynamic
    programming.  Returns a list of `world` index arrays (every sample exactly once).  With few samples per GPU the long
    series decide the step: the rank that holds them gets fewer samples, the ranks with short series more."""
    terms = np.asarray(terms, dtype=np.float64)
    B = len(terms)
    world = int(world)
    if world < 1 or B < world:
        raise ValueError("need at least one sample per rank")
    order = np.argsort(terms, kind="stable")
    ts = terms[order]

    def cost(i, j):                      # samples i..j-1 of the sorted order
        c = ts[j - 1] * forward_columns
        lo = i
        for w in _panel_widths(j - i, cap):
            hi = min(lo + min(w, cap), j)
            c += ts[hi - 1] * (latency_columns + w)
            lo = hi
        return c
    INF = float("inf")
    best = np.full((world + 1, B + 1), INF)
    arg = np.zeros((world + 1, B + 1), dtype=np.int64)
    best[0, 0] = 0.0
    for g in range(1, world + 1):
        for j in range(g, B - (world - g) + 1):
            for i in range(g - 1, j):
                if best[g - 1, i] == INF:
                    continue
                c = max(best[g - 1, i], cost(i, j))
                if c < best[g, j]:
                    best[g, j] = c
                    arg[g, j] = i
    cuts = [B]
    j = B
    for g in range(world, 0, -1):
        j = int(arg[g, j])
        cuts.append(j)
    cuts = cuts[::-1]
    return [np.sort(order[cuts[g]:cuts[g + 1]]) for g in range(world)]


def partition_cost(terms, groups, latency_columns=6.4, forward_columns=6.8, cap=32):
    """Model cost (in units of one column-term) of the most expensive group of a partition; see balanced_partition."""
    terms = np.asarray(terms, dtype=np.float64)
    worst = 0.0
    for g in groups:
        ts = np.sort(terms[np.asarray(g)])
        c = ts[-1] * forward_columns
        lo = 0
        for w in _panel_widths(len(ts), cap):
            hi = min(lo + min(w, cap), len(ts))
            c += ts[hi - 1] * (latency_columns + w)
            lo = hi
        worst = max(worst, c)
    return worst


class SampleEnsemble:
    def __init__(self, Mbar, rank=None, world=None, partition="block", terms=None):
        """partition: "block" (contiguous sample blocks per rank), "cyclic" (sample k on rank k mod world) or "balanced"
        (unequal shares by cost, balanced_partition; needs terms = cheby_exp.series_terms(t), the series length of every
        sample).  With the economised series the cost of a sample grows with |t_k|, and the sample times are ordered, so
        "cyclic" gives every rank the same mix of short and long series; "balanced" gives the rank with the longest series
        fewer samples, which is faster when a rank holds only a few."""
        try:
            import torch.distributed as dist
            self._dist = dist if dist.is_available() and dist.is_initialized() else None
        except Exception:                      # torch not importable: single process
            self._dist = None
        if rank is None:
            rank = self._dist.get_rank() if self._dist else 0
        if world is None:
            world = self._dist.get_world_size() if self._dist else 1
        self.Mbar, self.rank, self.world = int(Mbar), int(rank), int(world)
        if partition not in ("block", "cyclic", "balanced"):
            raise ValueError("partition must be 'block', 'cyclic' or 'balanced'")
        self.partition = partition
        self.offsets = block_partition(self.Mbar, self.world)
        self.groups = None
        if partition == "balanced":
            if terms is None or len(terms) != self.Mbar:
                raise ValueError("partition='balanced' needs terms (one series length per sample, cheby_exp.series_terms)")
            self.groups = balanced_partition(terms, self.world)
        self._lib_comm = None

    @property
    def local_slice(self):
        """Index of this rank's samples: a slice ("block", "cyclic") or an index array ("balanced")."""
        if self.partition == "balanced":
            return self.groups[self.rank]
        if self.partition == "cyclic":
            return slice(self.rank, self.Mbar, self.world)
        return slice(int(self.offsets[self.rank]), int(self.offsets[self.rank + 1]))

    @property
    def n_local(self):
        if self.partition == "balanced":
            return int(len(self.groups[self.rank]))
        return len(range(*self.local_slice.indices(self.Mbar)))

    def local_times_weights(self, dt, rho):
        t, w = sample_times_weights(self.Mbar, dt, rho)
        return np.ascontiguousarray(t[self.local_slice]), np.ascontiguousarray(w[self.local_slice])

    # ---- the one collective of the path
    def allreduce(self, v):
        """Sum v over the ensemble in place (numpy array or torch tensor)."""
        if self.world == 1:
            return v
        if self._dist is None:
            raise RuntimeError(f"SampleEnsemble(world={self.world}) has no communicator: torch.distributed is not initialised and "
                               "no library communicator was set up (init_library_comm); refusing to return this rank's partial sum")
        import torch
        if isinstance(v, np.ndarray):
            tv = torch.from_numpy(v)
            if self._dist.get_backend() == "nccl":
                tv = tv.cuda()
                self._dist.all_reduce(tv)
                v[...] = tv.cpu().numpy()
            else:
                self._dist.all_reduce(tv)
            return v
        self._dist.all_reduce(v)
        return v

    def init_library_comm(self, op):
        """NCCL communicator inside libkexp.so (kexp_comm_init): the all-reduce then runs on the
        library's own stream right behind the weighted column sum, without going through torch."""
        import ctypes as C
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            rc = op.lib.kexp_comm_unique_id(uid.ctypes.data_as(C.c_void_p))
            if rc:
                raise RuntimeError(f"kexp_comm_unique_id failed ({rc})")
        if self._dist is not None and self.world > 1:
            box = [uid.tobytes()]
            self._dist.broadcast_object_list(box, src=0)
            uid = np.frombuffer(box[0], dtype=np.uint8).copy()
        op._check(op.lib.kexp_comm_init(op.ctx, uid.ctypes.data_as(C.c_void_p), self.rank, self.world), "kexp_comm_init")
        self._lib_comm = op

    def averaged_apply(self, cheby, X_local, t_local, w_local, shared=False, fused=False):
        """sum over ALL ranks and samples of w_k exp(t_k L) X_k  (timestepping_methods.py:9-11)."""
        use_lib = self._lib_comm is cheby.op and self.world > 1
        v = cheby.apply_average(X_local, t_local, w_local, shared=shared, allreduce=use_lib, fused=fused)
        if not use_lib:
            self.allreduce(v)
        return v
