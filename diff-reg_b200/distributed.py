"""Multi-GPU paths of the hot path.

(i)  Independent units -- registration pairs / diffusion samples -- are sharded over the ranks with NO collective:
     `shard_units` is the partition (unit index mod world size); bench.py runs one sample per GPU this way.
(ii) One very large Sinkhorn (BASELINE.json configs[4]: N = M = 16384, 100 iterations) is ROW-SHARDED: rank r owns a
     contiguous block of rows, `v` is replicated, and the ranks all-reduce the per-column log-sum-exp partials once per
     iteration.  The reference has no counterpart (its matrix always lives on one GPU).

The collective is log-sum-exp, which NCCL does not offer: `lse_allreduce` does MAX on the maxima, rescales the sums and
does SUM (two small all-reduces), or -- once a per-column reference is known -- a single SUM of sums rescaled to that
reference (the previous iteration's column log-sum-exp, which the iteration converges to).
"""
import torch
import torch.distributed as dist

from . import ops


def shard_units(n_units, world_size, rank):
    """Indices of the independent units (pairs / samples) this rank processes."""
    return list(range(rank, n_units, world_size))


def shard_rows(n_rows, world_size, rank):
    """Contiguous-equal row block [start, stop) of rank `rank` (the first n_rows % world_size ranks get one extra row)."""
    base, extra = divmod(n_rows, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def lse_combine(partials):
    """Reference combine of a list of [.., 2] (max, sum) log2-domain partials (single process; used by tests and by the
    one-process emulation of the sharded path)."""
    m = torch.stack([p[..., 0] for p in partials]).max(dim=0)[0]
    s = sum(p[..., 1] * torch.exp2(p[..., 0] - m) for p in partials)
    return torch.stack((m, s), dim=-1)


def lse_allreduce(partial, group=None, ref=None):
    """All-reduce [.., 2] (max, sum) log2-domain partials over the process group.
    ref=None: exact two-step (MAX, then SUM of rescaled sums).  ref=[..] tensor: one SUM with every rank's sum rescaled
    to the shared reference (valid while |max - ref| stays far below the fp32 exponent range)."""
    m_loc, s_loc = partial[..., 0], partial[..., 1]
    if ref is None:
        m = m_loc.clone()
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    else:
        m = ref
    s = s_loc * torch.exp2(torch.clamp(m_loc - m, max=120.0))
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return torch.stack((m, s), dim=-1)


class P2PUnavailable(RuntimeError):
    """CUDA IPC / peer access is not available between the ranks (raised on every rank of the group together)."""


class P2PComm:
    """Peer-mapped exchange buffers of the ranks of one node (drg_p2p_*): every rank allocates an inbox, the CUDA IPC
    handles are all-gathered over the process group, and every rank maps every inbox.  Kernels then store into the
    peers' inboxes directly over NVLink / NVSwitch -- the all-reduce of the row-sharded Sinkhorn happens inside its own
    kernel instead of two NCCL calls per iteration.  One process per GPU, all on the same node, world size <= 8."""

    def __init__(self, slot_elems, nflags, group=None, device=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load_library()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.slot_elems, self.nflags = int(slot_elems), int(nflags)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        hb = int(self.lib.drg_p2p_handle_bytes())
        mine = (ctypes.c_ubyte * hb)()
        comm = ctypes.c_void_p()
        backend = dist.get_backend(group)
        cdev = dev if backend == "nccl" else torch.device("cpu")

        def all_ok(rc):        # every rank learns whether every rank succeeded, so that they fail (or go on) together
            flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=cdev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag.item()) == 1

        rc = self.lib.drg_p2p_create(self.rank, self.world, self.slot_elems, self.nflags, ctypes.byref(comm), mine)
        self.handle = comm if rc == 0 else None
        if not all_ok(rc):
            msg = self.lib.drg_last_error().decode() if rc else "another rank failed"
            self._abort()
            raise P2PUnavailable(f"drg_p2p_create: {msg}")
        # all-gather the handles through tensors of the group's backend
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=cdev)
        allh = torch.empty(self.world * hb, dtype=torch.uint8, device=cdev)
        dist.all_gather_into_tensor(allh, t, group=group)
        raw = bytes(allh.cpu().tolist())
        buf = (ctypes.c_ubyte * len(raw)).from_buffer_copy(raw)
        rc = self.lib.drg_p2p_connect(self.handle, buf)
        if not all_ok(rc):                 # also the barrier: nobody sends before everybody has mapped everybody
            msg = self.lib.drg_last_error().decode() if rc else "another rank failed"
            self._abort()
            raise P2PUnavailable(f"drg_p2p_connect: {msg}")

    def _abort(self):
        if self.handle is not None:
            self.lib.drg_p2p_destroy(self.handle)
            self.handle = None

    def fits(self, B, M):
        return B * (M + 1) <= self.slot_elems and B * ((M + 1 + 31) // 32) <= self.nflags

    def status(self):
        return int(self.lib.drg_p2p_status(self.handle))

    def close(self):
        if self.handle is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self.lib.drg_p2p_destroy(self.handle)
            self.handle = None


class RowShardedSinkhorn:
    """log_optimal_transport over a row-sharded score matrix.  Call on every rank with its local rows.

    exchange="p2p" (default when the backend is NCCL): the per-iteration all-reduce of the column partials runs inside
    the kernel over peer-mapped memory (P2PComm).  exchange="nccl": two (later one) NCCL all-reduces per iteration
    between the local pass and the update -- the portable path, also used under gloo in the CPU tests' logic."""

    def __init__(self, group=None, single_allreduce_after=2, exchange=None):
        self.group = group
        self.single_after = single_allreduce_after
        self.exchange = exchange
        self.comm = None
        self._shape = None
        self._p2p_failed = False

    def _p2p(self, B, M):
        mode = self.exchange
        if mode is None:
            mode = "p2p" if dist.get_backend(self.group) == "nccl" else "nccl"
        if mode != "p2p":
            return None
        if self._p2p_failed:
            return None
        if self.comm is None or self._shape != (B, M):   # the flag <-> column-chunk mapping is per shape
            if self.comm is not None:
                self.comm.close()
                self.comm = None
            try:
                self.comm = P2PComm(B * (M + 1), B * ((M + 1 + 31) // 32), self.group)
                self._shape = (B, M)
            except P2PUnavailable as e:       # raised on EVERY rank together (the ranks agree before connecting)
                if self.exchange == "p2p":
                    raise
                import warnings
                warnings.warn(f"RowShardedSinkhorn: peer-to-peer exchange unavailable ({e}); using NCCL all-reduces")
                self._p2p_failed = True
                return None
        return self.comm

    def _check_exchange(self, comm, device):
        """A wait inside the exchange kernel that timed out (a peer never arrived) leaves a status word instead of
        hanging the GPU; the result of that call is then invalid on at least one rank and the cumulative flag counters
        are out of step.  Every rank learns about it (MAX over the group), the comm is dropped and the call raises."""
        bad = torch.tensor([comm.status()], dtype=torch.int32, device=device if dist.get_backend(self.group) == "nccl" else "cpu")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if int(bad.item()) != 0:
            comm._abort()
            self.comm = None
            self._shape = None
            raise RuntimeError("RowShardedSinkhorn: the peer-to-peer exchange timed out on at least one rank; the result is "
                               "invalid (the exchange buffers were released; the next call reconnects)")

    @torch.no_grad()
    def __call__(self, scores_local, alpha, iters, src_mask_local, tgt_mask, out_mode="conf", apply_mask=False):
        st = ops.ShardedSinkhornState(scores_local, alpha, src_mask_local, tgt_mask, apply_mask)
        counts = st.local_counts()
        src_total = counts[:, 0].clone()
        dist.all_reduce(src_total, op=dist.ReduceOp.SUM, group=self.group)
        counts[:, 0] = src_total
        st.begin(counts)
        comm = self._p2p(st.B, st.M)
        if comm is not None:
            st.iterate_exchange(comm.handle, int(iters))
            out = st.final(out_mode)
            self._check_exchange(comm, scores_local.device)
            return out
        ref = None
        for it in range(int(iters)):
            partial = st.local()
            reduced = lse_allreduce(partial, self.group, ref if it >= self.single_after else None)
            st.update(reduced)
            # the column log-sum-exp this iteration found is the next iteration's rescaling reference
            ref = reduced[..., 0] + torch.log2(reduced[..., 1].clamp_min(1e-38))
        return st.final(out_mode)


class EmulatedRowShards:
    """The same algorithm with all P shards living in ONE process on one GPU (no process group): validates the sharded
    kernels against the unsharded result where only one GPU is available."""

    def __init__(self, n_shards):
        self.P = n_shards

    @torch.no_grad()
    def __call__(self, scores, alpha, iters, src_mask, tgt_mask, out_mode="conf", apply_mask=False):
        B, N, M = scores.shape
        bounds = [shard_rows(N, self.P, r) for r in range(self.P)]
        states = [ops.ShardedSinkhornState(scores[:, a:b].contiguous(), alpha, src_mask[:, a:b].contiguous(), tgt_mask, apply_mask)
                  for a, b in bounds]
        counts = states[0].local_counts()
        counts[:, 0] = sum(s.local_counts()[:, 0] for s in states)
        for s in states:
            s.begin(counts)
        for _ in range(int(iters)):
            reduced = lse_combine([s.local().clone() for s in states])
            for s in states:
                s.update(reduced)
        return torch.cat([s.final(out_mode)[:, : (b - a)] for s, (a, b) in zip(states, bounds)], dim=1)
