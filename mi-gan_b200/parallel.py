"""Batch sharding over the GPUs of one box: one process per GPU, weights replicated, each rank
runs `Generator.forward` on its shard and the outputs are all-gathered over NCCL (NVLink).

The reference has no inference parallelism at all; images are independent
(no batch statistics, migan_inference.py:165-167), so this is pure data parallelism with ONE
collective per batch and nothing else on the data path.

The all-gather runs asynchronously on NCCL's stream: `forward_async` returns a handle and the
next batch's kernels start immediately, so the transfer of batch t overlaps the compute of
batch t+1 (in a serving loop the collective is then off the critical path).
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def configure_overlap(reserve_sms: int = None, nccl_channels: int = 8, gather: str = "ce") -> None:
    """Call BEFORE `init_process_group` and before the first forward.

    gather = "ce" (default): the output all-gather moves its bytes with the copy engines (peer-to-peer writes over NVLink, see
    `ShardedGenerator`) and signals completion with stream memory operations (or, MIGAN_CE_SIGNAL=nccl, an 8-byte all-reduce):
    no collective kernel runs, the persistent tensor-core kernel keeps all 148 SMs (no reservation).
    gather = "nccl": `all_gather_into_tensor` runs on NCCL's SM-resident copy kernels, one SM per channel, concurrently with
    the next batch's kernels.  The persistent kernel uses one CTA per SM -- if they collide, its last CTAs run as a second
    wave -- so NCCL is limited to `nccl_channels` channels and the persistent grids leave `reserve_sms` SMs free.  Measured
    on 2 B200s (migan-512, 32 img/GPU, round 1): 14.27 ms/step without the reservation, 13.3-13.4 ms with 4-8 SMs reserved
    (12.9 ms on one GPU).  Respects values already present in the environment."""
    if gather == "nccl":
        os.environ.setdefault("NCCL_MAX_NCHANNELS", str(nccl_channels))
        os.environ.setdefault("NCCL_MIN_NCHANNELS", str(nccl_channels))
        os.environ.setdefault("MIGAN_TC_RESERVE_SMS", str(8 if reserve_sms is None else reserve_sms))
    else:
        os.environ.setdefault("MIGAN_TC_RESERVE_SMS", str(0 if reserve_sms is None else reserve_sms))


def shard_bounds(global_n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous rank-major shard [lo, hi) of a global batch; the first (global_n % world_size)
    ranks get one extra image."""
    if global_n < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard request (%d, %d, %d)" % (global_n, world_size, rank))
    base, rem = divmod(global_n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GatherHandle:
    """Result of `forward_async`: `wait()` makes the current stream wait for the gather and returns the gathered tensor
    [world*n_local, 3, R, R] (rank-major).  `local` is this rank's own output."""

    def __init__(self, out: torch.Tensor, work, local: torch.Tensor, event=None):
        self._out, self._work, self.local, self._event = out, work, local, event

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._event is not None:
            torch.cuda.current_stream(self._out.device).wait_event(self._event)
            self._event = None
        return self._out


class ShardedGenerator:
    """Data-parallel wrapper.  `model` is any callable x[n,4,R,R] -> y[n,3,R,R] on this rank's device
    (the B200 `Generator`; CPU stand-ins are used by the gloo tests).

    gather = "nccl": one `all_gather_into_tensor` per batch (NCCL's copy kernels, SM-resident).
    gather = "ce" (CUDA tensors, all ranks on one box): the same all-gather with the bytes moved by the copy engines.  Every
      rank publishes a ring of RING gathered-output buffers to its peers once (CUDA IPC, through torch's tensor sharing); per
      batch it (1) writes its rows into the current slot of every rank's ring with peer-to-peer `cudaMemcpyAsync` over NVLink
      (one stream per peer), then (2) joins an 8-byte NCCL all-reduce -- the only collective kernel -- whose completion means
      "every rank's rows have landed here".  No SM is taken from the compute kernels, so nothing has to be reserved for the
      collective, and the transfer of batch t overlaps the compute of batch t+1.  LIFETIME: `forward_async` returns a view of
      the ring; it stays valid until RING - 1 further `forward_async` calls (consume or clone it before).  `forward` /
      `forward_global` return a private copy.
      Ready signal (`signal`, env MIGAN_CE_SIGNAL): "memops" (default when the driver offers stream memory operations) --
      after its rows the sender copies a 4-byte step counter into the receiver's arrival table, and the receiver's stream
      waits on those words with `cuStreamWaitValue32`; no kernel runs for the collective at all, so the persistent
      tensor-core kernel (one CTA per SM, static tile assignment) never finds an SM taken by a spinning collective kernel.
      "nccl": the 8-byte all-reduce described above (measured on 8 B200s: its kernel, waiting for the slowest rank while
      holding an SM, made the compute kernels' last CTAs run as a second wave: 0.65 scaling efficiency).
    gather = "auto": "ce" when possible, else "nccl"."""

    RING = 4   # gathered buffers per rank (copy-engine path): the tensor of step t is valid until forward_async(t + RING - 1)

    def __init__(self, model: Callable[[torch.Tensor], torch.Tensor], group: Optional[dist.ProcessGroup] = None,
                 gather: str = "auto"):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun, one process per GPU)")
        if gather not in ("auto", "ce", "nccl"):
            raise ValueError("gather must be 'auto', 'ce' or 'nccl'")
        self.model = model
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.gather = gather
        self._ce = None          # lazily built state of the copy-engine path
        self._step = 0

    def check_replicas(self, state_dict) -> None:
        """All ranks must hold identical weights (analogue of the reference's
        torch_utils/misc.py:178-187 check_ddp_consistency): compare a checksum with rank 0."""
        total = torch.zeros(2, dtype=torch.float64)
        for v in state_dict.values():
            v = v.detach().double().cpu()
            total[0] += v.sum()
            total[1] += v.abs().sum()
        dev = next(iter(state_dict.values())).device
        total = total.to(dev if dev.type == "cuda" else "cpu")
        ref = total.clone()
        dist.broadcast(ref, src=0, group=self.group)
        if not torch.equal(ref, total):
            raise RuntimeError("rank %d holds different weights than rank 0" % self.rank)

    # -- copy-engine all-gather ---------------------------------------------------------------------------------------
    def _setup_ce(self, y: torch.Tensor):
        """Publish RING gathered-output buffers to every peer and map theirs (one-time, collective)."""
        from torch.multiprocessing.reductions import reduce_tensor
        n = y.shape[0]
        ring = [torch.empty((self.world_size * n,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device) for _ in range(self.RING)]
        mine = [reduce_tensor(t) for t in ring]                       # (rebuild_fn, args): CUDA IPC handle + offset
        everyone = [None] * self.world_size
        dist.all_gather_object(everyone, mine, group=self.group)
        peers = {}
        for p in range(self.world_size):
            if p != self.rank:
                peers[p] = [fn(*args) for fn, args in everyone[p]]    # tensors aliasing rank p's ring (peer-mapped)
                if tuple(peers[p][0].shape) != tuple(ring[0].shape):
                    raise RuntimeError("rank %d published buffers of a different shape" % p)
        self._ce = {"ring": ring, "peers": peers, "side": torch.cuda.Stream(device=y.device),
                    "push": {p: torch.cuda.Stream(device=y.device) for p in peers},
                    "flag": torch.zeros(1, device=y.device), "ready": None, "shape": tuple(y.shape)}
        signal = os.environ.get("MIGAN_CE_SIGNAL", "memops")
        if signal == "memops":
            from . import _abi
            if not _abi.load().b200_stream_memops_available():
                signal = "nccl"
        votes = [None] * self.world_size
        dist.all_gather_object(votes, signal, group=self.group)
        signal = "memops" if all(v == "memops" for v in votes) else "nccl"      # every rank must use the same protocol
        self._ce["signal"] = signal
        if signal == "memops":
            # arrival table: arrive[q] = (last step + 1) whose rows from rank q have landed in this rank's ring
            arrive = torch.zeros(self.world_size, dtype=torch.int32, device=y.device)
            stage = torch.zeros(self.world_size, dtype=torch.int32, device=y.device)    # per-destination source word of the 4-byte copy
            tables = [None] * self.world_size
            dist.all_gather_object(tables, reduce_tensor(arrive), group=self.group)
            self._ce["arrive"] = arrive
            self._ce["stage"] = stage
            self._ce["peer_arrive"] = {p: tables[p][0](*tables[p][1]) for p in peers}
            torch.cuda.synchronize(y.device)
            dist.barrier(group=self.group)              # every table is zeroed and mapped before the first signal is sent

    def _forward_ce(self, y: torch.Tensor) -> GatherHandle:
        """Push form: this rank writes its rows into slot t % RING of EVERY rank's gathered ring (posted peer-to-peer writes,
        one stream per peer so several copy engines / NVLink ports work at once), then joins the 8-byte all-reduce.  Its
        completion on a rank means every rank's rows have landed there.  Slot t % RING was last handed out at step t - RING and,
        by the lifetime rule of `forward_async`, is no longer in use once its owner has CALLED forward_async(t - 1) -- which is
        what the completed ready signal of step t - 1 certifies for every rank."""
        ce, t, n = self._ce, self._step, y.shape[0]
        if ce["signal"] == "memops":
            return self._forward_ce_memops(y)
        slot = t % self.RING
        cur = torch.cuda.current_stream(y.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        side = ce["side"]
        rows = slice(self.rank * n, (self.rank + 1) * n)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            if ce["ready"] is not None:
                ce["ready"].wait()                    # step t-1's signal: every rank is past its use of this slot
            base = torch.cuda.Event()
            base.record(side)
            ce["ring"][slot][rows].copy_(y, non_blocking=True)
            for k in range(1, self.world_size):       # staggered: at step k every rank writes to a different destination
                p = (self.rank + k) % self.world_size
                bufs = ce["peers"][p]
                st = ce["push"][p]
                st.wait_event(base)
                with torch.cuda.stream(st):
                    bufs[slot][rows].copy_(y, non_blocking=True)      # peer-to-peer write, copy engine
                    done_p = torch.cuda.Event()
                    done_p.record(st)
                side.wait_event(done_p)
            ready = dist.all_reduce(ce["flag"], group=self.group, async_op=True)   # "my rows are in every rank's slot"
        y.record_stream(side)
        for st in ce["push"].values():
            y.record_stream(st)
        ce["ready"] = ready
        return GatherHandle(ce["ring"][slot], ready, y)

    def _forward_ce_memops(self, y: torch.Tensor) -> GatherHandle:
        """Push form with no collective kernel.  Per destination q, on its own stream: wait until q's rows of step t-1 have
        landed HERE (arrive[q] >= t: q has then passed its forward_async(t-1), i.e. is done with the slot about to be
        overwritten -- the lifetime rule of `forward_async`), copy this rank's rows into slot t % RING of q's ring, then copy
        the step counter t+1 into q's arrival table.  The gathered tensor is complete here when arrive[q] >= t+1 for every q;
        those waits run on a side stream, the consumer waits for one event."""
        from . import _abi
        lib = _abi.load()
        ce, t, n = self._ce, self._step, y.shape[0]
        slot = t % self.RING
        cur = torch.cuda.current_stream(y.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        side = ce["side"]
        rows = slice(self.rank * n, (self.rank + 1) * n)
        arrive, stage = ce["arrive"], ce["stage"]
        word = arrive.element_size()
        tv, tn = t & 0xFFFFFFFF, (t + 1) & 0xFFFFFFFF
        with torch.cuda.device(y.device):
            for k in range(1, self.world_size):       # staggered: at step k every rank writes to a different destination
                p = (self.rank + k) % self.world_size
                st = ce["push"][p]
                st.wait_event(ev)
                _abi.check(lib.b200_stream_wait_value32(st.cuda_stream, arrive.data_ptr() + p * word, tv))
                with torch.cuda.stream(st):
                    ce["peers"][p][slot][rows].copy_(y, non_blocking=True)                       # peer-to-peer write, copy engine
                _abi.check(lib.b200_stream_write_value32(st.cuda_stream, stage.data_ptr() + p * word, tn))
                with torch.cuda.stream(st):
                    ce["peer_arrive"][p][self.rank:self.rank + 1].copy_(stage[p:p + 1], non_blocking=True)   # 4 bytes: "my rows of step t are in"
                y.record_stream(st)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                ce["ring"][slot][rows].copy_(y, non_blocking=True)
                for p in ce["peers"]:
                    _abi.check(lib.b200_stream_wait_value32(side.cuda_stream, arrive.data_ptr() + p * word, tn))
                done = torch.cuda.Event()
                done.record(side)
            y.record_stream(side)
        return GatherHandle(ce["ring"][slot], None, y, done)

    def forward_async(self, x_local: torch.Tensor) -> GatherHandle:
        y = self.model(x_local)
        if self.world_size == 1:
            return GatherHandle(y, None, y)
        mode = self.gather
        if mode in ("auto", "ce") and y.is_cuda:
            if self._ce is None or self._ce["shape"] != tuple(y.shape):
                try:
                    self._setup_ce(y)
                except Exception:
                    if mode == "ce":
                        raise
                    self.gather = mode = "nccl"       # e.g. no peer access between the devices
            if mode != "nccl":
                h = self._forward_ce(y)
                self._step += 1
                return h
        out = torch.empty((self.world_size * y.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        work = dist.all_gather_into_tensor(out, y, group=self.group, async_op=True)
        return GatherHandle(out, work, y)

    def forward(self, x_local: torch.Tensor) -> torch.Tensor:
        """Every rank passes its shard (equal sizes) and receives all outputs, rank-major (a tensor the caller owns)."""
        out = self.forward_async(x_local).wait()
        return out.clone() if self._ce is not None else out

    def forward_global(self, x_global: torch.Tensor) -> torch.Tensor:
        """Every rank passes the same global batch (size divisible by the world size)."""
        if x_global.shape[0] % self.world_size:
            raise ValueError("global batch %d is not divisible by world size %d" % (x_global.shape[0], self.world_size))
        lo, hi = shard_bounds(x_global.shape[0], self.world_size, self.rank)
        return self.forward(x_global[lo:hi].contiguous())

    def forward_host_async(self, x_host_local: torch.Tensor, out_host_local: torch.Tensor) -> None:
        """End-to-end serving step of one rank: pinned host shard -> device, forward, all-gather, and this rank's rows of the
        GATHERED tensor back to pinned host memory (each process returns its part of the batch; the gather sits on the
        path).  Three streams and two device input slots: the H2D copy of batch t+1 and the D2H copy of batch t-1 run under
        the kernels of batch t.  Enqueue-only: call `host_wait()` before reading `out_host_local`."""
        dev = torch.device("cuda", torch.cuda.current_device())
        hs = self.__dict__.get("_host")
        if hs is None or hs["shape"] != tuple(x_host_local.shape):
            hs = {"shape": tuple(x_host_local.shape), "x": [torch.empty(x_host_local.shape, dtype=x_host_local.dtype, device=dev) for _ in range(2)],
                  "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev), "done": [None, None], "k": 0}
            self._host = hs
        k = hs["k"]
        slot = k % 2
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(hs["h2d"]):
            if hs["done"][slot] is not None:
                hs["h2d"].wait_event(hs["done"][slot])       # the forward that read this slot two steps ago has finished
            hs["x"][slot].copy_(x_host_local, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(hs["h2d"])
        cur.wait_event(ev_in)
        if k >= 2 and hs.get("d2h_done", {}).get(k - 2) is not None:
            cur.wait_event(hs["d2h_done"].pop(k - 2))    # ring lifetime: the copy-out of step k-2 precedes this step's ready signal
        h = self.forward_async(hs["x"][slot])
        g = h.wait()
        ev_c = torch.cuda.Event()
        ev_c.record(cur)
        hs["done"][slot] = ev_c
        n = x_host_local.shape[0]
        with torch.cuda.stream(hs["d2h"]):
            hs["d2h"].wait_event(ev_c)
            out_host_local.copy_(g[self.rank * n:(self.rank + 1) * n], non_blocking=True)
            ev_o = torch.cuda.Event()
            ev_o.record(hs["d2h"])
        hs.setdefault("d2h_done", {})[k] = ev_o
        g.record_stream(hs["d2h"])
        hs["k"] = k + 1

    def close(self) -> None:
        """Drop the peer mappings of the copy-engine path (collective: call on every rank before the process group is destroyed,
        so that no rank frees a published buffer while a peer still maps it)."""
        if self._ce is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._ce["peers"].clear()
            self._ce.get("peer_arrive", {}).clear()
            dist.barrier(group=self.group)
            self._ce = None

    def host_wait(self) -> None:
        hs = self.__dict__.get("_host")
        if hs is not None:
            hs["d2h"].synchronize()

    __call__ = forward
