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


def configure_overlap(reserve_sms: int = 8, nccl_channels: int = 8) -> None:
    """Call BEFORE `init_process_group` and before the first forward.  The output all-gather runs concurrently with
    the next batch's kernels; NCCL's copy kernels use one SM per channel, and the persistent tensor-core kernel uses
    one CTA per SM -- if they collide, the persistent kernel's last CTAs run as a second wave.  So NCCL is limited to
    `nccl_channels` channels (enough for 100 MB per rank per ~13 ms step even at 8 ranks) and the persistent grids leave
    `reserve_sms` SMs free.  Measured on 2 B200s (migan-512, 32 img/GPU): 14.27 ms/step without the reservation,
    13.32-13.39 ms with 4-8 SMs reserved (12.9 ms on one GPU).  Respects values already present in the environment."""
    os.environ.setdefault("NCCL_MAX_NCHANNELS", str(nccl_channels))
    os.environ.setdefault("NCCL_MIN_NCHANNELS", str(nccl_channels))
    os.environ.setdefault("MIGAN_TC_RESERVE_SMS", str(reserve_sms))


def shard_bounds(global_n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous rank-major shard [lo, hi) of a global batch; the first (global_n % world_size)
    ranks get one extra image."""
    if global_n < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard request (%d, %d, %d)" % (global_n, world_size, rank))
    base, rem = divmod(global_n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GatherHandle:
    """Result of `forward_async`: `wait()` makes the current stream wait for the collective and
    returns the gathered tensor [world*n_local, 3, R, R] (rank-major)."""

    def __init__(self, out: torch.Tensor, work, local: torch.Tensor):
        self._out, self._work, self.local = out, work, local

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._out


class ShardedGenerator:
    """Data-parallel wrapper.  `model` is any callable x[n,4,R,R] -> y[n,3,R,R] on this rank's device
    (the B200 `Generator`; CPU stand-ins are used by the gloo tests)."""

    def __init__(self, model: Callable[[torch.Tensor], torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun, one process per GPU)")
        self.model = model
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

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

    def forward_async(self, x_local: torch.Tensor) -> GatherHandle:
        y = self.model(x_local)
        if self.world_size == 1:
            return GatherHandle(y, None, y)
        out = torch.empty((self.world_size * y.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        work = dist.all_gather_into_tensor(out, y, group=self.group, async_op=True)
        return GatherHandle(out, work, y)

    def forward(self, x_local: torch.Tensor) -> torch.Tensor:
        """Every rank passes its shard (equal sizes) and receives all outputs, rank-major."""
        return self.forward_async(x_local).wait()

    def forward_global(self, x_global: torch.Tensor) -> torch.Tensor:
        """Every rank passes the same global batch (size divisible by the world size)."""
        if x_global.shape[0] % self.world_size:
            raise ValueError("global batch %d is not divisible by world size %d" % (x_global.shape[0], self.world_size))
        lo, hi = shard_bounds(x_global.shape[0], self.world_size, self.rank)
        return self.forward(x_global[lo:hi].contiguous())

    __call__ = forward
