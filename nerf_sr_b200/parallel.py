"""Ray sharding across the GPUs of one node.

Every ray is independent (the s x s box average couples only the s*s consecutive rows of one LR
pixel), so inference needs NO data-path collective: rank g renders rows [lo_g, hi_g) with shard
boundaries at multiples of s*s, weights are replicated, results stay rank-local or are gathered once
(16 B per LR pixel).  The reference instead wraps each MLP in nn.DataParallel and scatters/gathers
the [P,90] point tensor around every call (models/networks.py:54-69) -- not reproduced.
Training's only collective is the gradient all-reduce of DDP (models/networks.py:72-86): one flat
4.77 MB fp32 bucket, mean over ranks -- `allreduce_mean_` below (NCCL on GPUs, gloo in the CPU tests).
Training batches are sharded the way the reference's DDP loader does it (data/__init__.py:94-113:
`DistributedSampler(seed=opt.seed)` + per-rank batch = batch_size / n_gpus): `epoch_indices` / `rank_batches` below.
One process per GPU; torch.distributed is used for plumbing only."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rays: int, world_size: int, group_rows: int = 1) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) row ranges, one per rank, with boundaries at multiples of `group_rows`
    (= s*s sub-pixel rays of one LR pixel).  Earlier ranks get the remainder groups."""
    if n_rays % group_rows:
        raise ValueError(f"n_rays={n_rays} is not a multiple of the sub-pixel group {group_rows}")
    groups = n_rays // group_rows
    base, rem = divmod(groups, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + (base + (1 if r < rem else 0)) * group_rows
        out.append((lo, hi))
        lo = hi
    return out


def render_sharded(render_fn: Callable[[torch.Tensor], Sequence[torch.Tensor]], rays: torch.Tensor, s: int = 1,
                   group: Optional[dist.ProcessGroup] = None, gather: bool = True):
    """Render this rank's shard of `rays` with `render_fn(rays_shard) -> (rgb[n/s^2,3], depth[n/s^2])`
    and (optionally) all-gather the LR results so every rank holds the full frame, in row order.
    Works with any backend (NCCL for CUDA tensors, gloo for the CPU tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(rays.shape[0], world, s * s)
    lo, hi = bounds[rank]
    outs = [o.reshape(o.shape[0], -1) for o in render_fn(rays[lo:hi])]
    if not gather or world == 1:
        return outs, bounds
    full = []
    for o in outs:
        sizes = [(b[1] - b[0]) // (s * s) for b in bounds]
        mx = max(sizes)
        pad = torch.zeros(mx, o.shape[1], dtype=o.dtype, device=o.device)
        pad[: o.shape[0]] = o
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full.append(torch.cat([p[:n] for p, n in zip(parts, sizes)], 0))
    return full, bounds


def _allreduce_mean_flat_(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> None:
    """Mean over ranks of ONE flat tensor, in place, with as few launches as the backend allows: NCCL averages inside
    the collective (ncclAvg); gloo has no AVG, so SUM then one division."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)


def allreduce_mean_(tensors: Sequence[torch.Tensor], group: Optional[dist.ProcessGroup] = None) -> None:
    """DDP-equivalent gradient averaging of arbitrary tensors (the autograd bridge hands over two separately allocated
    flat buffers): flatten into ONE bucket (2 x 595 844 fp32 = 4.77 MB for the coarse+fine nets), one all-reduce,
    scatter back in place.  `Trainer` does not come through here: its gradients already live in one bucket
    (`make_grad_reducer`)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if len(tensors) == 1 and tensors[0].is_contiguous():
        _allreduce_mean_flat_(tensors[0].view(-1), group)
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    _allreduce_mean_flat_(flat, group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


class _DevicePointer:
    """Zero-copy view of library-owned device memory as a torch tensor (``torch.as_tensor`` reads this protocol)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class TorchGradReducer:
    """The flat gradient bucket in ordinary device memory, averaged with ONE torch.distributed collective
    (NCCL: ncclAvg, no extra elementwise launch; gloo in the CPU tests)."""

    def __init__(self, n_floats: int, device, group=None):
        self.group = group
        self.buffer = torch.zeros(n_floats, device=device, dtype=torch.float32)
        self.impl = f"torch.distributed all_reduce on one flat bucket ({dist.get_backend(group)}" + \
                    (", ncclAvg)" if dist.get_backend(group) == "nccl" else ", sum + divide)")

    def allreduce_mean_(self) -> None:
        _allreduce_mean_flat_(self.buffer, self.group)

    def close(self) -> None:
        pass


class P2PGradReducer:
    """The flat gradient bucket in libnsr_b200's symmetric memory, averaged by the library's own one-kernel all-reduce over
    NVLink peer loads / stores (nsr_comm_allreduce_mean, csrc/nsr_comm.cu).  torch.distributed only carries the 64-byte
    CUDA IPC handles at construction; the collective itself is a single kernel launch on the current stream and its
    result is bit-identical on every rank (fixed summation order)."""

    def __init__(self, renderer, n_floats: int, group=None):
        import ctypes as C
        self.r, self.group = renderer, group
        lib = renderer.lib
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        c = C.c_void_p()
        renderer._check(lib.nsr_comm_create(renderer._h, rank, world, n_floats, C.byref(c)))
        self._c = c
        mine = (C.c_char * 64)()
        renderer._check(lib.nsr_comm_export(c, mine))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine), group=group)           # plumbing only: 64 bytes per rank
        blob = b"".join(handles)
        renderer._check(lib.nsr_comm_connect_ipc(c, blob, world))
        self.buffer = torch.as_tensor(_DevicePointer(int(lib.nsr_comm_buffer(c)), int(lib.nsr_comm_buffer_floats(c))),
                                      device=renderer.device)
        dist.barrier(group=group)                                            # everyone mapped everyone before first use
        self.impl = "nsr_comm_allreduce_mean: one kernel, two-shot over CUDA-IPC peer memory (NVLink ld/st), fixed rank order"

    def allreduce_mean_(self) -> None:
        self.r._check(self.r.lib.nsr_comm_allreduce_mean(self._c, self.r._stream()))

    def close(self) -> None:
        if getattr(self, "_c", None):
            self.buffer = None
            self.r.lib.nsr_comm_destroy(self._c)
            self._c = None


def make_grad_reducer(renderer, n_floats: int, group=None, mode: str = "auto"):
    """The gradient bucket + its all-reduce for `Trainer`.  mode: "p2p" | "nccl" | "auto" (p2p when the process group is NCCL
    on CUDA devices, else the torch.distributed collective).  The environment variable NSR_ALLREDUCE overrides "auto"."""
    import os
    if mode == "auto":
        mode = os.environ.get("NSR_ALLREDUCE", "auto")
    backend = dist.get_backend(group)
    if mode == "auto":
        mode = "p2p" if (backend == "nccl" and renderer.device.type == "cuda") else "nccl"
    if mode == "p2p":
        return P2PGradReducer(renderer, n_floats, group)
    if mode == "nccl":
        return TorchGradReducer(n_floats, renderer.device, group)
    raise ValueError(f"unknown all-reduce mode {mode!r}")


def epoch_indices(n_samples: int, world_size: int, rank: int, epoch: int = 0, seed: int = 0, shuffle: bool = True,
                  drop_last: bool = False) -> torch.Tensor:
    """The sample indices rank `rank` visits in `epoch`: torch.utils.data.DistributedSampler's rule, which the
    reference's DDP loader uses over the LR-pixel buffers (data/__init__.py:94-101).  One permutation per epoch from
    `seed + epoch` (identical on every rank), padded by wrapping around to a multiple of the world size (or truncated with
    `drop_last`), then rank r takes every world_size-th entry starting at r -- disjoint shards that cover the set."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n_samples, generator=g)
    else:
        idx = torch.arange(n_samples)
    if drop_last and n_samples % world_size:
        total = (n_samples // world_size) * world_size
        idx = idx[:total]
    else:
        total = ((n_samples + world_size - 1) // world_size) * world_size
        pad = total - n_samples
        if pad:
            reps = (pad + n_samples - 1) // n_samples
            idx = torch.cat([idx, idx.repeat(reps)[:pad]])
    return idx[rank:total:world_size]


def rank_batches(n_samples: int, batch_size: int, world_size: int, rank: int, epoch: int = 0, seed: int = 0,
                 shuffle: bool = True, keep_last: bool = False):
    """Per-rank batches of one epoch: the global `batch_size` must divide by the world size (the reference asserts it,
    data/__init__.py:94-96); every rank steps through its `epoch_indices` in chunks of batch_size / world_size, dropping the
    ragged last chunk unless `keep_last` (`--keep_last`).  Yields LongTensors for `scenes.take_batch`."""
    if batch_size % world_size:
        raise ValueError(f"batch_size {batch_size} is not divisible by the number of ranks {world_size}")
    per_rank = batch_size // world_size
    idx = epoch_indices(n_samples, world_size, rank, epoch, seed, shuffle)
    for lo in range(0, idx.numel(), per_rank):
        chunk = idx[lo:lo + per_rank]
        if chunk.numel() < per_rank and not keep_last:
            return
        yield chunk
