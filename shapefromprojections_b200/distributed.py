"""Angle-sharded data parallelism: the only multi-GPU strategy of this path.

Every projection angle is independent (``fp_vecs.py:101``, ``fp_opengl.py:89-90``), so rank ``r`` owns a
contiguous block of ``idx_angles``; the mesh is replicated; sinogram / target shards never leave
their GPU.  The one exchange per optimiser iteration is an all-reduce (SUM) of the vertex and
attenuation gradients plus the two loss scalars.  The reference has no distributed code at all.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; single process when unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def strided_angles(nangles: int, rank: int, world: int):
    """Angles ``rank, rank+world, ...``: every rank sees the whole orbit, so the per-angle cost
    variation (edge-on views carry fewer fragments than face-on ones) averages out; contiguous
    blocks were up to ~20 % imbalanced on the torus workload."""
    import torch
    if rank >= nangles:
        return torch.zeros(0, dtype=torch.int64)
    return torch.arange(rank, nangles, world, dtype=torch.int64)


def shard_angles(nangles: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of rank ``rank``; block sizes differ by at most one angle."""
    base, rem = divmod(nangles, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradientExchange:
    """Flat, pre-allocated all-reduce buffers the backward writes into directly (no pack kernel).

    ``flat`` (fp32) = [grad_verts (3V) | grad_mus (n)], ``scalars`` (fp64) = [loss_sum, n_valid].  Both are views
    of ONE byte buffer, so a single fill zeroes them, and the two all-reduces (different dtypes) are issued as
    one NCCL group (one launch) behind the backward kernels on the current stream.
    """

    def __init__(self, num_verts: int, num_mus: int, device):
        nflat = 3 * num_verts + num_mus
        off = (4 * nflat + 15) // 16 * 16                     # fp64 pair behind the fp32 part, 16-byte aligned
        self.storage = torch.zeros(off + 16, dtype=torch.uint8, device=device)
        self.flat = self.storage[:4 * nflat].view(torch.float32)
        self.grad_verts = self.flat[:3 * num_verts].view(num_verts, 3)
        self.grad_mus = self.flat[3 * num_verts:]
        self.scalars = self.storage[off:off + 16].view(torch.float64)

    def zero_(self):
        self.storage.zero_()

    @staticmethod
    def _active() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def all_reduce_tensors(self, flat, scalars):
        """SUM over ranks of an fp32 buffer and an fp64 pair as one grouped launch where the backend can."""
        if not self._active():
            return
        manager = getattr(dist, "_coalescing_manager", None)
        if manager is not None and dist.get_backend() == "nccl":
            try:
                with manager(device=flat.device):
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                    dist.all_reduce(scalars, op=dist.ReduceOp.SUM)
                return
            except (TypeError, RuntimeError):                 # API drift between torch versions: two plain calls
                pass
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM)

    def all_reduce(self):
        """Sum over ranks, enqueued on the current stream right behind the backward kernels."""
        self.all_reduce_tensors(self.flat, self.scalars)

    @staticmethod
    def kernels_per_exchange(world: int) -> int:
        """Kernels one zero_() + all_reduce() enqueues besides the library's own (for launch accounting)."""
        return 1 + (1 if world > 1 else 0)

    def loss_sum_and_count(self):
        return tuple(float(x) for x in self.scalars.tolist())

    def mean_gradients(self):
        """d(mean over valid pixels)/d(verts, mus) and the mean loss, from the all-reduced sums."""
        n = self.scalars[1].clamp(min=1.0)
        inv = (1.0 / n).to(torch.float32)
        return self.grad_verts * inv, self.grad_mus * inv, self.scalars[0] / n
