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


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pin this process (and therefore its pinned host buffers, first-touch) to the CPUs NVML reports as local to GPU
    ``local_rank``: with one process per GPU on a two-socket box, host->device copies of a rank whose pages sit on the
    other socket cross the socket interconnect and share it with every other such rank.  Returns what was done (for
    the bench record); never raises — a container whose cpuset does not intersect the GPU's CPUs is left alone."""
    if os.environ.get("CTDR_NO_NUMA_BIND"):
        return "off (CTDR_NO_NUMA_BIND)"
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(ncpu, 1024) + 63) // 64)
        local = {i for i in range(64 * len(words)) if (int(words[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        if not cpus:
            return f"skipped: GPU-local CPUs ({len(local)}) are outside this process's cpuset ({len(allowed)} CPUs)"
        if set(cpus) == set(allowed):
            return f"all {len(allowed)} allowed CPUs are local to the GPU"
        os.sched_setaffinity(0, cpus)
        return f"bound to {len(cpus)} of {len(allowed)} CPUs ({cpus[0]}-{cpus[-1]})"
    except Exception as e:                                  # noqa: BLE001 — placement is an optimisation, never a failure
        return f"skipped: {type(e).__name__}: {e}"


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


class PeerAllReduce:
    """SUM over the ranks of one node by ONE kernel that reads the peers' buffers over NVLink (csrc/ctdr_peer.cuh).

    Every rank owns a *peer block* (``ctdr_peer_block_create``), the 64-byte CUDA IPC handles travel once through
    ``all_gather_object`` and every rank maps the others' blocks.  ``run`` then replaces the library all-reduce of
    the gradient exchange: copy -> signal every peer -> wait for every peer -> add all blocks in rank order
    (bit-identical sums on every rank), one launch, no host involvement, capturable in a CUDA graph.
    ``blocks`` may also be handed in directly (all ranks in one process: the single-GPU emulation of the tests).
    """

    # one-shot: every rank reads every block — right for the 0.6 MB of a 50k-vertex mesh.  Beyond ~1 MB the library
    # all-reduce wins: measured on 4 GPUs with the 6 MB of a 500k-vertex mesh, NCCL 0.083 ms vs 0.16 ms for a two-phase
    # variant of this kernel (each slice summed by an owner rank, 48 CTAs), which was therefore not kept.
    MAX_PAYLOAD = 1 << 20
    MAX_RANKS = 8

    def __init__(self, payload_bytes: int, device, rank: int, world: int, blocks, owned: int, opened=()):
        self.payload_bytes, self.device, self.rank, self.world = int(payload_bytes), torch.device(device), rank, world
        self.blocks, self.owned, self.opened = list(blocks), owned, list(opened)
        import ctypes
        self._table = (ctypes.c_void_p * world)(*self.blocks)

    @staticmethod
    def create_block(payload_bytes: int, want_handle: bool = True):
        """(device pointer, 64-byte IPC handle or None) of a fresh peer block on the current device."""
        import ctypes
        from . import _lib
        L = _lib.lib()
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)() if want_handle else None
        _lib.check(L.ctdr_peer_block_create(payload_bytes, ctypes.byref(ptr), handle), "ctdr_peer_block_create")
        return ptr.value, (bytes(handle) if want_handle else None)

    @classmethod
    def connect(cls, payload_bytes: int, device):
        """Collective over the default process group (one process per GPU, one node); None when any rank cannot
        take part (no peer access, IPC refused, payload too large): callers then keep the library all-reduce."""
        import ctypes
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(), dist.get_rank()
        if world < 2 or world > cls.MAX_RANKS or payload_bytes > cls.MAX_PAYLOAD or dist.get_backend() != "nccl":
            return None
        L = _lib.lib()
        dev = torch.device(device)
        mine, handle, err = None, None, None
        try:
            with torch.cuda.device(dev):
                mine, handle = cls.create_block(payload_bytes)
        except Exception as e:                                 # noqa: BLE001 — reported below, the group falls back together
            err = repr(e)
        infos = [None] * world
        dist.all_gather_object(infos, (handle, os.uname().nodename, err))
        ok = all(h is not None and node == infos[0][1] for h, node, _ in infos)
        opened, blocks = [], [None] * world
        if ok:
            try:
                with torch.cuda.device(dev):
                    for r, (h, _, _) in enumerate(infos):
                        if r == rank:
                            blocks[r] = mine
                            continue
                        p = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                        _lib.check(L.ctdr_peer_block_open(buf, ctypes.byref(p)), "ctdr_peer_block_open")
                        opened.append(p.value); blocks[r] = p.value
            except Exception:                                  # noqa: BLE001
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # also orders "every block is mapped" before first use
        if int(flag.item()) == 0:
            for p in opened:
                L.ctdr_peer_block_close(p)
            if mine is not None:
                dist.barrier()                                 # nobody still maps it
                L.ctdr_peer_block_destroy(mine)
            return None
        return cls(payload_bytes, dev, rank, world, blocks, mine, opened)

    def run(self, storage: torch.Tensor, f32_bytes: int, f64_bytes: int) -> None:
        from . import _lib
        _lib.check(_lib.lib().ctdr_peer_allreduce(self._table, self.world, self.rank, self.payload_bytes, storage.data_ptr(),
                                                  f32_bytes, f64_bytes, _lib.stream_ptr(self.device)), "ctdr_peer_allreduce")

    def status(self) -> int:
        """0 = fine; 1 = some launch gave up waiting for a peer (its sums are invalid).  Synchronises."""
        import ctypes
        from . import _lib
        st = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ctdr_peer_block_status(self.owned, ctypes.byref(st)), "ctdr_peer_block_status")
        return st.value

    def close(self, collective: bool = True) -> None:
        from . import _lib
        L = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for p in self.opened:
                L.ctdr_peer_block_close(p)
            self.opened = []
            if collective and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.barrier()                                 # every rank has unmapped before anyone frees
            if self.owned is not None:
                L.ctdr_peer_block_destroy(self.owned)
                self.owned = None


class GradientExchange:
    """Flat, pre-allocated all-reduce buffers the backward writes into directly (no pack kernel).

    ``flat`` (fp32) = [grad_verts (3V) | grad_mus (n)], ``scalars`` (fp64) = [loss_sum, n_valid].  Both are views
    of ONE byte buffer, so a single fill zeroes them and ONE launch sums them over the ranks: the library's own
    kernel over NVLink peer memory (``PeerAllReduce``; NCCL process group, one node, payload up to 4 MB), else the
    two all-reduces (different dtypes) issued as one NCCL group behind the backward kernels on the current stream.
    ``peer``: None = use the peer kernel when it can be set up (``CTDR_PEER_ALLREDUCE=0`` disables), False = never.
    Construction is collective when a process group is active.
    """

    def __init__(self, num_verts: int, num_mus: int, device, peer: bool | None = None):
        nflat = 3 * num_verts + num_mus
        off = (4 * nflat + 15) // 16 * 16                     # fp64 pair behind the fp32 part, 16-byte aligned
        self.storage = torch.zeros(off + 16, dtype=torch.uint8, device=device)
        self.flat = self.storage[:4 * nflat].view(torch.float32)
        self.grad_verts = self.flat[:3 * num_verts].view(num_verts, 3)
        self.grad_mus = self.flat[3 * num_verts:]
        self.scalars = self.storage[off:off + 16].view(torch.float64)
        self._f32_bytes = off
        self.peer = None
        if peer is None:
            peer = os.environ.get("CTDR_PEER_ALLREDUCE", "1") != "0"
        if peer and self._active() and self.storage.is_cuda:
            self.peer = PeerAllReduce.connect(off + 16, self.storage.device)

    @property
    def collective(self) -> str:
        if not self._active():
            return "none (one rank)"
        return "k_peer_allreduce: one launch over NVLink peer memory" if self.peer is not None else "NCCL all-reduce (fp32 buffer + fp64 pair)"

    def zero_(self):
        self.storage.zero_()

    @staticmethod
    def _active() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def all_reduce_tensors(self, flat, scalars):
        """SUM over ranks of an fp32 buffer and an fp64 pair through the process group: two plain all-reduces.
        (torch's ``_coalescing_manager(device=...)`` around the pair returned WITHOUT reducing for these mixed
        dtypes on torch 2.11 / NCCL 2.28 — caught by tests/test_gpu_multi_rank.py on two GPUs — so the library
        path stays the simple one; the one-launch exchange is ``PeerAllReduce``.)"""
        if not self._active():
            return
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM)

    def all_reduce(self):
        """Sum over ranks, enqueued on the current stream right behind the backward kernels."""
        if self.peer is not None:
            self.peer.run(self.storage, self._f32_bytes, 16)
            return
        self.all_reduce_tensors(self.flat, self.scalars)

    def kernels_per_exchange(self, world: int) -> int:
        """Kernels one zero_() + all_reduce() enqueues besides the library's own (for launch accounting): the fill,
        plus the NCCL group when the peer kernel (counted by the library) is not in use."""
        return 1 + (2 if world > 1 and self.peer is None else 0)

    def loss_sum_and_count(self):
        return tuple(float(x) for x in self.scalars.tolist())

    def mean_gradients(self):
        """d(mean over valid pixels)/d(verts, mus) and the mean loss, from the all-reduced sums."""
        n = self.scalars[1].clamp(min=1.0)
        inv = (1.0 / n).to(torch.float32)
        return self.grad_verts * inv, self.grad_mus * inv, self.scalars[0] / n
