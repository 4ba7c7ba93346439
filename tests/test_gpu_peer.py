"""GPU: the one-launch all-reduce over peer memory (csrc/ctdr_peer.cuh, ``distributed.PeerAllReduce``).

One GPU is enough: N "ranks" live in this process, each with its own peer block, payload and stream, and their N
kernels run concurrently on the device — the same signalling (release store into the peer's flag array, acquire
spin on one's own) and the same rank-ordered sums as N processes over NVLink.  The N-process path over CUDA IPC is
covered by tests/test_gpu_multi_rank.py (needs two GPUs).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _payload(gen, nf, dev):
    """[nf float32 | pad to 16 B | 2 float64] like GradientExchange.storage; returns (storage, f32_bytes)."""
    off = (4 * nf + 15) // 16 * 16
    st = torch.zeros(off + 16, dtype=torch.uint8, device=dev)
    st[:4 * nf].view(torch.float32).copy_(torch.randn(nf, generator=gen, device=dev))
    st[off:].view(torch.float64).copy_(torch.randn(2, generator=gen, device=dev, dtype=torch.float64) * 1e6)
    return st, off


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_peer_allreduce_emulated_ranks(cuda_lib, world):
    from shapefromprojections_b200.distributed import PeerAllReduce
    dev = torch.device("cuda", 0)
    nf = 3 * 50000 + 2                                    # config 3: 600 KB
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + world)
    payload_bytes = (4 * nf + 15) // 16 * 16 + 16
    blocks = [PeerAllReduce.create_block(payload_bytes, want_handle=False)[0] for _ in range(world)]
    ranks = [PeerAllReduce(payload_bytes, dev, r, world, blocks, owned=blocks[r]) for r in range(world)]
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    try:
        for epoch in range(5):                            # both payload slots, several times over
            data, off = zip(*[_payload(gen, nf, dev) for _ in range(world)])
            off = off[0]
            want_f = torch.zeros(off // 4, device=dev)
            want_d = torch.zeros(2, device=dev, dtype=torch.float64)
            for r in range(world):                        # rank order, like the kernel
                want_f = want_f + data[r][:off].view(torch.float32)
                want_d = want_d + data[r][off:].view(torch.float64)
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    ranks[r].run(data[r], off, 16)
            torch.cuda.synchronize()
            for r in range(world):
                assert ranks[r].status() == 0, f"rank {r} gave up waiting (epoch {epoch})"
                assert torch.equal(data[r][:off].view(torch.float32), want_f), (world, r, epoch)
                assert torch.equal(data[r][off:].view(torch.float64), want_d), (world, r, epoch)
    finally:
        for p in ranks:
            p.close(collective=False)


def test_peer_allreduce_in_a_cuda_graph(cuda_lib):
    """The launch carries no per-step argument (the epoch lives in the block): replaying a captured launch works."""
    from shapefromprojections_b200.distributed import PeerAllReduce
    dev = torch.device("cuda", 0)
    nf = 4096
    payload_bytes = 4 * nf + 16
    block = PeerAllReduce.create_block(payload_bytes, want_handle=False)[0]
    pr = PeerAllReduce(payload_bytes, dev, 0, 1, [block], owned=block)
    st = torch.zeros(payload_bytes, dtype=torch.uint8, device=dev)
    try:
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            pr.run(st, 4 * nf, 16)                        # warm-up outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            pr.run(st, 4 * nf, 16)
        for k in range(4):
            st[:4 * nf].view(torch.float32).fill_(float(k + 1))
            g.replay()
            torch.cuda.synchronize()
            assert torch.all(st[:4 * nf].view(torch.float32) == float(k + 1))
        assert pr.status() == 0
    finally:
        pr.close(collective=False)


def test_peer_allreduce_rejects_bad_sizes(cuda_lib):
    import ctypes
    from shapefromprojections_b200 import _lib
    L = _lib.lib()
    p = ctypes.c_void_p()
    assert L.ctdr_peer_block_create(24, ctypes.byref(p), None) == -2      # CTDR_E_SHAPE: not a multiple of 16
    assert L.ctdr_peer_block_bytes(1024) == 4096 + 2 * 1024
    table = (ctypes.c_void_p * 1)(None)
    buf = torch.zeros(64, dtype=torch.uint8, device="cuda")
    assert L.ctdr_peer_allreduce(table, 1, 0, 64, buf.data_ptr(), 48, 16, None) == -1      # NULL block
    assert L.ctdr_peer_allreduce(table, 9, 0, 64, buf.data_ptr(), 48, 16, None) == -2      # too many ranks
