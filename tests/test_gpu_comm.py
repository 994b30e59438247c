"""GPU: libnsr_b200's one-kernel gradient all-reduce (csrc/nsr_comm.cu, include/nsr.h nsr_comm_*).

On ONE device: N comms (N "ranks", one handle and one stream each) wired with nsr_comm_connect_ptrs launch the
collective concurrently and must all end with the mean, summed in fixed rank order -- bit-identical everywhere and
bit-equal to ((g0 + g1) + g2 ...) * (1/N) evaluated in fp32.  The multi-process CUDA-IPC wiring of the same kernel
(P2PGradReducer) is exercised by the 2-GPU runs of bench.py (`train.allreduce_impl`) and tests/test_gpu_multi.py."""
import ctypes as C

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _view(lib, c):
    from nerf_sr_b200.parallel import _DevicePointer
    return torch.as_tensor(_DevicePointer(int(lib.nsr_comm_buffer(c)), int(lib.nsr_comm_buffer_floats(c))), device=DEV)


@pytest.mark.parametrize("world,n_floats", [(2, 2 * 595844), (4, 2 * 595844), (3, 10007), (8, 2 * 595844)])
def test_p2p_allreduce_mean_is_exact_and_identical_on_every_rank(world, n_floats):
    from nerf_sr_b200 import Renderer
    cfg = O.RenderConfig()
    rs = [Renderer(cfg, DEV, precision="bf16x3") for _ in range(world)]
    lib = rs[0].lib
    comms = []
    for rank, r in enumerate(rs):
        c = C.c_void_p()
        r._check(lib.nsr_comm_create(r._h, rank, world, n_floats, C.byref(c)))
        comms.append(c)
    ptrs = (C.c_void_p * world)(*[lib.nsr_comm_buffer(c) for c in comms])
    for r, c in zip(rs, comms):
        r._check(lib.nsr_comm_connect_ptrs(c, ptrs, world))
    bufs = [_view(lib, c) for c in comms]
    assert bufs[0].numel() >= n_floats and bufs[0].numel() % 4096 == 0
    streams = [torch.cuda.Stream(DEV) for _ in range(world)]
    g = torch.Generator(device=DEV).manual_seed(world)
    for it in range(3):                                   # epochs: the flags are never reset
        data = [torch.randn(n_floats, device=DEV, generator=g) * (10.0 ** (k - 2)) for k in range(world)]
        for b, d in zip(bufs, data):
            b[:n_floats].copy_(d)
        torch.cuda.synchronize()
        for r, c, st in zip(rs, comms, streams):
            with torch.cuda.stream(st):
                r._check(lib.nsr_comm_allreduce_mean(c, st.cuda_stream))
        torch.cuda.synchronize()
        want = data[0].clone()
        for d in data[1:]:
            want = want + d                               # fp32, rank order
        want = want * torch.tensor(1.0 / world, dtype=torch.float32, device=DEV)
        for k, b in enumerate(bufs):
            assert torch.equal(b[:n_floats], want), (world, it, k, float((b[:n_floats] - want).abs().max()))
            assert float(b[n_floats:].abs().max() if b.numel() > n_floats else 0.0) == 0.0      # the zero padding stays zero
    for r, c in zip(rs, comms):
        lib.nsr_comm_destroy(c)
        r.close()


def test_comm_argument_errors():
    from nerf_sr_b200 import NsrError, Renderer
    r = Renderer(O.RenderConfig(), DEV, precision="bf16x3")
    c = C.c_void_p()
    assert r.lib.nsr_comm_create(r._h, 2, 2, 100, C.byref(c)) == 1        # rank out of range
    assert r.lib.nsr_comm_create(r._h, 0, 17, 100, C.byref(c)) == 1       # world too large
    r._check(r.lib.nsr_comm_create(r._h, 0, 2, 100, C.byref(c)))
    with pytest.raises(NsrError):                                         # peers not connected yet
        r._check(r.lib.nsr_comm_allreduce_mean(c, None))
    r.lib.nsr_comm_destroy(c)
    one = C.c_void_p()
    r._check(r.lib.nsr_comm_create(r._h, 0, 1, 100, C.byref(one)))        # world 1: a no-op collective
    r._check(r.lib.nsr_comm_allreduce_mean(one, None))
    r.lib.nsr_comm_destroy(one)
    r.close()
