"""CPU (gloo, world_size 2) test of the slab-decomposition host logic and index math used by the CUDA engine
(csrc/engine_cufft.cu: k_slab_pack + all-to-all + strided z transform): the same pack / exchange / unpack formulas,
with NumPy FFTs standing in for the local transforms, must reproduce the global rfftn / irfftn."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _a2a(send):
    """send[r] goes to rank r; returns recv with recv[s] from rank s (what ncclSend/ncclRecv in a group does)."""
    t = torch.from_numpy(np.ascontiguousarray(send).view(np.float64))
    out = torch.empty_like(t)
    dist.all_to_all_single(out, t)
    return out.numpy().view(np.complex128).reshape(send.shape)


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ptf_b200.parallel import slab_extents
    nx, ny, nz = n
    nkr = nx // 2 + 1
    nzl, zoff = slab_extents(nz, world, rank)
    nyl, yoff = slab_extents(ny, world, rank)
    rng = np.random.default_rng(3)
    c = rng.standard_normal((nz, ny, nx))
    ref = np.fft.rfftn(c)
    # ---- forward: local 2-D r2c -> pack -> all-to-all -> z transform (engine_cufft.cu fwd()) ----
    T1 = np.fft.rfft2(c[zoff:zoff + nzl])                               # [nzl][ny][nkr]
    T2 = np.empty((world, nzl, nyl, nkr), dtype=np.complex128)
    for r in range(world):
        T2[r] = T1[:, r * nyl:(r + 1) * nyl, :]                         # T2[r][zl][jl][kx] = T1[zl][r*nyl+jl][kx]
    spec = _a2a(T2).reshape(nz, nyl, nkr)                               # [s][nzl][nyl][nkr] == [nz][nyl][nkr]
    spec = np.fft.fft(spec, axis=0)
    e_f = np.abs(spec - ref[:, yoff:yoff + nyl, :]).max() / np.abs(ref).max()
    # ---- inverse: z transform -> all-to-all (no send-side pack) -> unpack -> local 2-D c2r (inv()) ----
    s = np.fft.ifft(spec, axis=0) * nz                                  # unnormalised, as cuFFT
    T2 = _a2a(s.reshape(world, nzl, nyl, nkr))
    T1 = np.empty((nzl, ny, nkr), dtype=np.complex128)
    for sidx in range(world):
        T1[:, sidx * nyl:(sidx + 1) * nyl, :] = T2[sidx]               # T1[zl][s*nyl+jl][kx] = T2[s][zl][jl][kx]
    back = np.fft.irfft2(T1, s=(ny, nx)) / nz
    e_i = np.abs(back - c[zoff:zoff + nzl]).max()
    q.put((rank, float(e_f), float(e_i)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [(16, 8, 12), (10, 12, 8)])
def test_slab_transpose_scheme_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_f, e_i in res:
        assert e_f < 1e-13 and e_i < 1e-13, (rank, e_f, e_i)


def test_slab_extents():
    import ptf_b200
    from ptf_b200.parallel import slab_extents
    assert slab_extents(1024, 8, 3) == (128, 384)
    with pytest.raises(ValueError):
        slab_extents(10, 4, 0)


# ---------------------------------------------------------------------------------------------------------------------
# 2-D slab scheme (csrc/engine_slab2d.cu): physical rows sharded, spectral kr-columns sharded in kc = ceil(nkr/P) padded
# columns, spectral layout [kc][ny].  Same pack / exchange / unpack formulas as the kernels k2_pack_rows, k2_unpack_rows
# and the two transposes, with NumPy FFTs standing in for the local cuFFT batches.
# ---------------------------------------------------------------------------------------------------------------------
def _worker2d(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny = n
    nkr = nx // 2 + 1
    nyp = ny // world
    kc = -(-nkr // world)
    koff = rank * kc
    kvalid = max(0, min(kc, nkr - koff))
    rng = np.random.default_rng(5)
    c = rng.standard_normal((ny, nx))
    ref = np.fft.rfft2(c)                                                  # [ny][nkr]
    rows = np.fft.rfft(c[rank * nyp:(rank + 1) * nyp], axis=1)             # local 1-D r2c: [nyp][nkr]
    # k2_pack_rows: blocks[d][jl][cc] = rows[jl][d*kc + cc] (zero beyond nkr)
    blocks = np.zeros((world, nyp, kc), dtype=np.complex128)
    for d in range(world):
        w = max(0, min(kc, nkr - d * kc))
        blocks[d, :, :w] = rows[:, d * kc:d * kc + w]
    recv = _a2a(blocks)                                                    # [src][nyp][kc] == [ny][kc]
    spec = np.fft.fft(recv.reshape(ny, kc).T, axis=1)                      # transpose -> [kc][ny], c2c along y
    e_f = np.abs(spec[:kvalid] - ref[:, koff:koff + kvalid].T).max() / np.abs(ref).max()
    e_pad = np.abs(spec[kvalid:]).max() if kvalid < kc else 0.0            # padded columns stay exactly zero
    # inverse: c2c along y -> transpose ([ny][kc]: block d = rows of rank d) -> all-to-all -> k2_unpack_rows -> c2r along x
    s = np.fft.ifft(spec, axis=1) * ny
    recv = _a2a(np.ascontiguousarray(s.T).reshape(world, nyp, kc))
    rows2 = np.empty((nyp, nkr), dtype=np.complex128)
    for k in range(nkr):
        rows2[:, k] = recv[k // kc, :, k % kc]
    back = np.fft.irfft(rows2, n=nx, axis=1) / ny
    e_i = np.abs(back - c[rank * nyp:(rank + 1) * nyp]).max()
    q.put((rank, float(e_f), float(e_i), float(e_pad)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [(16, 8), (10, 12), (6, 4)])
def test_slab2d_transpose_scheme_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker2d, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_f, e_i, e_pad in res:
        assert e_f < 1e-13 and e_i < 1e-13 and e_pad == 0.0, (rank, e_f, e_i, e_pad)
