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


# ---------------------------------------------------------------------------------------------------------------------
# Fused 3-D engine, slab-decomposed (csrc/engine_fused3d.cu, fused3d_kernels.cuh): the producing kernels write the blocks
# of the exchange directly in the receiver's layout — P^xy as [p][kr][zl/8][ll][zl%8], A / C as [p][kr][ll][zl] — so the
# all-to-all moves contiguous blocks and nothing is packed or unpacked.  Same index formulas as the kernels, NumPy FFTs
# standing in for the in-kernel transforms.
# ---------------------------------------------------------------------------------------------------------------------
def _worker_fused3d(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = n
    nkr = nx // 2 + 1
    nzl, nyl = nz // world, ny // world
    zoff, yoff = rank * nzl, rank * nyl
    rng = np.random.default_rng(7)
    c = rng.standard_normal((nz, ny, nx))
    ref = np.fft.rfftn(c)                                                   # [m][l][kr]
    # forward: x row kernel + y-forward kernel give FFT_y FFT_x of the local planes ...
    T1 = np.fft.fft(np.fft.rfft(c[zoff:zoff + nzl], axis=2), axis=1)        # [zl][l][kr]
    # ... which k_yfwd3 stores as PXY[p][kr][zl/8][ll][zl%8] (p = l / nyl)
    send = np.empty((world, nkr, nzl // 8, nyl, 8), dtype=np.complex128)
    for zl in range(nzl):
        for l in range(ny):
            send[l // nyl, :, zl // 8, l % nyl, zl % 8] = T1[zl, l, :]
    recv = _a2a(send)                                                       # [r][kr][zl/8][ll][zl%8], r = source rank
    # the z-column kernel gathers column (kr, ll): element z lives in block r = z / nzl
    spec = np.empty((nkr, nyl, nz), dtype=np.complex128)                    # state layout [kr][ll][z]
    for z in range(nz):
        r, zl = z // nzl, z % nzl
        spec[:, :, z] = recv[r, :, zl // 8, :, zl % 8]
    spec = np.fft.fft(spec, axis=2)
    e_f = np.abs(spec - ref[:, yoff:yoff + nyl, :].transpose(2, 1, 0)).max() / np.abs(ref).max()
    # inverse: the z-column kernel stores A = IFFT_z(s) as ZA[p][kr][ll][zl] (p = z / nzl)
    A = np.fft.ifft(spec, axis=2) * nz
    sendA = np.empty((world, nkr, nyl, nzl), dtype=np.complex128)
    for z in range(nz):
        sendA[z // nzl, :, :, z % nzl] = A[:, :, z]
    recvA = _a2a(sendA)                                                     # [r][kr][ll][zl], r = l / nyl
    # k_yinv3 reads column (zl, kr): element l lives in block r = l / nyl
    Yc = np.empty((nzl, nkr, ny), dtype=np.complex128)                      # [zl][kr][l]
    for l in range(ny):
        Yc[:, :, l] = recvA[l // nyl, :, l % nyl, :].T
    Yp = np.fft.ifft(Yc, axis=2) * ny                                       # [zl][kr][y]
    back = np.fft.irfft(Yp.transpose(0, 2, 1), n=nx, axis=2) / (ny * nz)    # x row kernel: c2r along x
    e_i = np.abs(back - c[zoff:zoff + nzl]).max()
    q.put((rank, float(e_f), float(e_i)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [(8, 4, 16), (6, 8, 32)])
def test_fused3d_slab_layouts_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fused3d, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_f, e_i in res:
        assert e_f < 1e-13 and e_i < 1e-13, (rank, e_f, e_i)


@pytest.mark.parametrize("P", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("ny,nz", [(64, 64), (128, 256), (1024, 128)])
def test_fused3d_flattened_strides_match_the_block_layouts(P, ny, nz):
    """engine_fused3d.cu y3args(): element l = t + T*e of a y column is addressed as base(t) + e*s_e + (e >> esh)*s_r.
    Must equal the block layouts [r][kr][ll][zl] (k_yinv3 input) and [p][kr][zl/8][ll][zl%8] (k_yfwd3 output)."""
    nkx = 5
    nyl, nzl = ny // P, nz // P
    if nzl < 8:
        pytest.skip("each rank needs at least 8 planes")
    T = ny // 16
    blk = nkx * nyl * nzl
    esh = (16 // P).bit_length() - 1
    in_se, in_sr = T * nzl, blk - nyl * nzl
    out_se, out_sr = T * 8, blk - nyl * 8
    rng = np.random.default_rng(P + ny)
    for _ in range(20):
        kr, t, e = int(rng.integers(nkx)), int(rng.integers(T)), int(rng.integers(16))
        zl = 2 * int(rng.integers(nzl // 2))
        l = t + T * e
        r, ll = l // nyl, l % nyl
        want_in = ((r * nkx + kr) * nyl + ll) * nzl + zl
        got_in = (kr * nyl + t) * nzl + zl + e * in_se + (e >> esh) * in_sr
        assert want_in == got_in
        want_out = (((r * nkx + kr) * (nzl // 8) + zl // 8) * nyl + ll) * 8 + zl % 8
        got_out = ((kr * (nzl // 8) + zl // 8) * nyl + t) * 8 + zl % 8 + e * out_se + (e >> esh) * out_sr
        assert want_out == got_out


@pytest.mark.parametrize("P", [1, 2, 8])
def test_fused3d_z_column_offsets_match_the_block_layouts(P):
    """fused_kernels.cuh, k_fused_y<..., D3>: element z of column cid = kr*nyl + ll is gathered from block r = z >> zsh at
    pin + (zl >> 3)*nyl*8 + (zl & 7) and its A / C value stored into block p = z >> zsh at (cid << zsh) + zl."""
    nkx, ny, nz = 5, 64, 128
    nyl, nzl = ny // P, nz // P
    zsh = nzl.bit_length() - 1
    blk = nkx * nyl * nzl
    rng = np.random.default_rng(P)
    for _ in range(50):
        kr, ll, z = int(rng.integers(nkx)), int(rng.integers(nyl)), int(rng.integers(nz))
        cid = kr * nyl + ll
        r, zl = z >> zsh, z & (nzl - 1)
        assert (r, zl) == (z // nzl, z % nzl)
        # gather: Psrc[r] = base + r*blk ; [r][kr][zl/8][ll][zl%8]
        got = r * blk + (kr * (nzl >> 3) * nyl + ll) * 8 + (zl >> 3) * nyl * 8 + (zl & 7)
        want = (((r * nkx + kr) * (nzl // 8) + zl // 8) * nyl + ll) * 8 + zl % 8
        assert got == want
        # store: Adst[p] = base + p*blk ; [p][kr][ll][zl]
        got = r * blk + (cid << zsh) + zl
        want = ((r * nkx + kr) * nyl + ll) * nzl + zl
        assert got == want
