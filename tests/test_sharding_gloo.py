"""N > 1 host logic on CPU (gloo, world_size 2): the frame <-> pixel sharding exchange and the GroupNorm partial-sum
all-reduce reproduce the unsharded computation.  The CUDA kernels that realise the two index maps are checked against
the same maps in tests/test_kernels_gpu.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lavie_b200.sharding import frame_shard_ranks, scatter_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _temporal_mix(x):                       # stands in for LN -> qkv -> temporal attention -> out-proj on [F, n, C]
    w = torch.softmax(x.mean(-1, keepdim=True).transpose(0, 1) @ x.mean(-1, keepdim=True).transpose(0, 1).transpose(1, 2), -1)
    return torch.einsum("nij,jnc->inc", w, x)


def _worker(rank, world, port, F, HW, C, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    x = torch.randn(F, HW, C, dtype=torch.float64)
    P, f_loc, hwp = world, F // world, HW // world
    local = x[rank * f_loc:(rank + 1) * f_loc].reshape(f_loc * HW, C)
    # frame -> pixel sharding
    idx = scatter_rows(f_loc, HW, P)
    send = torch.empty_like(local)
    send[idx] = local
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    y = _temporal_mix(recv.reshape(F, hwp, C)).reshape(F * hwp, C)
    back = torch.empty_like(y)
    dist.all_to_all_single(back, y)
    out_local = local + back[idx]                                   # gather-add = inverse of the scatter
    want = (x + _temporal_mix(x))[rank * f_loc:(rank + 1) * f_loc].reshape(f_loc * HW, C)
    ok_a2a = torch.allclose(out_local, want, atol=1e-12)
    # GroupNorm across frames: all-reduced (sum, sumsq) == global statistics
    g = x.reshape(F * HW, 4, C // 4)
    sums = torch.stack([local.reshape(-1, 4, C // 4).sum((0, 2)), (local ** 2).reshape(-1, 4, C // 4).sum((0, 2))], -1)
    dist.all_reduce(sums)
    ok_gn = torch.allclose(sums[:, 0], g.sum((0, 2))) and torch.allclose(sums[:, 1], (g ** 2).sum((0, 2)))
    result[rank] = int(ok_a2a) + 2 * int(ok_gn)
    dist.destroy_process_group()


def test_frame_pixel_exchange_and_gn_allreduce_world2():
    world, F, HW, C = 2, 4, 6, 8
    result = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), F, HW, C, result), nprocs=world, join=True)
    assert dict(result) == {0: 3, 1: 3}


def test_scatter_rows_is_a_permutation_with_the_expected_blocks():
    f_loc, hw, p = 3, 8, 4
    idx = scatter_rows(f_loc, hw, p)
    assert sorted(idx.tolist()) == list(range(f_loc * hw))
    # chunk q of the send buffer holds pixel block q of every local frame, frame-major
    send = torch.empty(f_loc * hw, dtype=torch.long)
    send[idx] = torch.arange(f_loc * hw)
    chunk1 = send.reshape(p, f_loc, hw // p)[1]
    assert chunk1.tolist() == [[f * hw + 2, f * hw + 3] for f in range(f_loc)]


def test_rank_layout():
    p, frame_groups, pair_groups = frame_shard_ranks(8)
    assert p == 4 and frame_groups == [[0, 1, 2, 3], [4, 5, 6, 7]] and pair_groups[2] == [2, 6]


def _halo_worker(rank, world, port, result):
    """Frame-sharded (k,1,1) convolution of the VSR denoiser: pad-frame halo exchange + local conv == the conv on the
    whole video (zero padding at the video's ends only)."""
    import torch.nn.functional as F
    from lavie_b200.sharding import exchange_frame_halo
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    frames, hw, C = 8, 6, 4
    x = torch.randn(frames, hw, C, dtype=torch.float64)
    ok = 0
    for k in (3, 5):
        pad = k // 2
        w = torch.randn(C, C, k, 1, 1, dtype=torch.float64)
        want = F.conv3d(x.permute(2, 0, 1)[None, :, :, :, None], w, padding=(pad, 0, 0))[0, :, :, :, 0].permute(1, 2, 0)
        f_loc = frames // world
        buf = torch.zeros((f_loc + 2 * pad) * hw, C, dtype=torch.float64)
        buf[pad * hw:(pad + f_loc) * hw] = x[rank * f_loc:(rank + 1) * f_loc].reshape(f_loc * hw, C)
        exchange_frame_halo(buf, pad, f_loc, hw, dist.group.WORLD, world, rank)
        local = buf.reshape(f_loc + 2 * pad, hw, C)
        got = F.conv3d(local.permute(2, 0, 1)[None, :, :, :, None], w)[0, :, :, :, 0].permute(1, 2, 0)   # "valid" over frames
        ok += int(torch.allclose(got, want[rank * f_loc:(rank + 1) * f_loc], atol=1e-12))
    result[rank] = ok
    dist.destroy_process_group()


def test_frame_conv_halo_exchange_world2():
    world = 2
    result = mp.get_context("spawn").Manager().dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), result), nprocs=world, join=True)
    assert dict(result) == {0: 2, 1: 2}
