"""Host-side description of the frame-sharding layouts, shared by the tests and the docs.

Inside one CFG half the F frames are split over P ranks (F/P consecutive frames each).  Around every temporal
attention the tokens switch to pixel sharding with one all-to-all each way:

    local tokens  [F_loc, HW, C]  --scatter-->  send [P, F_loc, HW/P, C]  --all_to_all-->  recv [P, F_loc, HW/P, C]
                                                                                            == [F, HW/P, C]
    ... LayerNorm already applied, q/k/v GEMM, temporal attention, out-projection on [F, HW/P, C] ...
    y [F, HW/P, C] == [P, F_loc, HW/P, C]  --all_to_all-->  back [P(pixel block), F_loc, HW/P, C]  --gather-add--> local

The two index maps below are exactly what `lavie_layernorm_scatter_bf16` / `lavie_add_gathered_bf16` implement.
"""
from __future__ import annotations

import torch


def scatter_rows(f_loc: int, hw: int, p: int) -> torch.Tensor:
    """dst_row[src_row]: source row (f, pixel) of the [F_loc, HW] shard -> row of the [P, F_loc, HW/P] send buffer."""
    hwp = hw // p
    f = torch.arange(f_loc).repeat_interleave(hw)
    pix = torch.arange(hw).repeat(f_loc)
    blk = pix // hwp
    return (blk * f_loc + f) * hwp + pix % hwp


def frame_shard_ranks(world: int):
    """rank -> (cfg_half, frame_shard) and the two kinds of groups used by bench.py (rank = half * P + shard)."""
    assert world >= 2 and world % 2 == 0
    p = world // 2
    frame_groups = [list(range(h * p, h * p + p)) for h in range(2)]
    pair_groups = [[s, p + s] for s in range(p)]
    return p, frame_groups, pair_groups


def exchange_frame_halo(buf, pad: int, frames_local: int, hw: int, group, P: int, idx: int) -> None:
    """Halo exchange of a frame-sharded (k,1,1) convolution (the VSR denoiser's ResnetBlock3DCNN, vsr/models/resnet.py:
    253-254,269).  ``buf`` is this rank's frame-padded map [(frames_local + 2*pad) * hw, C]: its interior holds the local
    frames (already normalised and activated), the ``pad`` frames in front / behind must become the LAST / FIRST ``pad``
    frames of the left / right neighbour; the two ends of the video keep whatever they hold (zeros = the conv's padding).
    Works on any device / back end (NCCL on the GPU path, gloo in the CPU tests)."""
    import torch.distributed as dist
    if frames_local < pad:
        raise ValueError(f"frame sharding needs at least {pad} frames per rank for this frame convolution")
    n = pad * hw
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(P))
    ops = []
    if idx > 0:
        ops.append(dist.P2POp(dist.isend, buf[n:2 * n], ranks[idx - 1], group))                      # my first frames -> left
        ops.append(dist.P2POp(dist.irecv, buf[:n], ranks[idx - 1], group))                           # left's last frames
    if idx < P - 1:
        ops.append(dist.P2POp(dist.isend, buf[n + (frames_local - pad) * hw:n + frames_local * hw], ranks[idx + 1], group))
        ops.append(dist.P2POp(dist.irecv, buf[n + frames_local * hw:], ranks[idx + 1], group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
