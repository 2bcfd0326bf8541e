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
