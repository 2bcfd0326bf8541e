"""Peer-memory plumbing for the fused frame-sharding kernels (csrc/p2p.cu).

torch's symmetric-memory allocator is used ONLY to obtain buffers that every rank of the frame group can address
(CUDA IPC mappings) and their peer pointers; all loads / stores / flags on those buffers are issued by this repo's
kernels.  One `PeerContext` per frame group.
"""
from __future__ import annotations

import ctypes
from typing import List

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .ops import _Launch, _rows2d, _stream, BF16, F32


import os as _os
# Launch-count switches, both OFF by default because they measured SLOWER on 4 x B200 (profiles/r2_ab_p2p.txt, same box,
# ms per step: neither 9.12 | fused barriers 9.29 | one-launch GroupNorm exchange 10.01 | both 10.17):
#  * FUSED_BARRIERS: the flag barrier runs inside the scatter / halo-push (last block) and gather-add (every block polls in
#    its prologue) kernels instead of a separate one-warp lavie_rank_barrier launch (-32 launches per step);
#  * GN_ONE_LAUNCH: the single-block GroupNorm exchange kernel folds the producers' micro-group statistics itself
#    instead of a 64-block reduction launch in front of it (-45 launches, but one block reading up to 160 KB).
# The sharded step is bound by the latency of its ~500 dependent small kernels, not by their count.
FUSED_BARRIERS = _os.environ.get("LAVIE_P2P_FUSED_BARRIERS", "0") != "0"
GN_ONE_LAUNCH = _os.environ.get("LAVIE_P2P_GN_ONE_LAUNCH", "0") != "0"


class PeerContext:
    """Symmetric buffers of one frame group: flags, GroupNorm slots, and the two token exchange buffers."""

    def __init__(self, group, token_bytes: int, device: torch.device, frame_off: int = -1, halo_bytes: int = 0):
        """token_bytes: size of one all-to-all buffer (all frames x this rank's pixel slice x C at level 0);
        frame_off: global index of this rank's first frame (-1 = equal shards); halo_bytes: size of the
        [2 halo frames | local frames] q|k|v buffer of the interpolation model's SparseCausal attention (0 = none)."""
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.P = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.frame_off = frame_off
        flags_b = 256                                     # uint32[P], padded
        slots_b = 2 * self.P * (2 * 64 * 2) * 8           # double[2][P][samples<=2 * groups<=64 * 2]
        tok_b = (token_bytes + 255) // 256 * 256
        halo_b = (halo_bytes + 255) // 256 * 256
        self.layout = {"flags": 0, "slots": flags_b, "recv": flags_b + slots_b, "y": flags_b + slots_b + tok_b,
                       "kvx": flags_b + slots_b + 2 * tok_b}
        self.halo_bytes = halo_b
        total = flags_b + slots_b + 2 * tok_b + halo_b
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(self.buf, group)
        self.peer_base: List[int] = [int(p) for p in self.handle.buffer_ptrs]
        assert len(self.peer_base) == self.P and self.peer_base[self.rank] == self.buf.data_ptr()
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)     # "last block" counter of the fused barriers
        # pinned host words the kernels fill before trapping on a lost peer (readable after the context died)
        self.fault = torch.zeros(8, dtype=torch.int32).pin_memory()
        _lib.load().lavie_p2p_fault_buffer(self.fault.data_ptr(), 0)
        self.token_bytes = tok_b
        dist.barrier(group)                               # everybody has zeroed flags before the first kernel

    def describe_fault(self) -> str:
        """After a CUDA error on a sharded step: which peer this rank was waiting for (one rank's failure aborts the
        whole frame group, so the rank that reports a missing peer points at the rank that failed first)."""
        w = [int(v) & 0xFFFFFFFF for v in self.fault.tolist()]
        if w[0] != 0x4C564945:
            return "no peer-wait timeout recorded on this rank"
        return (f"rank {w[1]} of the frame group timed out waiting for peer {w[2]}: wanted epoch {w[3]}, "
                f"last seen {w[4]}")

    def _sync_args(self):
        """(flag_ptrs, epoch, ticket) of a fused barrier, or NULLs when the barrier is a separate launch."""
        if FUSED_BARRIERS:
            return self.ptrs("flags"), self.epoch.data_ptr(), self.ticket.data_ptr()
        return None, None, None

    def ptrs(self, what: str):
        off = self.layout[what]
        return (ctypes.c_void_p * self.P)(*[b + off for b in self.peer_base])

    def local(self, what: str, rows: int, cols: int) -> torch.Tensor:
        """bf16 [rows, cols] view of this rank's `recv` / `y` / `kvx` buffer."""
        off = self.layout[what]
        n = rows * cols * 2
        assert n <= (self.halo_bytes if what == "kvx" else self.token_bytes), (what, n)
        return self.buf[off:off + n].view(BF16).view(rows, cols)

    def push_halo(self, ext: torch.Tensor, n_local: int, hw: int):
        """ext = this rank's `kvx` view [(2 + n_local) * hw, cols] whose local frames (rows 2*hw ..) were just written:
        store frame 0 into every rank's halo block 0 (first rank) and my last frame into my right neighbour's block 1,
        then the flag barrier; afterwards ext's halo rows are valid on every rank."""
        lib = _lib.load()
        cols = ext.shape[1]
        frame_bytes = hw * cols * 2
        first = ext[2 * hw:3 * hw]
        last = ext[(1 + n_local) * hw:(2 + n_local) * hw]
        with _Launch("lavie_halo_push_p2p", 0.0, 2.0 * frame_bytes):      # flag barrier fused into the kernel's last block
            check(lib.lavie_halo_push_p2p(first.data_ptr(), last.data_ptr(), frame_bytes, self.ptrs("kvx"), self.P,
                                          self.rank, *self._sync_args(), _stream()), "lavie_halo_push_p2p")
        if not FUSED_BARRIERS:
            self.barrier()

    # ---- fused kernels ----
    def barrier(self):
        lib = _lib.load()
        with _Launch("lavie_rank_barrier"):
            check(lib.lavie_rank_barrier(self.ptrs("flags"), self.epoch.data_ptr(), self.P, self.rank, _stream()),
                  "lavie_rank_barrier")

    def gn_scale_shift(self, x, x2, samples, rows_local, gamma, beta, eps, groups=32, rows_global=None):
        """rows_global: rows of one sample over ALL ranks (default rows_local * P: equal shards)."""
        lib = _lib.load()
        rows_global = rows_local * self.P if rows_global is None else rows_global
        rows, c0, ld0 = _rows2d(x)
        c1, ld1 = 0, 0
        if x2 is not None:
            _, c1, ld1 = _rows2d(x2)
        C = c0 + c1
        from . import ops
        cs = ops.colsums(x, rows_local, x2)
        if cs is not None and not GN_ONE_LAUNCH:
            sums = ops.groupnorm_sums(x, samples, rows_local, groups, x2)      # multi-block reduction of the statistics
            ss = torch.empty((samples, C, 2), dtype=F32, device=x.device)
            with _Launch("lavie_gn_exchange_finalize"):
                check(lib.lavie_gn_exchange_finalize_sums(sums.data_ptr(), samples, groups, C,
                                                          rows_global * (C // groups), gamma.data_ptr(),
                                                          beta.data_ptr(), eps, ss.data_ptr(), self.ptrs("slots"),
                                                          self.ptrs("flags"), self.epoch.data_ptr(), self.P, self.rank,
                                                          _stream()), "lavie_gn_exchange_finalize_sums")
            return ss
        if cs is not None:
            # the statistics pass already happened in the producers' epilogues: local reduction of their micro-group
            # sums, peer exchange and finalize in ONE single-block launch
            ss = torch.empty((samples, C, 2), dtype=F32, device=x.device)
            with _Launch("lavie_gn_exchange_finalize"):
                check(lib.lavie_gn_exchange_finalize_colsums(cs[0].data_ptr(), c0, cs[1].data_ptr() if cs[1] is not None
                                                             else None, c1, samples, rows_local, groups,
                                                             rows_global * (C // groups), gamma.data_ptr(),
                                                             beta.data_ptr(), eps, ss.data_ptr(), self.ptrs("slots"),
                                                             self.ptrs("flags"), self.epoch.data_ptr(), self.P, self.rank,
                                                             _stream()), "lavie_gn_exchange_finalize_colsums")
            return ss
        chunks = lib.lavie_groupnorm_chunks(samples, rows_local)
        partial = torch.empty((samples, chunks, groups, 2), dtype=F32, device=x.device)
        with _Launch("lavie_groupnorm_stats", 0.0, 2.0 * rows * C, f"gn_stats rows={rows} C={C} samples={samples}"):
            check(lib.lavie_groupnorm_stats(x.data_ptr(), ld0, c0, x2.data_ptr() if x2 is not None else None, ld1, c1,
                                            samples, rows_local, groups, partial.data_ptr(), _stream()),
                  "lavie_groupnorm_stats")
        ss = torch.empty((samples, C, 2), dtype=F32, device=x.device)
        with _Launch("lavie_gn_exchange_finalize"):
            check(lib.lavie_gn_exchange_finalize(partial.data_ptr(), samples, chunks, groups, C,
                                                 rows_global * (C // groups), gamma.data_ptr(), beta.data_ptr(),
                                                 eps, ss.data_ptr(), self.ptrs("slots"), self.ptrs("flags"),
                                                 self.epoch.data_ptr(), self.P, self.rank, _stream()),
                  "lavie_gn_exchange_finalize")
        return ss

    def layernorm_scatter(self, x, gamma, beta, hw, eps=1e-5, frames_total=None):
        """LayerNorm + all-to-all (store side) + flag barrier (fused: the kernel's last block).  Returns this rank's
        receive buffer [F, hw/P, C], valid once the kernel has completed."""
        lib = _lib.load()
        rows, C, ldx = _rows2d(x)
        hwp = hw // self.P
        with _Launch("lavie_layernorm_scatter_p2p", 0.0, 4.0 * rows * C, f"ln_scatter_p2p rows={rows} C={C}"):
            check(lib.lavie_layernorm_scatter_p2p(x.data_ptr(), ldx, gamma.data_ptr(), beta.data_ptr(), eps,
                                                  self.ptrs("recv"), rows, C, hw, hwp, self.P, self.rank, self.frame_off,
                                                  *self._sync_args(), _stream()), "lavie_layernorm_scatter_p2p")
        if not FUSED_BARRIERS:
            self.barrier()
        frames_total = rows // hw * self.P if frames_total is None else frames_total
        return self.local("recv", frames_total * hwp, C)

    def add_gathered(self, res, hw):
        """res + all-to-all back (load side) of the peers' `y` buffers."""
        lib = _lib.load()
        rows, C, ldr = _rows2d(res)
        hwp = hw // self.P
        if not FUSED_BARRIERS:
            self.barrier()                                # every rank's y is complete
        out = torch.empty((rows, C), dtype=BF16, device=res.device)
        with _Launch("lavie_add_gathered_p2p", 0.0, 6.0 * rows * C):
            check(lib.lavie_add_gathered_p2p(res.data_ptr(), ldr, self.ptrs("y"), out.data_ptr(), C, rows, C, hw, hwp,
                                             self.P, self.rank, self.frame_off, *self._sync_args(), _stream()),
                  "lavie_add_gathered_p2p")
        return out
