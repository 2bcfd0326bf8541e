"""Frame-sharding kernels on the B200 against the host-side index maps (single GPU; the 2..8-GPU equivalence run is
tools/multigpu_check.py, results in DESIGN.md)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def test_layernorm_scatter_and_add_gathered_match_index_map():
    from lavie_b200 import ops
    from lavie_b200.sharding import scatter_rows
    f_loc, hw, p, C = 4, 160, 4, 640
    g = torch.Generator().manual_seed(0)
    x = torch.randn(f_loc * hw, C, generator=g).cuda().to(torch.bfloat16)
    gamma = (torch.randn(C, generator=g) * 0.1 + 1).cuda()
    beta = (torch.randn(C, generator=g) * 0.1).cuda()
    idx = scatter_rows(f_loc, hw, p).cuda()
    got = ops.layernorm_scatter(x, gamma, beta, hw, hw // p)
    plain = ops.layernorm(x, gamma, beta)
    want = torch.empty_like(plain)
    want[idx] = plain
    assert torch.equal(got, want)
    z = torch.randn(f_loc * hw, C, generator=g).cuda().to(torch.bfloat16)      # receive buffer [P, F_loc, hw/p, C]
    out = ops.add_gathered(x, z, hw, hw // p)
    ref = (x.float() + z[idx].float()).to(torch.bfloat16)
    assert torch.equal(out, ref)


def test_groupnorm_sums_path_equals_fused_path():
    from lavie_b200 import ops
    B, rows, C = 2, 3 * 200, 640
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(B * rows, C, generator=g) * 2 + 0.3).cuda().to(torch.bfloat16)
    gamma = (torch.randn(C, generator=g) * 0.1 + 1).cuda()
    beta = (torch.randn(C, generator=g) * 0.1).cuda()
    ss_fused = ops.groupnorm_scale_shift(x, B, rows, gamma, beta, 1e-5)
    # two "shards" of the rows of each sample, summed on the host like the NCCL all-reduce would
    xs = x.reshape(B, rows, C)
    s0 = ops.groupnorm_sums(xs[:, : rows // 2].reshape(-1, C).contiguous(), B, rows // 2)
    s1 = ops.groupnorm_sums(xs[:, rows // 2:].reshape(-1, C).contiguous(), B, rows - rows // 2)
    ss = ops.groupnorm_finalize_sums(s0 + s1, C, rows * (C // 32), gamma, beta, 1e-5)
    assert rel_l2(ss, ss_fused) < 1e-6
