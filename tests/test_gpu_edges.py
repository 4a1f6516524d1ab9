"""Edge cases of the C-ABI kernels on the GPU: maximum sizes of the sweep config (C5), ragged / degenerate shapes, and the
error behaviour of the boundary (every misuse must come back as a Python exception carrying sim_last_error_string(),
never as a crash or a silent fallback)."""

import pytest
import torch

from oracle import mamba, spectral, tokenizer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(lib):
    from si_mamba_b200 import ops as o
    return o


def scan_inputs(B, D, L, seed, N=16):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(B, D, L, generator=g)
    delta = 0.5 * torch.randn(B, D, L, generator=g)
    z = torch.randn(B, D, L, generator=g)
    Bm, Cm = torch.randn(B, N, L, generator=g), torch.randn(B, N, L, generator=g)
    A = -torch.exp(torch.log(torch.arange(1, N + 1, dtype=torch.float32))[None].repeat(D, 1)
                   + 0.2 * torch.randn(D, N, generator=g))
    Dv = torch.randn(D, generator=g)
    dt = torch.exp(torch.rand(D, generator=g) * 4.6 - 6.9).clamp(min=1e-4)
    bias = dt + torch.log(-torch.expm1(-dt))   # upstream dt-bias init
    return u, delta, A, Bm, Cm, Dv, z, bias


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


def test_scan_longest_sequence(ops):
    """L = 4096 (the longest sequence of the C5 sweep): 256 tiles, the state must survive 4096 multiplications."""
    B, D, L = 1, 64, 4096
    u, delta, A, Bm, Cm, Dv, z, bias = scan_inputs(B, D, L, 4096)
    ref = mamba.selective_scan_fp64(u, delta, A, Bm, Cm, Dv, z, bias, True)
    tm = lambda t: t.transpose(1, 2).contiguous().cuda()
    out = ops.selective_scan_tm(tm(u), tm(delta), A.cuda(), tm(Bm), tm(Cm), Dv.cuda(), tm(z), bias.cuda(), True)
    assert rel_err(out.cpu().transpose(1, 2).double(), ref) < 1e-3


def test_scan_huge_steps_do_not_overflow(ops):
    """delta so large that exp(delta * A) underflows to 0 and softplus takes its linear branch: finite, h = b exactly."""
    B, D, L = 1, 64, 20
    u = torch.randn(B, L, D).cuda()
    delta = torch.full((B, L, D), 60.0).cuda()
    A = -torch.arange(1, 17, dtype=torch.float32).repeat(D, 1).cuda()
    Bm, Cm = torch.randn(B, L, 16).cuda(), torch.randn(B, L, 16).cuda()
    out = ops.selective_scan_tm(u, delta, A, Bm, Cm, None, None, None, True)
    ref = 60.0 * u * (Bm * Cm).sum(-1, keepdim=True)  # every step forgets the past completely
    assert torch.isfinite(out).all() and rel_err(out, ref) < 1e-5


def test_fps_largest_cloud_and_single_group(ops):
    xyz = tokenizer.synthetic_clouds(1, 16384, 5, "ball")
    center, idx = ops.fps(xyz.cuda(), 8)
    ref = tokenizer.fps(xyz, 8)
    assert torch.equal(idx.cpu().long(), ref)
    c1, i1 = ops.fps(xyz.cuda(), 1)
    assert i1.cpu().tolist() == [[0]] and torch.equal(c1.cpu()[0, 0], xyz[0, 0])


def test_spectral_largest_graph(ops):
    """G = 512 patches (upper end of the north-star's 64-512 range, global-memory workspace path): eigenvalues 1e-5."""
    xyz = tokenizer.synthetic_clouds(1, 4096, 11, "surface")
    center = tokenizer.group(xyz, 512, 4)[1]
    vals, vecs, allv, S = spectral.spectral_eig(center, 20, 10.0, True, False, True, 4, True)
    out = ops.spectral_eig(center.cuda(), 20, 10.0, True, False, True, 4, True)
    assert torch.allclose(out["vals"].cpu().double(), vals.double(), rtol=1e-5, atol=1e-6)
    perm = out["perm"].cpu().long()
    assert torch.equal(perm.sort(-1).values, torch.arange(512).expand_as(perm))  # every order is a permutation


def test_boundary_errors_are_exceptions(ops):
    """Misuse of the ABI: CPU tensors, empty problems, unsupported sizes, misaligned views."""
    from si_mamba_b200._lib import SimError
    x = torch.randn(2, 64, 3)
    with pytest.raises(RuntimeError):
        ops.fps(x, 8)                                  # CPU tensor: there is no CPU fallback
    with pytest.raises(SimError):
        ops.fps(x.cuda(), 65)                          # more groups than points
    with pytest.raises(SimError):
        ops.fps(torch.randn(1, 20000, 3).cuda(), 8)    # above the built maximum
    u = torch.randn(1, 8, 64).cuda()
    A = -torch.ones(64, 16).cuda()
    bc = torch.randn(1, 8, 16).cuda()
    with pytest.raises(SimError):                      # D must be a multiple of 16
        ops.selective_scan_tm(u[..., :40].contiguous(), u[..., :40].contiguous(), A[:40], bc, bc)
    wide = torch.randn(1, 8, 70).cuda()
    # a slice whose base is not 16-byte aligned cannot be a TMA tensor: the wrapper repacks it, the result is unchanged
    a = ops.selective_scan_tm(wide[..., 1:65], u.abs() * 0.1, A, bc, bc)
    b = ops.selective_scan_tm(wide[..., 1:65].contiguous(), u.abs() * 0.1, A, bc, bc)
    assert torch.equal(a, b)
    with pytest.raises(SimError):
        ops.linear_split3(torch.zeros(3, 8, 36, dtype=torch.bfloat16).cuda(), torch.zeros(3, 8, 36, dtype=torch.bfloat16).cuda(), 36)
    # the error text comes from the library
    try:
        ops.fps(x.cuda(), 65)
    except SimError as e:
        assert "fps" in str(e)
