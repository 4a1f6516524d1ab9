"""Generate the committed golden vectors in tests/golden/ from the CPU oracle.

    python tools/make_golden.py

The reference ships no fixtures ("parity unpinned"), so these vectors are
the pins this build creates for itself: seeded inputs + oracle outputs for every hot-path row,
small enough to commit.  The eigen fixtures come from torch.linalg.eigh in fp64; the mixer fixture
is additionally cross-checked against transformers' MambaMixer.slow_forward in
tests/test_oracle.py.  Re-running this script must reproduce the files bit for bit.
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import mae, mamba, model, spectral, tokenizer  # noqa: E402

OUT = ROOT / "tests" / "golden"


def tokenizer_case(kind, seed, B=2, N=256, G=16, M=8):
    xyz = tokenizer.synthetic_clouds(B, N, seed, kind)
    nbr, center, org, fidx, kidx = tokenizer.group(xyz, G, M)
    return dict(kind=kind, xyz=xyz, G=G, M=M, fps_idx=fidx.int(), center=center, knn_idx=kidx.int(), nbr=nbr, org=org)


def spectral_case(seed, B, N, G, k_nn, alpha, symmetric, self_loop, binary, k, smallest, matrix, eps_mode):
    xyz = tokenizer.synthetic_clouds(B, N, seed, "surface")
    center = tokenizer.group(xyz, G, 4)[1]
    A = spectral.knn_adjacency(center, k_nn, alpha, symmetric, self_loop, binary)
    vals, vecs, allv, S = spectral.spectral_eig(center, k_nn, alpha, symmetric, self_loop, binary, k, smallest,
                                                matrix, eps_mode)
    return dict(center=center, k_nn=k_nn, alpha=alpha, symmetric=symmetric, self_loop=self_loop, binary=binary, k=k,
                smallest=smallest, matrix=matrix, eps_mode=eps_mode, adjacency=A, operator=S, vals=vals, vecs=vecs,
                all_vals=allv, perm=spectral.sast_perm(vecs).int())


def scan_case(seed, B=2, D=64, L=70, N=16):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(B, D, L, generator=g)
    delta = 0.5 * torch.randn(B, D, L, generator=g)
    z = torch.randn(B, D, L, generator=g)
    Bm = torch.randn(B, N, L, generator=g)
    Cm = torch.randn(B, N, L, generator=g)
    A = -torch.exp(torch.log(torch.arange(1, N + 1, dtype=torch.float32))[None].repeat(D, 1)
                   + 0.1 * torch.randn(D, N, generator=g))
    Dv = torch.randn(D, generator=g)
    bias = torch.rand(D, generator=g) * 4 - 6
    out = mamba.selective_scan_ref(u, delta, A, Bm, Cm, Dv, z, bias, True)
    out64 = mamba.selective_scan_fp64(u, delta, A, Bm, Cm, Dv, z, bias, True)
    w = torch.randn(D, 4, generator=g) * 0.5
    cb = torch.randn(D, generator=g) * 0.1
    conv = mamba.causal_conv1d_ref(u, w, cb, "silu")
    return dict(u=u, delta=delta, z=z, B=Bm, C=Cm, A=A, D=Dv, delta_bias=bias, out=out, out_fp64=out64.float(),
                conv_w=w, conv_b=cb, conv_out=conv)


def mixer_case(seed):
    sd = mamba.init_mamba_params(d_model=64, n_layer=2, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    h = torch.randn(2, 24, 64, generator=g)
    out = mamba.mamba_mixer({"m." + k: v for k, v in sd.items()}, "m.", h)
    return dict(params=sd, hidden=h, out=out)


def mae_case(seed, B=2, G=16, k=2, C=8):
    g = torch.Generator().manual_seed(seed)
    vecs = torch.randn(B, G, k, generator=g, dtype=torch.float64)
    perm = spectral.sast_perm(vecs)
    mask = mae.rand_mask(B, G, 0.6, seed)
    x = torch.randn(B, G, C, generator=g)
    x_vis = mae.compact_visible(x, perm, mask)
    mfull = mae.mask_full(mask, perm)
    tok = torch.randn(C, generator=g)
    full = mae.restore(x_vis, mfull, tok)
    return dict(perm=perm.int(), mask=mask, x=x, x_vis=x_vis, mask_full=mfull, mask_token=tok, x_full=full,
                x_rec=mae.gather_masked(full, mfull))


def hlt_case(seed, B=2, G=32, k=2, C=8):
    g = torch.Generator().manual_seed(seed)
    vecs = torch.randn(B, G, k, generator=g, dtype=torch.float64)
    noise = torch.rand(B, G, generator=g)
    order = spectral.hlt_order(vecs.float(), k, noise)
    x = torch.randn(B, G, C, generator=g)
    return dict(vecs=vecs.float(), noise=noise, k=k, order=order.int(), slots=spectral.hlt_slots(G, k).int(), x=x,
                out=spectral.hlt_layout(x, order, k))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save({"ball": tokenizer_case("ball", 11), "surface": tokenizer_case("surface", 12),
                "duplicates": tokenizer_case("duplicates", 13)}, OUT / "tokenizer.pt")
    torch.save({
        "cls_binary": spectral_case(21, 3, 512, 32, 8, 100.0, True, False, True, 4, True, "laplacian", "add1e-6"),
        "seg_weighted": spectral_case(22, 2, 512, 48, 10, 10.0, True, True, False, 4, True, "laplacian", "add1e-6"),
        "mae_clamp": spectral_case(23, 2, 512, 32, 12, 10.0, True, False, True, 4, True, "laplacian", "clamp1e-12"),
        "largest": spectral_case(24, 2, 512, 32, 8, 100.0, True, False, True, 3, False, "laplacian", "add1e-6"),
        "symnorm": spectral_case(25, 2, 512, 32, 8, 100.0, True, False, True, 3, True, "symmetric", "add1e-6"),
    }, OUT / "spectral.pt")
    torch.save({"a": scan_case(31), "b": scan_case(32, B=1, D=32, L=33)}, OUT / "scan_conv.pt")
    torch.save(mixer_case(41), OUT / "mixer.pt")
    torch.save(mae_case(51), OUT / "mae.pt")
    torch.save(hlt_case(61), OUT / "hlt.pt")
    for p in sorted(OUT.glob("*.pt")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
