"""Pins oracle/fit.py + oracle/hashgrid.py against the golden fixtures produced with the reference's own
SingleImageDenoiser / adjust_learning_rate (tests/golden/make_fit_golden.py), and checks the host-side pieces the
CUDA path shares with it (level table, schedule, coordinate conventions)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import fit as OF
from oracle import hashgrid as HG

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _golden(name):
    z = np.load(os.path.join(GOLD, f"fit_{name}.npz"))
    cfg = {k: v for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    for k in ("C", "h", "w", "V", "bsz", "n_levels", "num_iters", "warmup_iters", "log_every", "seed"):
        cfg[k] = int(cfg[k])
    return cfg, z


@pytest.mark.parametrize("name", ["small_L6_ls1", "hashed_L16"])
def test_oracle_reproduces_reference_golden(name):
    cfg, z = _golden(name)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    meta = HG.grid_meta(cfg["n_levels"])
    feats, coords = OF.synthetic_bank(cfg["V"], cfg["h"], cfg["w"], cfg["C"], seed=cfg["seed"])
    init = OF.init_params(cfg["C"], cfg["h"], cfg["w"], meta, seed=cfg["seed"])
    idx = np.random.RandomState(cfg["seed"]).randint(0, cfg["V"] * cfg["h"] * cfg["w"], (cfg["num_iters"], cfg["bsz"]))
    assert int(idx.sum()) == int(z["idx_checksum"][0])
    out = OF.fit(feats, coords, cfg["h"], cfg["w"], meta, init, idx, lr=cfg["lr"], min_lr=cfg["min_lr"],
                 weight_decay=cfg["weight_decay"], warmup_iters=cfg["warmup_iters"], freeze_after=cfg["freeze_after"],
                 loss_scale=cfg["loss_scale"], log_every=cfg["log_every"])
    assert np.allclose(out["logs"], z["logs"], rtol=2e-3, atol=1e-5)
    d = (out["denoised_feats"] - torch.from_numpy(z["denoised_feats"])).abs().max().item()
    assert d < 2e-3, d
    assert (out["params"]["G"] - torch.from_numpy(z["G_final"])).abs().max().item() < 1e-3


def test_level_table_matches_survey():
    m = HG.grid_meta(16)
    assert list(map(int, m.res)) == [16, 22, 28, 37, 49, 65, 85, 112, 148, 195, 257, 338, 446, 589, 777, 1025]
    assert m.n_entries == 2467720 and m.n_params == 19741760
    assert list(m.hashed) == [False] * 15 + [True]
    m10 = HG.grid_meta(10)
    assert list(map(int, m10.res)) == [16, 26, 41, 64, 102, 162, 256, 407, 646, 1024]
    assert not m10.hashed.any() and m10.n_params == 13923712


def test_product_level_table_equals_oracle():
    from dvt.models.hashgrid_meta import make_meta
    for L in (1, 2, 6, 10, 16):
        a, b = make_meta(L), HG.grid_meta(L)
        assert np.array_equal(a.scale, b.scale) and np.array_equal(a.res, b.res) and np.array_equal(a.size, b.size)
        assert np.array_equal(a.offset, b.offset) and np.array_equal(a.hashed.astype(bool), b.hashed)


def test_hashgrid_properties():
    meta = HG.grid_meta(16)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(meta.n_entries, 8, generator=g)
    c = torch.rand(512, 2, generator=g)
    # interpolation weights are a partition of unity and indices stay inside their level
    for l in range(16):
        idx, w = HG.corner_indices_weights(c, meta, l)
        assert torch.allclose(w.sum(1), torch.ones(512), atol=1e-6)
        assert int(idx.min()) >= 0 and int(idx.max()) < int(meta.size[l])
    # linearity in the table
    e1 = HG.encode(table, c, meta)
    e2 = HG.encode(2.5 * table, c, meta)
    assert torch.allclose(e2, 2.5 * e1, rtol=1e-5, atol=1e-6)
    # a constant table encodes to that constant (weights sum to one)
    e3 = HG.encode(torch.full((meta.n_entries, 8), 0.75), c, meta)
    assert torch.allclose(e3, torch.full_like(e3, 0.75), atol=1e-5)


def test_lr_schedule_and_coordinates():
    # dvt/utils/misc.py:306-322
    assert OF.lr_at(0, 0.01, 0.001, 200, 2000) == 0.0
    assert math.isclose(OF.lr_at(100, 0.01, 0.001, 200, 2000), 0.005)
    assert math.isclose(OF.lr_at(200, 0.01, 0.001, 200, 2000), 0.01)
    assert math.isclose(OF.lr_at(2000, 0.01, 0.001, 200, 2000), 0.001)
    c = OF.make_patch_coordinates(3, 5)
    assert c.shape == (3, 5, 2) and float(c[0, 0, 0]) == -1 and float(c[0, 4, 0]) == 1 and float(c[2, 0, 1]) == 1
    from dvt.fit import make_patch_coordinates
    assert torch.equal(make_patch_coordinates(37, 37, 0, 1), OF.make_patch_coordinates(37, 37, 0, 1))


def test_patch_cell_indexing_is_exact():
    """Row r of the flattened bank belongs to noise-map cell r % (h*w): the reference's tiled linspace(-1,1)
    G-coordinates (main_img_denoising.py:58-62) sample exactly that node of G (bit-exact patch indexing)."""
    h, w, V, C = 5, 7, 3, 4
    G = torch.randn(1, C, h, w, generator=torch.Generator().manual_seed(0))
    gc = OF.make_patch_coordinates(h, w).unsqueeze(0).repeat(V, 1, 1, 1).reshape(-1, 2)
    rows = torch.arange(V * h * w)
    s = torch.nn.functional.grid_sample(G, gc[None, None, ...], mode="bilinear", align_corners=True)
    s = s.squeeze().permute(1, 0)
    cell = rows % (h * w)
    direct = G[0].permute(1, 2, 0).reshape(h * w, C)[cell]
    assert torch.allclose(s, direct, atol=1e-5)


@pytest.mark.parametrize("phase2", [False, True])
def test_module_forward_equals_oracle(phase2):
    """`dvt.models.SingleImageDenoiser.forward` (the drop-in module, written from its own pieces) against the oracle
    restatement of offline_denoiser.py:62-171 -- training (2-D) call incl. off-node artifact coordinates (bilinear,
    align_corners=True, zero padding) and the 4-D visualisation call, both phases, values and gradients."""
    import dvt.models as DVT
    C, h, w, n = 32, 5, 6, 64
    g = torch.Generator().manual_seed(3)
    meta = HG.grid_meta(4)
    init = OF.init_params(C, h, w, meta, seed=3)
    p = {k: v.clone().requires_grad_(True) for k, v in init.items()}
    den = DVT.SingleImageDenoiser(h, w, C)
    with torch.no_grad():
        den.shared_artifacts.copy_(init["G"])
        for i in (0, 2, 4):
            den.residual_predictor[i].weight.copy_(init[f"res.{i}.weight"])
            den.residual_predictor[i].bias.copy_(init[f"res.{i}.bias"])
    if phase2:
        den.start_residual_predictor()
    raw = torch.randn(n, C, generator=g)
    pc = torch.rand(n, 2, generator=g)
    gc = torch.rand(n, 2, generator=g) * 2.2 - 1.1          # some samples fall outside [-1, 1]: zero padding
    gc[: n // 2] = OF.make_patch_coordinates(h, w).reshape(-1, 2)[torch.randint(0, h * w, (n // 2,), generator=g)]
    field = lambda c: OF.field_forward(p, c, meta)  # noqa: E731
    got = den(raw, pc, neural_field=field, shared_artifact_coords=gc)
    want = OF.denoiser_forward(p, raw, pc, meta, gc, phase2)
    assert set(got) == set(want)
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-6), k
    got["loss"].backward()
    gG = den.shared_artifacts.grad.clone()
    want["loss"].backward()
    assert torch.allclose(gG, p["G"].grad, rtol=1e-4, atol=1e-7)
    if phase2:
        assert torch.allclose(den.residual_predictor[4].weight.grad, p["res.4.weight"].grad, rtol=1e-4, atol=1e-7)
    # 4-D call: one map per leading index, G cell by cell
    raw_hw = torch.randn(1, h, w, C, generator=g)
    co_hw = OF.make_patch_coordinates(h, w, 0, 1)[None]
    with torch.no_grad():
        vis = den(raw_hw, co_hw, neural_field=field, return_visualization=True)
        q = OF.query({k: v.detach() for k, v in p.items()}, raw_hw, co_hw, meta, phase2)
    for k in ("denoised_feats", "shared_patterns", "denoised_features") + (("pred_residual",) if phase2 else ()):
        assert vis[k].shape == (1, h, w, C) and torch.allclose(vis[k], q[k], rtol=1e-5, atol=1e-6), k
    with pytest.raises(AssertionError):
        den(raw, pc, neural_field=field)                      # 2-D call without artifact coordinates
    with pytest.raises(AssertionError):
        den(raw, pc, neural_field=field, shared_artifact_coords=gc, return_visualization=True)


@pytest.mark.parametrize("h,w", [(37, 37), (16, 16), (5, 7)])
def test_artifact_axis_table_reproduces_grid_sample(h, w):
    """The per-node tables the CUDA loss kernel uses (dvt.fit.artifact_axis_table) reproduce ATen's
    F.grid_sample(bilinear, align_corners=True) at the reference's linspace(-1, 1) nodes BIT for bit: the interpolated
    value and the gradient every cell receives (for 37 x 37, 818 of the 1369 cells get a gradient != 1)."""
    from dvt.fit import artifact_axis_table
    C = 3
    g = torch.Generator().manual_seed(0)
    G = torch.randn(1, C, h, w, generator=g, requires_grad=True)
    nodes = OF.make_patch_coordinates(h, w).reshape(-1, 2)
    out = torch.nn.functional.grid_sample(G, nodes[None, None], mode="bilinear", align_corners=True)
    out = out.squeeze(0).squeeze(1).t()                               # [h*w, C]
    up = torch.randn(h * w, C, generator=g)
    (out * up).sum().backward()
    xi, xw0, xw1 = artifact_axis_table(w)
    yi, yw0, yw1 = artifact_axis_table(h)
    cells = G.detach()[0].permute(1, 2, 0).reshape(h * w, C)
    val = torch.zeros(h * w, C)
    grad = torch.zeros(h * w, C)
    for r in range(h):
        for c in range(w):
            acc = torch.zeros(C)
            for k in range(4):                                        # nw, ne, sw, se
                xx, yy = int(xi[c]) + (k & 1), int(yi[r]) + (k >> 1)
                wgt = (xw1[c] if k & 1 else xw0[c]) * (yw1[r] if k >> 1 else yw0[r])
                if wgt != 0 and 0 <= xx < w and 0 <= yy < h:
                    acc = acc + cells[yy * w + xx] * wgt
                    grad[yy * w + xx] += up[r * w + c] * wgt
            val[r * w + c] = acc
    # values: the same four products; ATen's vectorised kernel may contract / order the sum differently (<= 1 ulp)
    assert torch.allclose(val, out.detach(), rtol=0, atol=1e-6 * float(cells.abs().max()))
    assert (val == out.detach()).float().mean().item() > 0.8
    ref_grad = G.grad[0].permute(1, 2, 0).reshape(h * w, C)
    assert torch.allclose(grad, ref_grad, rtol=0, atol=1e-6 * float(up.abs().max()))
    # the SAME cells receive gradient: where grid_sample leaks into a neighbour, so do the tables (and nowhere else)
    plain = torch.zeros(h * w, C)
    plain += up                                                       # weight-1 attribution to the node's own cell
    leak_ref, leak_tab = (ref_grad != plain), (grad != plain)
    assert torch.equal(leak_ref.any(1), leak_tab.any(1))
    if (h, w) == (37, 37):
        ones = torch.zeros(h * w)
        for r in range(h):
            for c in range(w):
                for k in range(4):
                    xx, yy = int(xi[c]) + (k & 1), int(yi[r]) + (k >> 1)
                    wgt = (xw1[c] if k & 1 else xw0[c]) * (yw1[r] if k >> 1 else yw0[r])
                    if wgt != 0 and 0 <= xx < w and 0 <= yy < h:
                        ones[yy * w + xx] += wgt
        assert int((ones != 1).sum()) == 818
