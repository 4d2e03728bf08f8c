"""Stage-1 public call on the GPU (`Stage1Pipeline.run_images`): the software-pipelined schedule over images (bank
extraction of image i+1 beside the fit of image i, two bank buffers) must give the results of the strictly sequential
schedule, from device-resident and from pinned-host views alike."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _pipe():
    import dvt.models as DVT
    from dvt.stage1 import Stage1Config, Stage1Pipeline
    torch.manual_seed(0)
    vit = DVT.PretrainedViTWrapper("vit_small_patch14_dinov2.lvd142m", stride=14)
    with torch.no_grad():
        for b in vit.model.blocks:
            b.ls1.gamma.fill_(1.0)
            b.ls2.gamma.fill_(1.0)
    vit = vit.cuda().eval()
    cfg = Stage1Config(num_iters=60, warmup_iters=6, n_levels=6, extract_bsz=4, pixel_bsz=64, graph_steps=7)
    return Stage1Pipeline(vit, layer_index=11, input_size=(70, 84), cfg=cfg)


def _run(pipe, views_list, coords, overlap):
    torch.manual_seed(7)                 # per-image module re-initialisation (torch's CUDA generator)
    pipe._gen.manual_seed(11)            # hash-table initialisation
    n_rows = coords.shape[0] * pipe.h * pipe.w

    def idx_fn(i):
        return np.random.RandomState(100 + i).randint(0, n_rows, (pipe.cfg.num_iters, pipe.cfg.pixel_bsz))

    def finalize(i, out):
        return out["denoised_feats"].cpu(), out["raw"].cpu()

    ev = []
    res = pipe.run_images(len(views_list), lambda i: views_list[i], lambda i: coords, idx_fn, finalize, events=ev,
                          overlap=overlap)
    torch.cuda.synchronize()
    assert sorted(k for k, _, _ in ev) == ["hp1"] * len(views_list) + ["hp2"] * len(views_list)
    assert all(a.elapsed_time(b) > 0 for _, a, b in ev)
    return res


def test_run_images_overlap_equals_sequential():
    from dvt import _lib
    pipe = _pipe()
    V, n_img = 7, 3
    g = torch.Generator(device="cuda").manual_seed(3)
    views = [torch.randn(V, 3, 70, 84, device="cuda", generator=g) for _ in range(n_img)]
    coords = torch.rand(V, pipe.h, pipe.w, 2, device="cuda", generator=g)
    seq = _run(pipe, views, coords, overlap=False)
    ovl = _run(pipe, views, coords, overlap=True)
    host = [v.cpu().pin_memory() for v in views]
    e2e = _run(pipe, host, coords, overlap=True)
    assert _lib.device_error() == 0
    for i in range(n_img):
        for other in (ovl, e2e):
            assert torch.equal(seq[i][1], other[i][1]), f"image {i}: raw feature map differs between schedules"
            c = F.cosine_similarity(seq[i][0].reshape(-1, pipe.C), other[i][0].reshape(-1, pipe.C), dim=-1).min().item()
            assert c > 0.999, f"image {i}: denoised map cosine {c} between schedules"
    # different images must give different maps (the two bank buffers are not mixed up)
    assert not torch.equal(seq[0][1], seq[1][1])


# ---- view generation --------------------------------------------------------------------------
def _views_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "views_small.npz"))


@pytest.mark.parametrize("which", ["reference_stream", "downscale_and_full"])
def test_view_crops_match_reference_golden(which):
    """dvt_view_crops against what the reference's own transform returned (tests/golden/make_views_golden.py):
    fp32 resampling within 2e-5 (summation order / FMA contraction), coordinates within 1 ulp of [0, 1] values."""
    from dvt import _lib, ops
    z = _views_golden()
    img = torch.from_numpy(z["image"]).cuda()
    P, S = (int(v) for v in z["patch"])
    pre = "" if which == "reference_stream" else "extra_"
    size = tuple(int(v) for v in z[pre + "size"])
    hp, wp = (size[0] - P) // S + 1, (size[1] - P) // S + 1
    views, coords = ops.view_crops(img, z[pre + "boxes"], z[pre + "flips"], size, hp, wp)
    torch.cuda.synchronize()
    assert _lib.device_error() == 0
    ref_v, ref_c = torch.from_numpy(z[pre + "views"]), torch.from_numpy(z[pre + "coords"])
    assert (views.cpu() - ref_v).abs().max().item() < 2e-5
    assert (coords.cpu() - ref_c).abs().max().item() <= 1.2e-7
    vb, _ = ops.view_crops(img, z[pre + "boxes"], z[pre + "flips"], size, hp, wp, dtype=torch.bfloat16)
    assert (vb.float().cpu() - ref_v).abs().max().item() < 4e-2     # bf16 output for the bf16 ViT path


def test_view_crops_against_oracle_at_full_size_and_errors():
    from dvt import _lib, ops
    from dvt.dataset import GpuViewGenerator
    from oracle import views as OV
    g = torch.Generator().manual_seed(5)
    img = torch.randn(3, 518, 518, generator=g)
    gen = GpuViewGenerator((518, 518), num_views=5)
    torch.manual_seed(3)
    np.random.seed(3)
    views, coords = gen(img.cuda())
    torch.cuda.synchronize()
    assert views.shape == (6, 3, 518, 518) and coords.shape == (6, 37, 37, 2)
    ref_v, ref_c = OV.make_views(img, gen.last_boxes, gen.last_flips, (518, 518))
    assert (views[:5].cpu() - ref_v).abs().max().item() < 5e-5
    assert (coords[:5].cpu() - ref_c).abs().max().item() <= 1.2e-7
    assert torch.equal(views[5].cpu(), img)                                   # the un-augmented image is the last view
    assert coords[5, 0, 0].tolist() == [0.0, 0.0] and coords[5, -1, -1].tolist() == [1.0, 1.0]
    with pytest.raises(_lib.DvtError, match="outside"):
        ops.view_crops(img.cuda(), [[500, 0, 40, 40]], [0], (64, 64), 4, 4)
    with pytest.raises(_lib.DvtError):
        gen(img)                                                                # CPU tensor: no fallback
