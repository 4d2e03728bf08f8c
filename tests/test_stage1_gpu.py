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
    vit = DVT.PretrainedViTWrapper("vit_small_patch14_dinov2.lvd142m", stride=14, allow_random_init=True)
    with torch.no_grad():
        for b in vit.model.blocks:
            b.ls1.gamma.fill_(1.0)
            b.ls2.gamma.fill_(1.0)
    vit = vit.cuda().eval()
    cfg = Stage1Config(num_iters=60, warmup_iters=6, n_levels=6, extract_bsz=4, pixel_bsz=64, graph_steps=7)
    return Stage1Pipeline(vit, layer_index=11, input_size=(70, 84), cfg=cfg)


def _run(pipe, views_list, coords, overlap):
    pipe._image_counter = 0              # per-image parameter initialisation is a function of (seed, image number)
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


# ---- end to end against the oracle (BASELINE.json configs[0]) -------------------------------------------------------
def test_stage1_config1_end_to_end_matches_oracle():
    """BASELINE config 1, the reference's own CPU-runnable case: one 224 x 224 image, DINOv2 ViT-S/14 (random init,
    non-degenerate LayerScale, position embedding resampled 37 -> 16), 8 augmented views + the image, 50-step fit, loss
    scale 1 -- `Stage1Pipeline` (views -> bank -> fit -> map, all on the GPU) against oracle ViT -> oracle fit on the same
    views, initial parameters and sampling stream.  Tolerance: cosine >= 0.999 on the bank and on the denoised map."""
    import dvt.models as DVT
    from dvt.dataset import GpuViewGenerator
    from dvt.stage1 import Stage1Config, Stage1Pipeline
    from oracle import fit as OF
    from oracle import hashgrid as HG
    from oracle import vit as OV
    ident = "vit_small_patch14_dinov2.lvd142m"
    vcfg = OV.CONFIGS[ident]
    sd = OV.random_state_dict(vcfg, seed=0)                 # LayerScale gammas resampled U(0.5, 1.5)
    vit = DVT.PretrainedViTWrapper(ident, stride=14, allow_random_init=True)
    vit.model.load_state_dict(sd)
    vit = vit.cuda().eval()
    V, T, bsz, L = 8, 50, 256, 16
    cfg = Stage1Config(num_iters=T, warmup_iters=5, n_levels=L, extract_bsz=4, pixel_bsz=bsz, loss_scale=1.0, graph_steps=5)
    pipe = Stage1Pipeline(vit, layer_index=vcfg.depth - 1, input_size=(224, 224), cfg=cfg)
    assert (pipe.h, pipe.w, pipe.C) == (16, 16, 384)
    img = torch.rand(3, 224, 224, generator=torch.Generator().manual_seed(0))
    mean, std = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    img = (img - mean) / std
    torch.manual_seed(1)
    np.random.seed(1)
    gen = GpuViewGenerator((224, 224), num_views=V)
    views, coords = gen(img.cuda())
    bank = pipe.extract_bank(views)
    # oracle bank from the SAME views (view generation has its own golden test)
    ref_bank = OV.forward_intermediates(sd, vcfg, views.cpu(), [vcfg.depth - 1], stride=14)[0].permute(0, 2, 3, 1).contiguous()
    assert F.cosine_similarity(bank.cpu().reshape(-1, 384), ref_bank.reshape(-1, 384), dim=-1).min().item() > 0.999
    meta = HG.grid_meta(L)
    init = OF.init_params(384, 16, 16, meta, seed=3)
    idx = np.random.RandomState(3).randint(0, (V + 1) * 256, (T, bsz))
    hyper = dict(lr=cfg.lr, min_lr=cfg.min_lr, weight_decay=cfg.weight_decay, warmup_iters=cfg.warmup_iters,
                 freeze_after=cfg.freeze_shared_artifacts_after, loss_scale=cfg.loss_scale)
    ora = OF.fit(ref_bank, coords.cpu(), 16, 16, meta, init, idx, **hyper)
    out = pipe.denoise(bank, coords, idx, init=init)
    losses = pipe.engine.losses()
    torch.cuda.synchronize()
    from dvt import _lib
    assert _lib.device_error() == 0
    got, ref = out["denoised_feats"].cpu(), ora["denoised_feats"]
    assert got.shape == ref.shape == (1, 16, 16, 384)
    c = F.cosine_similarity(got.reshape(-1, 384), ref.reshape(-1, 384), dim=-1).min().item()
    assert c > 0.999, f"end-to-end denoised map min cosine {c}"
    assert torch.equal(out["raw"], bank[-1])
    for row in ora["logs"]:
        s = int(row[0])
        assert abs(losses[s, 0] - row[1]) <= 0.03 * abs(row[1]) + 1e-3, f"step {s}: loss {losses[s, 0]} vs oracle {row[1]}"


def test_stage1_cli_writes_the_store(tmp_path, monkeypatch, capsys):
    """The drop-in command line itself (main_img_denoising.py) on synthetic JPEGs: two images pipelined, `.npy` store
    written by the background writer in the reference's layout, the raw map equal to the ViT features of the un-augmented
    image, and a second invocation skipping both images (resume rule, reference :303-307)."""
    import os
    import sys
    from PIL import Image
    monkeypatch.setenv("DVT_ALLOW_RANDOM_INIT", "1")        # no pretrained checkpoints on the test box
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import main_img_denoising as M
    from dvt.store import load_pair
    from dvt.utils import misc
    data_root = str(tmp_path / "data") + "/"
    rels = ["set/a.jpg", "set/b.jpg"]
    rs = np.random.RandomState(0)
    for rel in rels:
        os.makedirs(os.path.dirname(os.path.join(data_root, rel)), exist_ok=True)
        Image.fromarray(rs.randint(0, 255, (90, 120, 3), dtype=np.uint8)).save(os.path.join(data_root, rel))
    lst = tmp_path / "list.txt"
    lst.write_text("".join(f"{r}\n" for r in rels))
    argv = ["--model", "vit_small_patch14_dinov2.lvd142m", "--input_size", "70", "84", "--stride_size", "14",
            "--img_path", str(lst), "--data_root", data_root, "--save_root", str(tmp_path / "feats"), "--num_views", "6",
            "--num_iters", "40", "--warmup_iters", "4", "--n_levels", "6", "--extract_bsz", "4", "--pixel_bsz", "64",
            "--output_dir", str(tmp_path / "work"), "--seed", "3", "--collate_out", str(tmp_path / "maps.pt")]
    args = M.get_args(argv)
    M.main(args)
    text = capsys.readouterr().out
    assert text.count("Step 39/39: Loss = ") == 2 and "Saving denoised features to" in text
    for rel in rels:
        raw_p, den_p = misc.feature_paths(args, os.path.join(data_root, rel))
        assert os.path.isfile(raw_p) and os.path.isfile(den_p)
        raw, den = load_pair(den_p)
        assert raw.shape == (5, 6, 384) and den.shape == (5, 6, 384) and raw.dtype == np.float32
        assert np.isfinite(raw).all() and np.isfinite(den).all() and np.abs(den).max() > 0
        assert np.load(den_p).shape == (1, 5, 6, 384)
    packed = torch.load(str(tmp_path / "maps.pt"))
    assert packed["denoised_feats"].shape == (2, 5, 6, 384) and len(packed["files"]) == 2
    assert np.array_equal(packed["denoised_feats"][1].numpy(), load_pair(misc.feature_paths(args, os.path.join(data_root, rels[1]))[1])[1])
    # the raw map is the ViT feature map of the un-augmented image
    import dvt.models as DVT
    from dvt.dataset import load_image
    misc.fix_random_seeds(3)     # the CLI's randomly initialised backbone is seeded by fix_random_seeds(args.seed)
    vit = DVT.PretrainedViTWrapper(args.model, stride=14).cuda().eval()
    norm = vit.transformation.transforms[-1]
    x = load_image(os.path.join(data_root, rels[0]), (70, 84), norm.mean, norm.std)[None].cuda()
    feat = vit.get_intermediate_layers(x, n=[11], reshape=True)[-1].permute(0, 2, 3, 1)[0].cpu().numpy()
    raw0, _ = load_pair(misc.feature_paths(args, os.path.join(data_root, rels[0]))[1])
    assert np.abs(raw0 - feat).max() < 1e-4 * max(1.0, np.abs(feat).max())
    # second run: everything is skipped
    M.main(M.get_args(argv))
    text = capsys.readouterr().out
    assert text.count("Skipping") == 2 and "Step 39/39" not in text
