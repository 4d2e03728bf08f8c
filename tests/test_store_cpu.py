"""CPU: the host-side contract around the hot paths against goldens minted from the reference's own dvt/utils/misc.py
(tests/golden/make_store_golden.py): `.npy` feature-store paths / resume check, the LR schedule (8a-4),
and the files the stage-1 driver writes (main_img_denoising.py:131-146: raw (h, w, C) and denoised (1, h, w, C), float32)."""
import json
import os
import sys
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "store_and_schedule.json")))


def test_feature_store_paths_match_reference():
    from dvt.utils import misc
    for c in GOLD["paths"]:
        args = Namespace(data_root=c["data_root"], save_root=c["save_root"], model=c["model"])
        assert misc.feature_paths(args, c["filename"]) == (c["raw"], c["denoised"])


def test_lr_schedule_matches_reference():
    from dvt.utils import misc
    from oracle import fit as OF
    for s in GOLD["schedules"]:
        a = Namespace(lr=s["lr"], min_lr=s["min_lr"], warmup_iters=s["warmup_iters"], num_iters=s["num_iters"])

        class Opt:
            param_groups = [{"lr": None}, {"lr": None, "lr_scale": 0.5}]
        for it, ref in zip(s["iterations"], s["values"]):
            o = Opt()
            got = misc.adjust_learning_rate(o, it, a)
            assert abs(got - ref) <= 1e-15 * max(1.0, abs(ref)) + 1e-18, (s, it, got, ref)
            assert o.param_groups[0]["lr"] == got and o.param_groups[1]["lr"] == got * 0.5
            assert abs(OF.lr_at(it, s["lr"], s["min_lr"], s["warmup_iters"], s["num_iters"]) - ref) <= 1e-15


def test_stage1_driver_writes_the_reference_store(tmp_path):
    """`denoise_an_image` of the drop-in CLI with a stub pipeline: file locations, dtypes and shapes of the store, and the
    resume check that makes a second run skip the image."""
    sys.path.insert(0, ROOT)
    import main_img_denoising as M
    from dvt.utils import misc
    h, w, C, V = 4, 5, 32, 3

    class Engine:
        def losses(self):
            return np.zeros((20, 5), np.float32)

    class Pipe:
        engine = Engine()

        def denoise(self, feats, coords, idx):
            assert idx.shape == (20, 8) and idx.max() < V * h * w
            return {"denoised_feats": torch.ones(1, h, w, C), "raw": feats[-1], "denoiser": None}

    data_root = str(tmp_path / "imgs")
    args = Namespace(num_iters=20, pixel_bsz=8, lr=0.01, min_lr=0.001, warmup_iters=2, data_root=data_root,
                     save_root=str(tmp_path / "store"), model="vit_base_patch14_dinov2.lvd142m")
    img = os.path.join(data_root, "sub", "x.jpg")
    assert not misc.check_if_file_exists(args, img)
    feats = torch.arange(V * h * w * C, dtype=torch.float32).reshape(V, h, w, C)
    np.random.seed(0)
    M.denoise_an_image(args, Pipe(), feats, torch.zeros(V, h, w, 2), img_pth=img)
    raw_p, den_p = misc.feature_paths(args, img)
    # (the doubled slash is the reference's: data_root without a trailing slash is replaced by a directory with one)
    assert os.path.normpath(raw_p).endswith("store/raw_features/vit_base_patch14_dinov2.lvd142m/sub/x.npy")
    raw, den = np.load(raw_p), np.load(den_p)
    assert raw.dtype == np.float32 and raw.shape == (h, w, C) and np.array_equal(raw, feats[-1].numpy())
    assert den.dtype == np.float32 and den.shape == (1, h, w, C)
    with open(raw_p, "rb") as f:
        assert f.read(8) == bytes([0x93]) + b"NUMPY" + bytes([1, 0])          # NPY format version 1.0
    assert misc.check_if_file_exists(args, img)
