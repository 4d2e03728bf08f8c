"""DVT stage 1 (per-image denoising) on B200 -- drop-in for the reference's main_img_denoising.py.

Same command line (reference main_img_denoising.py:152-217), same outputs
(`{save_root}/raw_features/{model}/<rel>.npy` (h, w, C) float32 and `{save_root}/denoised_features/{model}/<rel>.npy`
(1, h, w, C) float32, :131-146), same skip/resume rule (:303-307).  The two hot paths run in libdvt_b200.so:
feature-bank extraction (769 ViT forwards) and the neural-field fit (`denoise_an_image`).

Differences that are deliberate and documented in DESIGN.md: the PCA visualisation of every `vis_freq`-th image
(reference :101-117) is not produced (matplotlib / torch_kmeans are outside the hot path); `--dtype` selects only the
storage type of the host views (the kernels always compute bf16 x bf16 -> fp32).
"""
import argparse
import datetime
import glob
import json
import os
import sys
import time

import numpy as np
import torch
from torchvision import transforms

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "denoising-vit_b200"))

import dvt.models as DVT  # noqa: E402
import dvt.utils.misc as misc  # noqa: E402
from dvt.dataset import GpuViewGenerator, RandomResizedCropFlip, SingleImageDataset  # noqa: E402
from dvt.fit import make_patch_coordinates  # noqa: E402
from dvt.stage1 import Stage1Config, Stage1Pipeline  # noqa: E402


def denoise_an_image(args, pipeline: Stage1Pipeline, all_raw_features, all_pixel_coords, img_pth=None):
    """Counterpart of the reference function (main_img_denoising.py:28-149): fit, final query, save."""
    n_rows = all_raw_features.shape[0] * all_raw_features.shape[1] * all_raw_features.shape[2]
    # the reference draws np.random.randint(0, n_rows, pixel_bsz) once per step from the global legacy RNG (:73);
    # drawing all steps at once consumes the identical MT19937 stream
    idx_stream = np.random.randint(0, n_rows, (args.num_iters, args.pixel_bsz))
    out = pipeline.denoise(all_raw_features, all_pixel_coords, idx_stream)
    losses = pipeline.engine.losses()
    for step in sorted(set(list(range(0, args.num_iters, 1000)) + [args.num_iters - 1])):
        lr = misc.learning_rate_at(step, args.lr, args.min_lr, args.warmup_iters, args.num_iters)
        l = losses[step]
        print(f"Step {step}/{args.num_iters - 1}: Loss = {l[0]:.4f}, Patch Loss = {l[1]:.4f}, CosSim Loss = {l[2]:.4f}, "
              f"Residual Loss = {l[3]:.4f}, Residual Sparsity Loss = {l[4]:.4f}, LR = {lr:.4f}")
    if args.data_root is not None:
        raw_path, den_path = misc.feature_paths(args, img_pth)
        os.makedirs(os.path.dirname(raw_path), exist_ok=True)
        os.makedirs(os.path.dirname(den_path), exist_ok=True)
        np.save(raw_path, out["raw"].float().cpu().numpy())
        np.save(den_path, out["denoised_feats"].float().cpu().numpy())
        print(f"Saved denoised features to {den_path} and raw features to {raw_path}")
    return out


def get_args(argv=None):
    p = argparse.ArgumentParser(description="DVT Stage-1: Single Image Denoising")
    p.add_argument("--model", type=str, default="vit_base_patch14_dinov2.lvd142m", choices=DVT.MODEL_LIST)
    p.add_argument("--input_size", type=int, default=518, nargs="+")
    p.add_argument("--stride_size", type=int, default=14)
    p.add_argument("--layer_depth_ratio", type=float, default=1.0)
    p.add_argument("--img_path", type=str, default="demo/assets/demo/cat.jpg")
    p.add_argument("--dtype", type=str, default="float32")
    p.add_argument("--data_root", type=str, default=None)
    p.add_argument("--save_root", type=str, default=None)
    p.add_argument("--start_idx", type=int, default=0)
    p.add_argument("--num_imgs", type=int, default=100)
    p.add_argument("--num_views", type=int, default=768)
    p.add_argument("--num_iters", type=int, default=25000)
    p.add_argument("--warmup_iters", type=int, default=2500)
    p.add_argument("--n_levels", type=int, default=16)
    p.add_argument("--freeze_shared_artifacts_after", type=float, default=0.5)
    p.add_argument("--lr", type=float, default=0.01)
    p.add_argument("--min_lr", type=float, default=0.001)
    p.add_argument("--weight_decay", type=float, default=1e-5)
    p.add_argument("--extract_bsz", type=int, default=32)
    p.add_argument("--pixel_bsz", type=int, default=2048)
    p.add_argument("--output_dir", type=str, default="./work_dirs/demo")
    p.add_argument("--num_vis_samples", type=int, default=5)
    p.add_argument("--vis_freq", type=int, default=100)
    p.add_argument("--seed", type=int, default=0)
    # B200 extension (not a reference flag): where the augmented views are produced
    p.add_argument("--view_backend", type=str, default="gpu", choices=["gpu", "cpu"],
                   help="gpu: one dvt_view_crops launch per image; cpu: the reference's DataLoader of CPU transforms")
    args = p.parse_args(argv)
    assert os.path.exists(args.img_path), f"Image not found: {args.img_path}"
    if isinstance(args.input_size, int):
        args.input_size = (args.input_size, args.input_size)
    elif len(args.input_size) == 1:
        args.input_size = (args.input_size[0], args.input_size[0])
    args.input_size = tuple(args.input_size)
    assert args.input_size[0] % args.stride_size == 0, "height must be divisible by stride_size"
    assert args.input_size[1] % args.stride_size == 0, "width must be divisible by stride_size"
    return args


def main(args):
    os.makedirs(args.output_dir, exist_ok=True)
    misc.fix_random_seeds(args.seed)
    print(f"Arguments:\n{json.dumps(vars(args), indent=4)}")
    assert torch.cuda.is_available(), "the B200 stage-1 driver needs a CUDA device (no CPU fallback)"
    device = "cuda"
    if os.path.isfile(args.img_path):
        if args.img_path.endswith("txt"):
            with open(args.img_path) as f:
                filenames = f.read().splitlines()
        else:
            filenames = [args.img_path]
    else:
        filenames = glob.glob(os.path.join(args.img_path, "**/*"), recursive=True)
    filenames = filenames[args.start_idx:args.start_idx + args.num_imgs]

    vit = DVT.PretrainedViTWrapper(model_identifier=args.model, stride=args.stride_size).to(device).eval()
    layer_index = int(args.layer_depth_ratio * vit.last_layer_index)
    args.layer_index, args.feat_dim = layer_index, vit.n_output_dims
    normalizer = vit.transformation.transforms[-1]
    assert isinstance(normalizer, transforms.Normalize), "last transform must be norm"
    host_dtype = torch.float32 if args.dtype == "float32" else torch.bfloat16
    cfg = Stage1Config(num_iters=args.num_iters, warmup_iters=args.warmup_iters, n_levels=args.n_levels,
                       freeze_shared_artifacts_after=args.freeze_shared_artifacts_after, lr=args.lr, min_lr=args.min_lr,
                       weight_decay=args.weight_decay, extract_bsz=args.extract_bsz, pixel_bsz=args.pixel_bsz)
    pipe = Stage1Pipeline(vit, layer_index, args.input_size, cfg)
    args.noise_map_height, args.noise_map_width = pipe.h, pipe.w

    num_samples = args.num_views + 1  # + the un-augmented image
    coords = torch.zeros((num_samples, pipe.h, pipe.w, 2), dtype=torch.float32, device=device)
    if args.view_backend == "gpu":
        views, views_dev = None, torch.empty((num_samples, 3) + args.input_size, dtype=host_dtype, device=device)
        view_gen = GpuViewGenerator(args.input_size, num_views=args.num_views, scale=(0.1, 0.5), patch_size=vit.patch_size,
                                    stride=args.stride_size, dtype=host_dtype)
    else:
        views = torch.zeros((num_samples, 3) + args.input_size, dtype=host_dtype).pin_memory()
    dataset = SingleImageDataset(
        size=args.input_size,
        base_transform=transforms.Compose([transforms.ToPILImage(), transforms.Resize(args.input_size),
                                           transforms.ToTensor(), normalizer]),
        final_transform=RandomResizedCropFlip(size=args.input_size, horizontal_flip=True, scale=(0.1, 0.5),
                                              patch_size=vit.patch_size, stride=args.stride_size),
        num_views=args.num_views)

    done, start = 0, time.time()
    for idx, filename in enumerate(filenames):
        filename = filename.strip().split(" ")[0]
        if args.data_root is not None:
            filename = os.path.join(args.data_root, filename)
            if misc.check_if_file_exists(args, filename):
                print(f"Skipping {filename}")
                continue
        dataset.set_image(filename)
        t0 = time.time()
        if args.view_backend == "gpu":
            # one kernel launch for all views (dvt_view_crops); the host only draws the crop boxes / flips, with the
            # reference's own RNG calls (dvt/dataset/gpu_views.py)
            gpu_views, gpu_coords = view_gen(dataset.image.to(device, torch.float32), views_out=views_dev, coords_out=coords)
            bank = pipe.extract_bank(gpu_views)
        else:
            loader = torch.utils.data.DataLoader(dataset, args.extract_bsz, num_workers=8)
            for i, data in enumerate(loader):
                s = slice(i * args.extract_bsz, i * args.extract_bsz + data["transformed_view"].shape[0])
                views[s] = data["transformed_view"].to(host_dtype)
                coords[s] = data["pixel_coords"].to(device)
            views[-1] = data["full_image"][0].to(host_dtype)
            coords[-1] = make_patch_coordinates(pipe.h, pipe.w, start=0, end=1)
            bank = pipe.extract_bank(views)
        torch.cuda.synchronize()
        t1 = time.time()
        print(f"Feature extraction time: {t1 - t0:.2f}s")
        denoise_an_image(args, pipe, bank, coords, img_pth=filename)
        torch.cuda.synchronize()
        t2 = time.time()
        done += 1
        print(f"Denoising time: {t2 - t1:.2f}s")
        elapsed = time.time() - start
        eta = elapsed / done * (len(filenames) - done)
        print(f"[{idx + 1}/{len(filenames)}] ETA: {datetime.timedelta(seconds=int(eta))}, "
              f"Elapsed: {datetime.timedelta(seconds=int(elapsed))}")
        print("-" * 80)
    print(f"Total time: {datetime.timedelta(seconds=int(time.time() - start))}")


if __name__ == "__main__":
    main(get_args())
