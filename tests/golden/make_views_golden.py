#!/usr/bin/env python
"""Golden vectors for the view generation: runs the REFERENCE's own
`dvt/dataset/transform.py::RandomResizedCropFlip` (imported unmodified from /root/reference) with seeded torch / numpy
RNGs on a small synthetic image and stores what it returned, plus the parameters it drew (re-derived from the same
seeds with the oracle's sampler, asserted equal through the outputs).  Run in the build container only."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import views as OV  # noqa: E402


def load_reference_transform():
    spec = importlib.util.spec_from_file_location("ref_transform", "/root/reference/dvt/dataset/transform.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_transform()
    H, W, size, P, S, V, seed = 64, 80, (56, 70), 14, 14, 6, 7
    img = torch.randn(3, H, W, generator=torch.Generator().manual_seed(seed))
    tf = ref.RandomResizedCropFlip(size=size, horizontal_flip=True, scale=(0.1, 0.5), patch_size=P, stride=S)
    torch.manual_seed(seed)
    np.random.seed(seed)
    views, coords = zip(*[tf(img) for _ in range(V)])
    views, coords = torch.stack(views), torch.stack(coords)
    # the same RNG stream through the oracle's sampler + restatement must reproduce the reference bit for bit
    torch.manual_seed(seed)
    np.random.seed(seed)
    boxes, flips = OV.sample_view_params(img, V)
    v2, c2 = OV.make_views(img, boxes, flips, size, P, S)
    assert torch.equal(views, v2) and torch.equal(coords, c2), "oracle/views.py does not reproduce the reference"
    assert flips.sum() not in (0, V), "seed gives no mix of flipped / unflipped views"
    # one down-scaling crop (scale > 1: widened anti-aliasing support) and the full image, for the kernel tests
    extra_boxes = np.array([[0, 0, H, W], [3, 5, 60, 72]], dtype=np.int32)
    extra_flips = np.array([0, 1], dtype=np.int32)
    ev, ec = OV.make_views(img, extra_boxes, extra_flips, (28, 42), P, S)
    out = os.path.join(HERE, "views_small.npz")
    np.savez_compressed(out, image=img.numpy(), size=np.array(size), patch=np.array([P, S]), seed=np.array([seed]),
                        boxes=boxes, flips=flips, views=views.numpy(), coords=coords.numpy(),
                        extra_boxes=extra_boxes, extra_flips=extra_flips, extra_size=np.array([28, 42]),
                        extra_views=ev.numpy(), extra_coords=ec.numpy())
    print("wrote", out, os.path.getsize(out), "bytes; boxes", boxes.tolist(), "flips", flips.tolist())


if __name__ == "__main__":
    main()
