"""CPU: the view-generation oracle against the golden vectors minted from the reference's own transform
(tests/golden/make_views_golden.py), and the host-side sampler of the product against the oracle's."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "views_small.npz")


def test_oracle_reproduces_reference_views():
    from oracle import views as OV
    z = np.load(GOLD)
    img = torch.from_numpy(z["image"])
    P, S = (int(v) for v in z["patch"])
    views, coords = OV.make_views(img, z["boxes"], z["flips"], tuple(int(v) for v in z["size"]), P, S)
    assert torch.equal(views, torch.from_numpy(z["views"]))          # bit for bit: same library calls as the reference
    assert torch.equal(coords, torch.from_numpy(z["coords"]))
    ev, ec = OV.make_views(img, z["extra_boxes"], z["extra_flips"], tuple(int(v) for v in z["extra_size"]), P, S)
    assert torch.equal(ev, torch.from_numpy(z["extra_views"])) and torch.equal(ec, torch.from_numpy(z["extra_coords"]))


def test_sampler_replays_the_reference_rng_stream():
    """Same torch / numpy RNG calls, in the same order, as transform.py:48,69: the drawn boxes and flips equal the ones the
    reference drew when the goldens were minted -- for the oracle's sampler and for the product's (dvt.dataset.gpu_views)."""
    from dvt.dataset import gpu_views as GV
    from oracle import views as OV
    z = np.load(GOLD)
    img = torch.from_numpy(z["image"])
    seed = int(z["seed"][0])
    for sampler in (OV.sample_view_params, GV.sample_view_params):
        torch.manual_seed(seed)
        np.random.seed(seed)
        boxes, flips = sampler(img, len(z["flips"]))
        assert np.array_equal(boxes, z["boxes"]) and np.array_equal(flips, z["flips"]), sampler.__module__
    coords = z["coords"]
    assert coords.min() >= 0.0 and coords.max() <= 1.0
