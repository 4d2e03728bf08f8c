"""Stage-1 of DVT on one GPU: feature-bank extraction (HP-1) + per-image neural-field fit (HP-2).

This is the public call the stage-1 CLI (`main_img_denoising.py`) and `bench.py` make per image; it replaces the
body of the reference's per-image loop (main_img_denoising.py:301-343): 769 ViT forwards into a device-resident
bank, then `denoise_an_image`.  Everything that computes runs in libdvt_b200.so."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import models as DVT
from .fit import FitEngine, make_patch_coordinates


@dataclass
class Stage1Config:
    num_iters: int = 25000
    warmup_iters: int = 2500
    n_levels: int = 16
    freeze_shared_artifacts_after: float = 0.5
    lr: float = 0.01
    min_lr: float = 0.001
    weight_decay: float = 1e-5
    extract_bsz: int = 32
    pixel_bsz: int = 2048
    loss_scale: float = 1024.0      # torch.amp.GradScaler("cuda", 2**10), never unscaled (main_img_denoising.py:55,88)
    graph_steps: int = 20
    log_losses: bool = False        # keep the per-step loss table of every image (out["losses"], pinned host memory)
    fit_engines: int = 2            # fits kept in flight by run_images (one engine each): the step chain of a fit is a
                                    # sequence of small latency-bound kernels, two independent chains fill each other's gaps


class Stage1Pipeline:
    def __init__(self, vit: DVT.PretrainedViTWrapper, layer_index: int, input_size, cfg: Stage1Config, seed: int = 0):
        self.vit, self.layer_index, self.cfg = vit, layer_index, cfg
        H, W = input_size
        P, S = vit.patch_size, vit.stride
        self.h, self.w = (H - P) // S + 1, (W - P) // S + 1
        self.C = vit.n_output_dims
        self.input_size = (H, W)
        # module parameters live on the GPU: the per-image re-initialisation and the upload into the engine are then
        # device-side copies (the reference also builds its modules on the GPU, main_img_denoising.py:39-47)
        self.field = DVT.NeuralFeatureField(feat_dim=self.C, n_levels=cfg.n_levels).cuda()
        self.engines = [FitEngine(self.C, self.h, self.w, cfg.pixel_bsz, self.field.meta) for _ in range(max(1, cfg.fit_engines))]
        self.engine = self.engines[0]
        self.base_seed = int(seed)               # parameter initialisation of image k uses (base_seed, k)
        self._image_counter = 0
        # coordinates of the final query (main_img_denoising.py:121-130), uploaded once: a pageable host-to-device copy
        # after the fit would block the host until the fit has finished and serialise run_images
        self._full_coords = make_patch_coordinates(self.h, self.w, 0, 1).to("cuda")
        assert self._full_coords.min() >= 0 and self._full_coords.max() <= 1
        self._banks = [None] * (len(self.engines) + 1)   # feature banks: one per fit in flight + the one being extracted
        self._stage: Optional[torch.Tensor] = None
        self._extract_stream: Optional[torch.cuda.Stream] = None

    # ---- HP-1 ----------------------------------------------------------------------------------------------
    def extract_bank(self, views: torch.Tensor, slot: int = 0) -> torch.Tensor:
        """views [V, 3, H, W] (cuda, or pinned host memory) -> bank [V, h, w, C] fp32 (cuda), written into bank buffer
        `slot`.  Host views are copied batch by batch on a side stream into two staging buffers, so the copy of
        batch k+1 overlaps the forward of batch k."""
        V = views.shape[0]
        if self._banks[slot] is None or self._banks[slot].shape[0] != V:
            self._banks[slot] = torch.empty((V, self.h, self.w, self.C), device="cuda", dtype=torch.float32)
        bank = self._banks[slot]
        bsz = self.cfg.extract_bsz
        starts = list(range(0, V, bsz))
        on_host = not views.is_cuda
        if on_host:
            if self._stage is None or self._stage[0].shape[0] != bsz or self._stage[0].dtype != views.dtype:
                self._stage = [torch.empty((bsz,) + tuple(views.shape[1:]), device="cuda", dtype=views.dtype) for _ in range(2)]
                self._copy_stream = torch.cuda.Stream()
                self._copied = [torch.cuda.Event(), torch.cuda.Event()]
                self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
            main = torch.cuda.current_stream()

            def issue_copy(k):
                s0 = starts[k]
                n = min(bsz, V - s0)
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(self._consumed[k % 2])  # staging buffer free again
                    self._stage[k % 2][:n].copy_(views[s0:s0 + n], non_blocking=True)
                    self._copied[k % 2].record(self._copy_stream)

            for e in self._consumed:
                e.record(main)
            issue_copy(0)
        for k, s0 in enumerate(starts):
            n = min(bsz, V - s0)
            if on_host:
                if k + 1 < len(starts):
                    issue_copy(k + 1)
                main.wait_event(self._copied[k % 2])
                x = self._stage[k % 2][:n]
            else:
                x = views[s0:s0 + n]
            self.vit.extract_into(x, self.layer_index, bank[s0:s0 + n])  # NHWC, no NCHW round trip
            if on_host:
                self._consumed[k % 2].record(main)
        return bank

    # ---- HP-2 ----------------------------------------------------------------------------------------------
    def denoise(self, bank: torch.Tensor, coords: torch.Tensor, idx_stream: np.ndarray,
                init: Optional[Dict[str, torch.Tensor]] = None, seed: Optional[int] = None,
                validate: bool = True, engine: int = 0) -> Dict[str, torch.Tensor]:
        """bank [V, h, w, C] f32 cuda, coords [V, h, w, 2] in [0,1]; the last view is the un-augmented image.
        Returns denoised_feats [1, h, w, C] (= neural_field(coords[-1]), what the reference saves) and raw [h, w, C].
        Every image starts from fresh parameters, like the reference (new SingleImageDenoiser + NeuralFeatureField per
        image, main_img_denoising.py:39-47); they are drawn on the device.  Nothing here waits for the GPU when
        validate=False (run_images), so the next image can be enqueued while this fit is running."""
        cfg = self.cfg
        eng = self.engines[engine]
        if seed is None:
            self._image_counter += 1
            seed = (self.base_seed << 20) + self._image_counter
        eng.init_params(seed)
        if init is not None:
            for k, v in init.items():
                eng.set_param(k, v)
        V = bank.shape[0]
        eng.begin(bank.reshape(V * self.h * self.w, self.C), coords.reshape(-1, 2).to("cuda", torch.float32).contiguous(),
                          idx_stream, lr=cfg.lr, min_lr=cfg.min_lr, warmup_iters=cfg.warmup_iters,
                          freeze_after=cfg.freeze_shared_artifacts_after, weight_decay=cfg.weight_decay,
                          loss_scale=cfg.loss_scale, validate=validate)
        eng.run(graph_steps=cfg.graph_steps)
        denoised = eng.query(self._full_coords, assume_valid=True).reshape(1, self.h, self.w, self.C)
        extra = {}
        if cfg.log_losses:
            extra["losses"] = eng.losses_async()      # valid once `losses_ready` has completed
            extra["losses_ready"] = torch.cuda.Event()
            extra["losses_ready"].record()
        # (a copy: the bank buffer is recycled for the image after next before a pipelined caller reads its result)
        return {"denoised_feats": denoised, "raw": bank[-1].clone(), **extra}

    def export_modules(self):
        """The fitted parameters as the reference's module objects (`SingleImageDenoiser`, `NeuralFeatureField`), e.g. for
        `visualize_offline_denoised_samples`-style inspection.  Blocks; not used by the per-image loop."""
        with torch.device("cuda"):
            den = DVT.SingleImageDenoiser(self.h, self.w, self.C, layer_index=self.layer_index)
        self.engine.store_modules(den, self.field)
        return den, self.field

    # ---- both paths, software-pipelined over images -----------------------------------------------------------
    def run_images(self, n_images: int, views_fn, coords_fn, idx_fn, finalize, events: Optional[list] = None,
                   overlap: bool = True):
        """Processes images 0 .. n_images-1.  The bank extraction of image i+1 (HP-1: tensor-core bound, on a low-priority
        stream) runs beside the fit of image i (HP-2: latency / HBM bound, on the engine's high-priority streams); the fit
        of an image needs its complete bank, so this is the only overlap the data dependencies allow.
          views_fn(i)  -> [V, 3, H, W] cuda or pinned-host views       coords_fn(i) -> [V, h, w, 2] global coordinates
          idx_fn(i)    -> int [num_iters, pixel_bsz] sampled bank rows  finalize(i, out) -> result (may block, e.g. D2H)
        events: optional list receiving ("hp1" | "hp2", start_event, end_event) per image (timing events recorded on the
        stream that runs the path).  overlap=False: strictly one image after the other (for A/B measurements)."""
        if n_images <= 0:
            return []
        ne = len(self.engines) if overlap else 1          # fits in flight
        nslots = len(self._banks)
        if self._extract_stream is None:
            self._extract_stream = torch.cuda.Stream(priority=0)   # lowest priority: the fit's short kernels go first
            self._fit_streams = [torch.cuda.Stream(priority=-1) for _ in self.engines]
            self._ext_done = [torch.cuda.Event() for _ in range(nslots)]
            self._fit_done = [torch.cuda.Event() for _ in range(nslots)]
        main = torch.cuda.current_stream()
        sx = self._extract_stream if overlap else main
        tev = (lambda: torch.cuda.Event(enable_timing=True)) if events is not None else None

        def make_views(i):
            # views_fn may launch GPU work (GpuViewGenerator): it goes to the extraction stream, in front of the forwards
            with torch.cuda.stream(sx):
                return views_fn(i)

        def enqueue_extract(i, views):
            slot = i % nslots
            sx.wait_event(self._fit_done[slot])    # the fit that read this bank buffer `nslots` images ago
            with torch.cuda.stream(sx):
                if tev:
                    a = tev()
                    a.record(sx)
                bank = self.extract_bank(views, slot=slot)
                if tev:
                    b = tev()
                    b.record(sx)
                    events.append(("hp1", a, b))
                self._ext_done[slot].record(sx)
            return bank

        for e in self._fit_done:
            e.record(main)
        sx.wait_stream(main)                       # inputs produced on the caller's stream
        for fs in self._fit_streams:
            fs.wait_stream(main)
        results = []
        bank = enqueue_extract(0, make_views(0))
        idx = idx_fn(0)
        pending = []                               # (i, out) of fits in flight: finalised `ne` images late
        for i in range(n_images):
            # The views of the NEXT image are produced now, i.e. behind the forwards of image i on the extraction stream.
            views_next = make_views(i + 1) if i + 1 < n_images else None
            fs = self._fit_streams[i % ne] if overlap else main
            fs.wait_event(self._ext_done[i % nslots])
            with torch.cuda.stream(fs):
                if tev:
                    a = tev()
                    a.record(fs)
                # enqueue only: the host does not wait for the fit
                out = self.denoise(bank, coords_fn(i), idx, validate=False, engine=i % ne)
                if tev:
                    b = tev()
                    b.record(fs)
                    events.append(("hp2", a, b))
                self._fit_done[i % nslots].record(fs)
                done = torch.cuda.Event()
                done.record(fs)
            out["_done"] = done
            if i + 1 < n_images:                   # enqueued while the GPU runs the fits of the previous images
                bank = enqueue_extract(i + 1, views_next)
                idx = idx_fn(i + 1)
            # finalize() may block (device-to-host reads): results are collected only after `ne` later fits have been
            # enqueued completely, so the GPU always has work queued behind the running fits
            pending.append((i, out))
            if len(pending) > ne:
                j, o = pending.pop(0)
                main.wait_event(o.pop("_done"))
                results.append(finalize(j, o))
        for j, o in pending:
            main.wait_event(o.pop("_done"))
            results.append(finalize(j, o))
        main.wait_stream(sx)
        for fs in self._fit_streams:
            main.wait_stream(fs)
        for eng in self.engines:
            eng.check()       # input validation of all fits of this call (blocks: the results are about to be read anyway)
        return results
