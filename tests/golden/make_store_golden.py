#!/usr/bin/env python
"""Golden values for the host-side contract around the hot paths (the `.npy` feature store; 8a-4: the
LR schedule), minted from the REFERENCE's own `dvt/utils/misc.py` (imported unmodified from /root/reference):
  * the paths `check_if_file_exists` probes for an image (recorded through a patched os.path.isfile),
  * `adjust_learning_rate` over whole schedules, incl. the CLI default warmup_iters (2500) > num_iters (2000).
Run in the build container only."""
import importlib.util
import json
import os
from argparse import Namespace
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_misc", "/root/reference/dvt/utils/misc.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = []
    for data_root, save_root, model, fn in [
            ("data/VOCdevkit", "work/feats", "vit_base_patch14_dinov2.lvd142m", "data/VOCdevkit/VOC2012/JPEGImages/2007_000027.jpg"),
            ("/abs/imgs/", "/abs/out", "vit_small_patch14_reg4_dinov2.lvd142m", "/abs/imgs/a/b/c.JPEG"),
            ("demo", "out", "vit_base_patch16_224.dino", "demo/cat.png")]:
        args = Namespace(data_root=data_root, save_root=save_root, model=model)
        seen = []
        with mock.patch("os.path.isfile", side_effect=lambda p: (seen.append(p), True)[1]):
            assert ref.check_if_file_exists(args, fn) is True
        cases.append({"data_root": data_root, "save_root": save_root, "model": model, "filename": fn, "raw": seen[0],
                      "denoised": seen[1]})
    sched = []
    for lr, min_lr, warm, n in [(0.01, 0.001, 200, 2000), (0.01, 0.001, 2500, 2000), (0.01, 0.001, 2500, 25000)]:
        a = Namespace(lr=lr, min_lr=min_lr, warmup_iters=warm, num_iters=n)

        class Opt:
            param_groups = [{"lr": None}, {"lr": None, "lr_scale": 0.5}]
        its = sorted(set(list(range(0, n, max(1, n // 40))) + [1, warm - 1, warm, warm + 1, n - 1]) & set(range(n)))
        vals = []
        for it in its:
            o = Opt()
            v = ref.adjust_learning_rate(o, it, a)
            assert o.param_groups[0]["lr"] == v and o.param_groups[1]["lr"] == v * 0.5
            vals.append(v)
        sched.append({"lr": lr, "min_lr": min_lr, "warmup_iters": warm, "num_iters": n, "iterations": its, "values": vals})
    # ---- stage 2: CosineScheduler (dvt/utils/misc.py:211-241) and the two samplers (dvt/dataset/sampler.py:7-45) ----
    import itertools
    stage2 = []
    for base, final, total in [(2e-4 * (32 * 8 / 256) ** 0.5, 1e-6, 400), (3e-3, 1e-5, 30), (1e-3, 0.0, 7)]:
        sc = ref.CosineScheduler(base_value=base, final_value=final, total_iters=total, warmup_iters=int(total * 0.15),
                                 start_warmup_value=0)
        stage2.append({"base_value": base, "final_value": final, "total_iters": total, "warmup_iters": int(total * 0.15),
                       "values": [float(sc[i]) for i in range(total + 3)]})
    spec2 = importlib.util.spec_from_file_location("ref_sampler", "/root/reference/dvt/dataset/sampler.py")
    smp = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(smp)
    samplers = {"infinite_n5_first12": list(itertools.islice(iter(smp.InfiniteSampler(range(5))), 12))}
    for world in (2, 3):
        for rank in range(world):
            d = smp.DistributedInfiniteSampler(range(11), num_replicas=world, rank=rank)
            samplers[f"distributed_n11_w{world}_r{rank}_first14"] = [int(i) for i in itertools.islice(iter(d), 14)]
            samplers[f"distributed_n11_w{world}_r{rank}_len"] = len(d)
    out = os.path.join(HERE, "store_and_schedule.json")
    json.dump({"paths": cases, "schedules": sched, "stage2_schedules": stage2, "samplers": samplers}, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
