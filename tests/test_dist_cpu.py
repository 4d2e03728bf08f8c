"""N>1 path on CPU: world_size-2 gloo run of the image sharding + the single all-gather collate."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "denoising-vit_b200"))
    from dvt.dist import collate_maps, shard_indices
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(num_items, rank, world)
    local = torch.stack([torch.full((2, 3, 4), float(i)) for i in mine]) if mine else torch.zeros(0, 2, 3, 4)
    full = collate_maps(local, num_items)
    ok = full.shape == (num_items, 2, 3, 4) and all(float(full[i, 0, 0, 0]) == i for i in range(num_items))
    q.put((rank, ok, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_collate_world2():
    ctx = mp.get_context("spawn")
    for num_items in (7, 8):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, num_items, q)) for r in range(2)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in procs]
        [p.join(timeout=60) for p in procs]
        assert all(ok for _, ok, _ in res)
        covered = sorted(i for _, _, m in res for i in m)
        assert covered == list(range(num_items))  # every image on exactly one rank


def test_shard_indices_partition():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "denoising-vit_b200"))
    from dvt.dist import shard_indices
    for n in (0, 1, 5, 1000):
        for world in (1, 2, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
