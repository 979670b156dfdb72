"""N>1 path on CPU: world_size-2 gloo, each rank "samples" its shard of the images through the real
host code (engine / fused-free generic sampler) with shape-only kernels whose output encodes the
image identity, then the final all_gather must return every image exactly once, in global order —
i.e. the N-rank result equals the 1-rank result on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from custom_diffusion360_b200.parallel import gather_images, image_shard
        mine = image_shard(n_images, rank, world)
        # stand-in for the per-image trajectories: latent i is filled with i + 0.5
        x_local = torch.stack([torch.full((4, 8, 8), i + 0.5) for i in mine]) if mine else torch.zeros(0, 4, 8, 8)
        full = gather_images(x_local, n_images)
        q.put((rank, mine, full[:, 0, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5, 1])
def test_image_parallel_gather_two_ranks(n_images):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [i + 0.5 for i in range(n_images)]
    shards = []
    for rank, mine, full in results:
        assert full == expect  # every rank holds all images in global order
        shards += mine
    assert sorted(shards) == list(range(n_images))  # each image sampled by exactly one rank


def test_shard_balance():
    from custom_diffusion360_b200.parallel import image_shard
    for n in (1, 7, 8, 32, 144):
        for w in (1, 2, 4, 8):
            sizes = [len(image_shard(n, r, w)) for r in range(w)]
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _train_worker(rank, world, port, q):
    """Data-parallel training exchange on CPU/gloo: flat parameter / gradient buffers, one bucket
    per pose block, per-block all-reduce started by the backward walk's callbacks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sgm_oracle as O
        from custom_diffusion360_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
        from custom_diffusion360_b200.sgm.optim import PoseAdamW
        torch.manual_seed(0)
        unet = UNetModel(**dict(O.TINY_CFG))
        named = [(n, p) for n, p in unet.named_parameters() if "pose" in n]
        opt = PoseAdamW(named, lr=1e-3)
        opt.attach_overlap(unet)
        blocks = [m for _, m in unet.pose_blocks()]
        assert len(opt.buckets) == len(blocks)
        # rank r's "gradient" of parameter k is (r + 1) * (k + 1); blocks finish in reverse order
        for k, (_, p) in enumerate(named):
            p.grad.fill_(float((rank + 1) * (k + 1)))
        for blk in reversed(blocks):
            blk.__dict__["_grads_ready"]()
        opt.wait_reduce()
        tot = sum(r + 1 for r in range(world))
        ok = all(bool((p.grad == float(tot * (k + 1))).all()) for k, (_, p) in enumerate(named))
        # graph mode (GraphedTrainStep): the callbacks are suspended, step() reduces everything afterwards
        opt.suspend_overlap = True
        for k, (_, p) in enumerate(named):
            p.grad.fill_(float((rank + 1) * (k + 1)))
        for blk in reversed(blocks):
            blk.__dict__["_grads_ready"]()
        assert not opt._pending
        ok = ok and all(bool((p.grad == float((rank + 1) * (k + 1))).all()) for k, (_, p) in enumerate(named))
        opt.reduce_all()
        opt.wait_reduce()
        ok = ok and all(bool((p.grad == float(tot * (k + 1))).all()) for k, (_, p) in enumerate(named))
        # views survive: parameters and grads still alias the flat buffers
        alias = all(p.data_ptr() == opt.flat.data.data_ptr() + 4 * o for p, o in zip(opt.flat.params, opt.flat.offsets))
        q.put((rank, ok, alias, len(opt.buckets), opt.flat.numel))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_buckets_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, alias, nb, numel in results:
        assert ok and alias and nb == 9 and numel > 0


def _refs_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sgm_oracle as O
        from custom_diffusion360_b200.parallel import gather_references
        from custom_diffusion360_b200.sgm.modules.diffusionmodules.openaimodel import UNetModel
        unet = UNetModel(**dict(O.TINY_CFG))
        caps, null = {}, {}
        for name, m in unet.pose_blocks():
            c = m.pose_emb_layers.weight.shape[0]
            # view j of rank k is filled with 10 * j + k
            caps[name] = torch.stack([torch.full((4, c), 10.0 * j + rank) for j in range(3)])
            null[name] = torch.full((4, c), -1.0)
        refs = gather_references(unet, caps, null_row=null)
        ok = True
        for name, m in unet.pose_blocks():
            r = m.references
            expect = [10.0 * j + k for j in range(3) for k in range(world)] + [-1.0]   # main.py:601 ordering
            ok = ok and r.shape[0] == 3 * world + 1 and r[:, 0, 0].tolist() == expect and r is refs[name]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_validation_reference_gather_two_ranks():
    """per-pose-block all_gather of the captured reference tokens, interleaved like
    main.py:600-601, registered as `references`."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_refs_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results)
