"""World-size-2 tests of the samples-per-pixel sharding on CPU (gloo).

(1) the batch form: each rank renders its share of the iteration indices, one reduce at the end;
(2) the per-frame form of mygpuraytracer_b200/distributed.py: one frame = one iteration per rank, one reduce to
    rank 0 per frame, rank 0 hands the running sum to the host after EVERY frame (apps/src/pathtrace.cu:662-668).

Each rank renders its share of the iteration indices -- with the CPU oracle
standing in for the per-rank renderer, since there is no GPU here -- into its
own accumulator, then one reduce(sum) combines them, exactly the plumbing
bench.py uses with NCCL.  Checks: the two ranks cover every iteration index
exactly once, and the reduced image equals the sequential image within float
summation order."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(rank: int, world: int, total: int):
    """(first, count, stride) of rank: iteration indices rank+1, rank+1+world, ..."""
    count = (total - rank + world - 1) // world if total > rank else 0
    return rank + 1, count, world


def test_shards_partition_the_iteration_indices():
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 7, 8, 64, 100):
            seen = []
            for r in range(world):
                first, count, stride = shard(r, world, total)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(1, total + 1)), (world, total)


def _worker(rank, world, port, total, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from mygpuraytracer_b200 import abi
    from oracle import oracle
    from util import load_golden

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle.set_num_threads(2)
    scene, *_ = load_golden("cornellGlass_32x32")
    first, count, stride = shard(rank, world, total)
    img, alb, _nl, _seg = oracle.render(scene, abi.default_options(), first, count, stride)
    acc = torch.from_numpy(img)
    dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    # the albedo AOV comes from the rank that owns iteration 1
    a = torch.from_numpy(alb)
    dist.broadcast(a, src=0)
    if rank == 0:
        np.save(out, np.stack([acc.numpy(), a.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_matches_sequential_render(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    from mygpuraytracer_b200 import abi
    from oracle import oracle
    from util import load_golden

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    total = 9  # odd on purpose: rank 0 renders 5 iterations, rank 1 renders 4
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = np.load(out)
    scene, *_ = load_golden("cornellGlass_32x32")
    ref, ref_alb, *_ = oracle.render(scene, abi.default_options(), 1, total, 1)
    assert np.allclose(got[0], ref, rtol=1e-5, atol=1e-6)
    assert np.array_equal(got[1], ref_alb)
    # and the shards really differ from the whole (the reduce did something)
    half, *_ = oracle.render(scene, abi.default_options(), 1, 5, 2)
    assert not np.allclose(half, ref, rtol=1e-3)


class OracleShard:
    """The B2ptShard interface (api.Shard) with the CPU oracle as the per-rank renderer: lets the host-side frame
    protocol of distributed.FrameRenderer run where there is no GPU."""

    def __init__(self, pod, rank, world):
        import torch

        self.pod, self.rank, self.world = pod, rank, world
        self.n_pixels = pod.n_pixels
        self.sum = np.zeros((self.n_pixels, 3), np.float32)
        self.albedo = np.zeros((self.n_pixels, 3), np.float32)
        self.contribution = torch.zeros(self.n_pixels * 3)
        self.frames = []

    def frame_begin(self, first):
        from mygpuraytracer_b200 import abi
        from oracle import oracle

        it = first + self.rank
        self.frames.append(it)
        img, alb, _nl, _seg = oracle.render(self.pod, abi.default_options(), it, 1, 1)
        self.contribution.copy_(__import__("torch").from_numpy(img.reshape(-1)))
        if it == 1:
            self.albedo[:] = alb

    def frame_image_tensor(self):
        return self.contribution

    def frame_merge(self):
        if self.rank == 0:
            self.sum += self.contribution.numpy().reshape(-1, 3)
        self.contribution.zero_()

    def frame_end(self, image, albedo):
        if image is not None:
            image[:] = self.sum
        if albedo is not None:
            albedo[:] = self.albedo

    def sync(self):
        pass

    def close(self):
        pass


def _frame_worker(rank, world, port, frames, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from mygpuraytracer_b200 import distributed
    from oracle import oracle
    from util import load_golden

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle.set_num_threads(2)
    scene, *_ = load_golden("cornellGlass_32x32")
    shard = OracleShard(scene, rank, world)
    fr = distributed.FrameRenderer(scene, None, reduce="nccl", shard=shard)
    n = scene.n_pixels
    img, alb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    got = []
    for f in range(frames):
        fr.pathtrace(1 + f * world, img, alb)
        got.append(img.copy())
    assert shard.frames == [1 + f * world + rank for f in range(frames)]
    if rank == 0:
        np.save(out, np.stack(got + [alb]))
    else:
        assert not img.any(), "only rank 0 is handed pixels"
    dist.barrier()
    dist.destroy_process_group()


def test_per_frame_reduce_two_ranks(tmp_path):
    """One reduce per frame to rank 0; the host image after EVERY frame equals the sequential render of the same
    iterations (float summation order), the albedo comes from the rank that owns iteration 1."""
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    from mygpuraytracer_b200 import abi
    from oracle import oracle
    from util import load_golden

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    frames, world = 4, 2
    out = str(tmp_path / "frames.npy")
    mp.spawn(_frame_worker, args=(world, port, frames, out), nprocs=world, join=True)
    got = np.load(out)
    scene, *_ = load_golden("cornellGlass_32x32")
    for f in range(frames):
        ref, ref_alb, *_ = oracle.render(scene, abi.default_options(), 1, (f + 1) * world, 1)
        assert np.allclose(got[f], ref, rtol=1e-5, atol=1e-6), f"frame {f}"
    assert np.array_equal(got[frames], ref_alb)
