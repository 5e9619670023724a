"""pathtrace() over several GPUs (csrc/multi.cu): one frame = one iteration per member, combined by k_frame_reduce.

The contract is the reference's accumulator (apps/src/pathtrace.cu:508, 662-668): after a frame the host image is
the running sum including every iteration of the frame.  k_frame_reduce adds the members' contributions in
iteration order, so every case compares BIT FOR BIT with b2pt_pathtrace on one context rendering the same
iterations one after the other.  Members may share a device, so the whole protocol (peer pointers, slices,
rendezvous, re-arming) is exercised on a one-GPU box; with more GPUs the same tests also run across devices.
"""
import os
import socket
import sys

import numpy as np
import pytest

from mygpuraytracer_b200 import abi, api, assets, scenes
from util import assert_same_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mesh_scene(tmp_path, name, w, h, tris):
    root = assets.prepare(str(tmp_path / "run"), triangles=tris, procedural_size=256)
    return api.Scene(assets.scene_file(name, w, h, root=root)).pod


def _reference_sums(pod, n_iters):
    """Running sum after every iteration on one context, and the albedo AOV."""
    n = pod.n_pixels
    img, alb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    sums = []
    with api.Renderer(pod, abi.default_options()) as one:
        for it in range(1, n_iters + 1):
            one.pathtrace(it, img, alb)
            sums.append(img.copy())
    return sums, alb


def _device_lists(members):
    n = api.device_count()
    lists = [[0] * members]
    if n >= 2:
        lists.append([g % n for g in range(members)])
    return lists


@pytest.mark.parametrize("members,lanes", [(1, 3), (2, 1), (2, 4), (3, 2), (4, 2), (8, 1)])
def test_multi_frames_bitexact(tmp_path, members, lanes):
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 128, 72, 5000)
    frames = 5
    sums, ref_alb = _reference_sums(pod, members * frames)
    n = pod.n_pixels
    for devices in _device_lists(members):
        img, alb = np.full((n, 3), -1, np.float32), np.full((n, 3), -1, np.float32)
        with api.MultiRenderer(pod, abi.default_options(), devices=devices, lanes=lanes) as multi:
            for f in range(frames):
                multi.pathtrace(1 + f * members, img, alb)
                assert_same_bits(sums[(f + 1) * members - 1], img, f"running sum after frame {f} ({members} members on {devices})")
                assert_same_bits(ref_alb, alb, f"albedo after frame {f}")
            assert multi.launch_count() > 0


def test_multi_odd_size_reset_and_restart(tmp_path):
    # 131*77*3 floats: slices that are not all the same length and a scalar tail in the last one
    pod = api.Scene(scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=131, height=77)).pod
    n = pod.n_pixels
    sums, ref_alb = _reference_sums(pod, 12)
    img, alb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    with api.MultiRenderer(pod, abi.default_options(), devices=[0, 0, 0], lanes=2) as multi:
        for f in range(3):
            multi.pathtrace(1 + 3 * f, img, alb)
            assert_same_bits(sums[3 * f + 2], img, f"frame {f}")
        multi.reset()
        for f in range(4):  # Free + Init: the same frames again from zero
            multi.pathtrace(1 + 3 * f, img, alb)
            assert_same_bits(sums[3 * f + 2], img, f"after reset, frame {f}")
            assert_same_bits(ref_alb, alb, "albedo after reset")
        # a call that does not continue the sequence: speculated frames are dropped, the sum keeps every call
        multi.pathtrace(1, img, alb)
        with api.Renderer(pod, abi.default_options()) as one:
            ref = np.zeros((n, 3), np.float32)
            for it in list(range(1, 13)) + [1, 2, 3]:
                one.pathtrace(it, ref, None)
        assert_same_bits(ref, img, "after a restart at iteration 1")


def test_multi_device_resident_frames(tmp_path):
    """No host pointers: frames stay on the device, the sum is read once at the end."""
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 1000)
    n = pod.n_pixels
    sums, _ = _reference_sums(pod, 8)
    img = np.zeros((n, 3), np.float32)
    with api.MultiRenderer(pod, abi.default_options(), devices=[0, 0], lanes=3) as multi:
        for f in range(3):
            multi.pathtrace(1 + 2 * f, None, None)
        multi.pathtrace(7, img, None)
    assert_same_bits(sums[7], img, "sum after four frames, three of them device resident")


def test_multi_rejects_bad_arguments(tmp_path):
    pod = api.Scene(scenes.write_scene("cornell", str(tmp_path / "s.txt"), width=32, height=32)).pod
    with pytest.raises(api.B2ptError):
        api.MultiRenderer(pod, abi.default_options(), devices=[0, 99], lanes=2)
    with pytest.raises(api.B2ptError):
        api.MultiRenderer(pod, abi.default_options(), devices=[0] * 17, lanes=1)
    with pytest.raises(api.B2ptError):
        api.Shard(pod, abi.default_options(), rank=2, world=2, lanes=1)
    with api.Shard(pod, abi.default_options(), rank=0, world=2, lanes=1) as s:
        with pytest.raises(api.B2ptError):
            s.frame_reduce()          # no frame begun
        s.frame_begin(1)
        with pytest.raises(api.B2ptError):
            s.frame_reduce()          # world 2 but never connected


def test_shard_world1_equals_pipe(tmp_path):
    pod = _mesh_scene(tmp_path, "cornellSpaceship", 96, 54, 1000)
    n = pod.n_pixels
    sums, ref_alb = _reference_sums(pod, 6)
    img, alb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    with api.Shard(pod, abi.default_options(), rank=0, world=1, lanes=4) as s:
        for it in range(1, 7):
            s.frame_begin(it)
            s.frame_reduce()
            s.frame_end(img, alb)
            assert_same_bits(sums[it - 1], img, f"iteration {it}")
        assert_same_bits(ref_alb, alb, "albedo")
        assert s.misses() == 0


def _rank_worker(rank, world, port, scene_txt, reduce, frames, lanes, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from mygpuraytracer_b200 import abi, api, distributed

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    use_nccl = ndev >= world
    dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", dev)} if use_nccl else {}))
    pod = api.Scene(scene_txt).pod
    n = pod.n_pixels
    img, alb = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    got = []
    with distributed.FrameRenderer(pod, abi.default_options(device=dev), lanes=lanes, reduce=reduce) as fr:
        for f in range(frames):
            fr.pathtrace(1 + f * world, img, alb)
            if rank == 0:
                got.append(img.copy())
        fr.sync()
        dist.barrier()
    if rank == 0:
        np.save(out, np.stack(got + [alb]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("reduce", ["p2p", "nccl"])
def test_one_process_per_gpu_frames(tmp_path, reduce):
    """Two ranks, CUDA IPC pointers, the three-phase frame protocol through distributed.FrameRenderer.  With two
    GPUs the barriers are NCCL all-reduces on the shard's stream; on one GPU both ranks use device 0 and gloo."""
    import torch.multiprocessing as mp

    pod_path = scenes.write_scene("cornellGlass", str(tmp_path / "s.txt"), width=96, height=64)
    pod = api.Scene(pod_path).pod
    world, frames = 2, 4
    sums, ref_alb = _reference_sums(pod, world * frames)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "frames.npy")
    mp.spawn(_rank_worker, args=(world, port, pod_path, reduce, frames, 3, out), nprocs=world, join=True)
    got = np.load(out)
    for f in range(frames):
        ref = sums[(f + 1) * world - 1]
        if reduce == "p2p":
            assert_same_bits(ref, got[f], f"frame {f}")
        else:
            assert np.allclose(ref, got[f], rtol=1e-5, atol=1e-6), f"frame {f}"
    assert_same_bits(ref_alb, got[frames], "albedo")
