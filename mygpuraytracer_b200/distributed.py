"""``pathtrace()`` over several GPUs with one process per GPU: ``torch.distributed`` plumbing around ``B2ptShard``.

A FRAME is one iteration per rank: rank r of a frame that starts at iteration i renders i + r (the reference's
iterations are independent; their only shared state is ``image[pixel] += color * PI``,
apps/src/pathtrace.cu:508).  Every rank calls :meth:`FrameRenderer.pathtrace` with the same argument; after the
call rank 0 holds the running sum including all iterations of the frame -- on the device, and in host memory if
it passed arrays (the reference's D2H of apps/src/pathtrace.cu:662-668, once per frame and on ONE rank).

``reduce="p2p"`` (default): the ranks exchange CUDA IPC handles once; a frame is combined by ``k_frame_reduce``
(csrc/multi.cu) over NVLink peer memory -- every rank reduces its own pixel slice of all contributions in
iteration order -- between two 4-byte NCCL all-reduces that act as stream-ordered barriers.  The image is
bit-identical to a single GPU's.  ``reduce="nccl"``: one ``dist.reduce`` of the contribution to rank 0 per frame,
then rank 0 merges (float summation order differs from the sequential one).

torch is used for the process group, the barriers and the collective only; all rendering goes through
``libb2pt.so``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import abi, api


class _DevArray:
    """A raw device pointer as ``__cuda_array_interface__`` (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


class FrameRenderer:
    def __init__(self, scene, options: Optional[abi.Options] = None, lanes: int = 4, reduce: str = "p2p", group=None,
                 shard=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if reduce not in ("p2p", "nccl"):
            raise ValueError("reduce must be 'p2p' or 'nccl'")
        # `shard`: an object with the Shard interface (the CPU tests pass one backed by the oracle)
        self.shard = shard if shard is not None else api.Shard(scene, options, rank=self.rank, world=self.world, lanes=lanes)
        self.n_pixels = self.shard.n_pixels
        self.on_gpu = shard is None
        self.reduce = reduce
        self._stream = None
        self._token = None
        # with a CPU backend (gloo) the barriers are host barriers after a stream sync: same protocol, no overlap
        # (used by the tests, which also run two ranks on ONE GPU, something NCCL refuses)
        self.host_barrier = dist.is_initialized() and dist.get_backend(group) != "nccl"
        if self.on_gpu:
            dev = torch.device("cuda", self.shard.options.device)
            self._stream = torch.cuda.ExternalStream(self.shard.stream_ptr(), device=dev)
            self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.world > 1 and reduce == "p2p":
            self._connect()

    # -- setup ---------------------------------------------------------------------------------
    def _connect(self) -> None:
        torch, dist = self.torch, self.dist
        blob = self.shard.export()
        dev = torch.device("cpu") if self.host_barrier else self._token.device
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.group)
        self.shard.connect([bytes(t.cpu().numpy().tobytes()) for t in gathered])
        dist.barrier(group=self.group)

    def _barrier(self) -> None:
        """Stream-ordered rendezvous of all ranks on the shard's stream: a 4-byte all-reduce."""
        if self.world == 1:
            return
        if self.host_barrier:
            self.shard.sync()
            self.dist.barrier(group=self.group)
            return
        with self.torch.cuda.stream(self._stream):
            self.dist.all_reduce(self._token, group=self.group)

    # -- one frame ------------------------------------------------------------------------------
    def pathtrace(self, first_iteration: int, image: Optional[np.ndarray] = None, albedo: Optional[np.ndarray] = None) -> None:
        s = self.shard
        s.frame_begin(first_iteration)
        if self.reduce == "p2p" or self.world == 1:
            self._barrier()
            s.frame_reduce()
            self._barrier()
        else:
            self._reduce_contribution()
            s.frame_merge()
        if self.rank == 0:
            s.frame_end(image, albedo)
        else:
            s.frame_end(None, None)

    def _reduce_contribution(self) -> None:
        torch, dist = self.torch, self.dist
        if self.on_gpu and self.host_barrier:
            t = torch.as_tensor(_DevArray(self.shard.frame_image_ptr(), self.n_pixels * 3), device=self._token.device)
            self.shard.sync()
            h = t.cpu()
            dist.reduce(h, dst=0, group=self.group)
            t.copy_(h)
            torch.cuda.synchronize()
        elif self.on_gpu:
            t = torch.as_tensor(_DevArray(self.shard.frame_image_ptr(), self.n_pixels * 3), device=self._token.device)
            with torch.cuda.stream(self._stream):
                dist.reduce(t, dst=0, group=self.group)
        else:
            dist.reduce(self.shard.frame_image_tensor(), dst=0, group=self.group)

    def sync(self) -> None:
        self.shard.sync()

    def close(self) -> None:
        self.shard.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
