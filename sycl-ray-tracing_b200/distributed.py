"""Multi-GPU frame rendering: one process per GPU (torchrun), scene replicated, image cut into interleaved 16x16 tiles
(tile number % world == rank), one gather of the per-rank tile buffers to rank 0 per frame (NCCL over NVLink when the
tensors are CUDA tensors), then the un-tile kernel. The reference has no counterpart (it is one OpenMP process,
render_kernel.cpp:198); pixels are independent (per-pixel seed 31 + x*y*spp, :77), so there is no exchange step during
rendering and the N-rank image equals the 1-rank image bit for bit.

torch is plumbing only here: device buffers, streams and torch.distributed. The fill/untile steps are injectable so the
tile bookkeeping can be exercised on CPU with the gloo backend (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import numpy as np

TILE = 16
TILE_PX = 256


import os

TILE_SKEW = max(0, int(os.environ.get("B200RT_TILE_SKEW", "0")))      # csrc/device_types.h tile_xy (0 = row-major, the default: a skew measured slower): tile number L = ty * tiles_x + (tx + TILE_SKEW * ty) % tiles_x


def tile_xy(L, tiles_x):
    """Tile number -> (tx, ty) (device_types.h tile_xy); works on ints and numpy arrays."""
    ty = L // tiles_x
    return (L % tiles_x - TILE_SKEW * ty) % tiles_x, ty


def tile_number(tx, ty, tiles_x):
    return ty * tiles_x + (tx + TILE_SKEW * ty) % tiles_x


def tiles_for_rank(w: int, h: int, rank: int, world: int) -> int:
    n = ((w + TILE - 1) // TILE) * ((h + TILE - 1) // TILE)
    return (n - rank + world - 1) // world


def tile_slot_coords(w: int, h: int, rank: int, world: int, tiles_padded: int) -> np.ndarray:
    """(tiles_padded*256, 2) int32 pixel coordinates (x, y) of every slot of a rank's tile-major buffer; (-1, -1) for
    slots outside the frame or in padding tiles. Host-side statement of the kernels' unit_pixel() mapping."""
    tiles_x = (w + TILE - 1) // TILE
    out = np.full((tiles_padded * TILE_PX, 2), -1, np.int32)
    n = tiles_for_rank(w, h, rank, world)
    slot = np.arange(TILE_PX)
    sub, lane = slot // 32, slot % 32
    lx = (sub & 1) * 8 + (lane & 7)
    ly = (sub >> 1) * 4 + (lane >> 3)
    for k in range(n):
        tx, ty = tile_xy(rank + k * world, tiles_x)
        x, y = tx * TILE + lx, ty * TILE + ly
        ok = (x < w) & (y < h)
        out[k * TILE_PX:(k + 1) * TILE_PX, 0] = np.where(ok, x, -1)
        out[k * TILE_PX:(k + 1) * TILE_PX, 1] = np.where(ok, y, -1)
    return out


def untile_numpy(gathered: np.ndarray, w: int, h: int, world: int) -> np.ndarray:
    """numpy statement of the k_untile kernel (test helper for the CPU/gloo path)."""
    tiles_padded = gathered.shape[1] // TILE_PX
    img = np.zeros((h, w, gathered.shape[2]), gathered.dtype)
    for r in range(world):
        xy = tile_slot_coords(w, h, r, world, tiles_padded)
        ok = xy[:, 0] >= 0
        img[xy[ok, 1], xy[ok, 0]] = gathered[r][ok]
    return img


class FrameGatherer:
    """Owns the per-rank tile buffer, the gathered buffer and the final image (rank 0) and runs
    fill -> gather -> untile for one frame."""

    def __init__(self, w: int, h: int, rank: int, world: int, device, fill_tiles, untile, channels: int = 4, init_image=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.w, self.h, self.rank, self.world = w, h, rank, world
        self.tiles_padded = tiles_for_rank(w, h, 0, world)
        self.n_my_tiles = tiles_for_rank(w, h, rank, world)
        self.fill_tiles, self.untile, self.init_image = fill_tiles, untile, init_image
        self.tiles = torch.zeros((self.tiles_padded * TILE_PX, channels), dtype=torch.float32, device=device)
        self.gathered = None
        self.image = None
        if rank == 0:
            self.gathered = torch.zeros((world, self.tiles_padded * TILE_PX, channels), dtype=torch.float32, device=device)
            self.image = torch.zeros((h, w, channels), dtype=torch.float32, device=device)

    def frame(self, fb_in_host=None):
        """Renders this rank's tiles, gathers to rank 0 and un-tiles there. Returns the image tensor on rank 0, else None.
        fb_in_host (rank 0, optional): the incoming framebuffer as a pinned host tensor (h, w, 4), copied up inside the frame;
        default Color::Black(). The un-tile step does `framebuffer += final; tone map` (render_kernel.cpp:169-180) on it."""
        if self.rank == 0 and self.image is not None and self.init_image is not None:
            self.init_image(self.image, fb_in_host)
        self.fill_tiles(self.tiles)
        if self.world > 1:
            if self.rank == 0:
                self.dist.gather(self.tiles, gather_list=list(self.gathered.unbind(0)), dst=0)
            else:
                self.dist.gather(self.tiles, gather_list=None, dst=0)
            src = self.gathered
        else:
            src = self.tiles.unsqueeze(0)
        if self.rank == 0:
            self.untile(src, self.image)
            return self.image
        return None


def make_cuda_gatherer(scene, camera, w, h, spp, max_bounces, rank, world, device, integrator=0, flags=0):
    """FrameGatherer whose fill/untile steps are the CUDA kernels, launched on torch's current stream. Every rank renders its tiles
    as mean radiance (B200RT_FLAG_LINEAR_TILES); rank 0 holds the incoming framebuffer and does `framebuffer += final; tone map`
    while un-tiling the gathered buffer (b200rt_untile_accumulate_device), so the incoming framebuffer is honoured without
    shipping it to every rank."""
    import torch
    from .binding import FLAG_LINEAR_TILES

    def fill(tiles):
        st = torch.cuda.current_stream(device).cuda_stream
        scene.render_tiles_device(camera, w, h, spp, max_bounces, tiles.data_ptr(), stream_ptr=st, integrator=integrator,
                                  flags=flags | FLAG_LINEAR_TILES, rank=rank, world=world)

    g = None
    black = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float32, device=device)

    def init_image(image, fb_in_host):
        if fb_in_host is not None:
            image.copy_(fb_in_host, non_blocking=True)
        else:
            image.copy_(black.expand_as(image))           # Color::Black(), image.h:34

    def untile(src, image):
        st = torch.cuda.current_stream(device).cuda_stream
        scene.untile_accumulate_device(src.data_ptr(), g.tiles_padded, world, w, h, image.data_ptr(), stream_ptr=st)

    g = FrameGatherer(w, h, rank, world, device, fill, untile, init_image=init_image)
    return g
