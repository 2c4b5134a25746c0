"""not gpu: the N>1 frame path (interleaved tiles -> gather to rank 0 -> un-tile) on CPU with the gloo backend,
world_size 2 and 3. The CUDA fill/untile steps are replaced by numpy statements of the same mapping."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    from sycl_ray_tracing_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fill(tiles):
        xy = D.tile_slot_coords(w, h, rank, world, tiles.shape[0] // D.TILE_PX)
        v = np.zeros((len(xy), 4), np.float32)
        ok = xy[:, 0] >= 0
        v[ok, 0] = xy[ok, 0]; v[ok, 1] = xy[ok, 1]; v[ok, 2] = rank; v[ok, 3] = xy[ok, 1] * w + xy[ok, 0]
        tiles.copy_(torch.from_numpy(v))

    def untile(src, image):
        image.copy_(torch.from_numpy(D.untile_numpy(src.numpy(), w, h, world)))

    g = D.FrameGatherer(w, h, rank, world, torch.device("cpu"), fill, untile)
    for _ in range(2):          # two frames: buffers are reused
        img = g.frame()
    if rank == 0:
        np.save(out_path, img.numpy())

    # the accumulating flavour (what make_cuda_gatherer wires up): rank 0 starts every frame from the incoming framebuffer and the
    # un-tile step adds the gathered tiles to it — a second frame must not see the first one's result
    def init_image(image, fb_in_host):
        image.copy_(fb_in_host if fb_in_host is not None else torch.zeros_like(image))

    def untile_add(src, image):
        image.add_(torch.from_numpy(D.untile_numpy(src.numpy(), w, h, world)))

    g2 = D.FrameGatherer(w, h, rank, world, torch.device("cpu"), fill, untile_add, init_image=init_image)
    fb = torch.full((h, w, 4), 1000.0) if rank == 0 else None
    for _ in range(2):
        img2 = g2.frame(fb_in_host=fb)
    if rank == 0:
        np.save(out_path.replace(".npy", "_acc.npy"), img2.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 100, 70), (3, 64, 48), (2, 33, 17)])
def test_gloo_tile_gather_reassembles_frame(tmp_path, world, w, h):
    import torch.multiprocessing as mp
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, out), nprocs=world, join=True)
    img = np.load(out)
    ys, xs = np.mgrid[0:h, 0:w]
    assert np.array_equal(img[..., 0], xs) and np.array_equal(img[..., 1], ys)
    assert np.array_equal(img[..., 3], ys * w + xs)
    tiles_x = (w + 15) // 16
    from sycl_ray_tracing_b200 import distributed as D
    owner = D.tile_number(xs // 16, ys // 16, tiles_x) % world
    assert np.array_equal(img[..., 2], owner), "pixel written by the rank that owns its tile"
    acc = np.load(out.replace(".npy", "_acc.npy"))
    assert np.array_equal(acc, img + 1000.0), "framebuffer += gathered tiles, once per frame"
