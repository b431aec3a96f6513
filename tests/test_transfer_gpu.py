"""Transfer queues of the C-ABI (vhr_image_upload_async / vhr_image_download_async / vhr_wait_download): the copies that
overlap the compute stream must deliver the same bytes as the in-order ones, and a download must be a snapshot."""
import numpy as np
import pytest

from vulkanhybridrenderer_b200 import capi
from vulkanhybridrenderer_b200 import types as T

pytestmark = pytest.mark.gpu

F4 = T.VK_FORMAT_R16G16B16A16_SFLOAT


def test_async_upload_is_consumed_by_the_next_user():
    import torch
    W, H = 256, 128
    rng = np.random.default_rng(7)
    with capi.Context(W, H) as ctx:
        ctx.actualize_image("a", F4)
        slot = ctx.upload_new_storage_image(W, H, F4)
        for it in range(4):
            src = torch.from_numpy(rng.integers(0, 0x3c00, (H, W, 4), dtype=np.uint16).view(np.int16)).pin_memory()
            ctx.image_upload_async("a", src)
            ctx.blit_transient_to_storage("a", slot)                      # first user: must see the uploaded bytes
            back = ctx.storage_image_download(slot)
            assert np.array_equal(back.view(np.uint16), src.numpy().view(np.uint16)), it


def test_async_download_is_a_snapshot():
    import torch
    W, H = 512, 256
    rng = np.random.default_rng(8)
    with capi.Context(W, H) as ctx:
        ctx.actualize_image("a", F4)
        outs = [torch.zeros(H, W, 4, dtype=torch.int16).pin_memory() for _ in range(3)]
        imgs = [rng.integers(0, 0x3c00, (H, W, 4), dtype=np.uint16) for _ in range(3)]
        tickets = []
        for img, out in zip(imgs, outs):
            ctx.image_upload("a", img.view(np.float16))
            tickets.append(ctx.image_download_async("a", out))             # the next upload overwrites "a" right away
        for t in tickets:
            ctx.wait_download(t)
        for img, out in zip(imgs, outs):
            assert np.array_equal(out.numpy().view(np.uint16), img)
        with pytest.raises(Exception):
            ctx.wait_download(tickets[-1] + 1)                             # never issued
