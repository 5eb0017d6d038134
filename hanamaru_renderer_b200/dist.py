"""Multi-GPU plumbing: one process per GPU, interleaved row tiles, ONE all-gather of the HDR framebuffer.

Every (pixel, sub-pixel, pass) is independent, so the path shards with no data-path
collective: rank k owns row tiles k, k+N, k+2N, ... (`hnm_shard`).  Each pixel has exactly one
owner, hence the gathered f64 buffer is bit-identical to a single-GPU render.  The only exchange
is the all-gather before `update_imgbuf` (the 3x3 bilateral filter needs neighbouring rows).
torch.distributed is plumbing only; the gather works on any backend (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

DEFAULT_TILE_ROWS = 4  # measured at N = 8, 1080p: 2 / 4 / 8 rows within 1.5 %; small tiles balance glass vs sky rows


def padded_rows(height, num_ranks, tile_rows=DEFAULT_TILE_ROWS):
    """Rows every rank allocates (equal on all ranks so that the all-gather is regular)."""
    if num_ranks == 1:
        return height
    ntiles = (height + tile_rows - 1) // tile_rows
    return ((ntiles + num_ranks - 1) // num_ranks) * tile_rows


def local_row_to_global(local_row, rank, num_ranks, tile_rows=DEFAULT_TILE_ROWS):
    """Image row of a rank's local row; >= height for padding rows.  Mirrors hnm_local_row_to_global."""
    lt, r = divmod(int(local_row), tile_rows)
    return (lt * num_ranks + rank) * tile_rows + r


def owned_rows(height, rank, num_ranks, tile_rows=DEFAULT_TILE_ROWS):
    """[(local_row, image_row)] for the rows of `rank` that exist in the image."""
    if num_ranks == 1:
        return [(y, y) for y in range(height)]
    out = []
    for lr in range(padded_rows(height, num_ranks, tile_rows)):
        y = local_row_to_global(lr, rank, num_ranks, tile_rows)
        if y < height:
            out.append((lr, y))
    return out


def deinterleave_numpy(gathered, height, num_ranks, tile_rows=DEFAULT_TILE_ROWS):
    """[num_ranks][padded_rows][W][3] -> [height][W][3] (host reference of the k_deinterleave kernel)."""
    gathered = np.asarray(gathered)
    n, pr, w, c = gathered.shape
    assert n == num_ranks
    full = np.zeros((height, w, c), gathered.dtype)
    for rank in range(num_ranks):
        for lr, y in owned_rows(height, rank, num_ranks, tile_rows):
            full[y] = gathered[rank, lr]
    return full


def all_gather_framebuffer(local, group=None):
    """local: torch tensor [padded_rows][W][3] f64 (any device) -> [world][padded_rows][W][3] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)  # concatenation along dim 0, rank-major
    return out.view((world,) + tuple(local.shape))


class DevicePointerTensor:
    """Wraps a raw CUDA allocation of the C ABI as a torch tensor (no copy) via __cuda_array_interface__."""

    def __init__(self, ptr, nbytes, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._nbytes = nbytes


def accum_as_tensor(ctx):
    """The renderer's device-resident accumulation shard as a torch CUDA tensor [padded_rows][W][3] (f64)."""
    import torch
    ptr, nbytes = ctx.accum_device_ptr()
    holder = DevicePointerTensor(ptr, nbytes, (ctx.owned_rows, ctx.width, 3))
    t = torch.as_tensor(holder, device="cuda:%d" % ctx.scene.device)
    assert t.data_ptr() == ptr
    return t


def gather_and_resolve(ctx, sampling, group=None):
    """All-gather the HDR shards over NCCL, scatter them into image order on the device and run the
    resolve kernels (replicated on every rank).  Returns the uint8 image and the full f64 buffer (torch)."""
    import torch
    import torch.distributed as dist
    local = accum_as_tensor(ctx)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ctx.resolve(sampling), local
    ctx.synchronize()
    torch.cuda.synchronize(local.device)
    gathered = all_gather_framebuffer(local, group)
    full = torch.empty((ctx.height, ctx.width, 3), dtype=torch.float64, device=local.device)
    torch.cuda.synchronize(local.device)
    ctx.deinterleave(gathered.data_ptr(), full.data_ptr())
    img = ctx.resolve(sampling, accum_full_device=full.data_ptr())
    return img, full
