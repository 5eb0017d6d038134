"""CPU test of the N>1 path: world_size 2 over gloo.  Each rank renders only the row tiles it owns (the
oracle stands in for the GPU kernels -- this is test code), the HDR shards are all-gathered through the same
`dist.all_gather_framebuffer` the NCCL path uses, de-interleaved, and must be BIT-identical to a one-process render."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, tile_rows, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import hanamaru_renderer_b200 as hr
    from hanamaru_renderer_b200 import dist as hd
    from oracle_ffi import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = Oracle("det")
        scene = hr.build_scene("diamond", hr.AssetStore.from_pack())
        pr = hd.padded_rows(h, world, tile_rows)
        local = np.zeros((pr, w, 3), np.float64)
        full_scratch = np.zeros((h, w, 3), np.float64)
        for lr, y in hd.owned_rows(h, rank, world, tile_rows):
            full_scratch[:] = 0
            oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, 2, accum=full_scratch, rows=(y, y + 1), counters=False)
            local[lr] = full_scratch[y]
        gathered = hd.all_gather_framebuffer(torch.from_numpy(local)).numpy()
        full = hd.deinterleave_numpy(gathered, h, world, tile_rows)
        np.save(os.path.join(out_dir, "full_%d.npy" % rank), full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("h,tile_rows", [(20, 8), (16, 4), (7, 8)])
def test_two_rank_gather_is_bit_identical(tmp_path, h, tile_rows):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hanamaru_renderer_b200 as hr
    from oracle_ffi import Oracle
    w, world = 24, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, w, h, tile_rows, str(tmp_path)), nprocs=world, join=True)
    want, _ = Oracle("det").render(hr.build_scene("diamond", hr.AssetStore.from_pack()), w, h, hr.MODE_PATHTRACING, 1, 2, counters=False)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "full_%d.npy" % rank))
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), rank


def test_row_mapping_matches_c_abi_semantics():
    from hanamaru_renderer_b200 import dist as hd
    for h, n, t in [(1080, 8, 8), (270, 2, 8), (7, 4, 8), (2160, 8, 16), (10, 3, 2)]:
        pr = hd.padded_rows(h, n, t)
        seen = []
        for r in range(n):
            rows = hd.owned_rows(h, r, n, t)
            assert all(lr < pr for lr, _ in rows)
            ys = [y for _, y in rows]
            assert ys == sorted(ys)
            # rows that exist form a prefix of the local rows (the kernels rely on it)
            assert [lr for lr, _ in rows] == list(range(len(rows)))
            seen += ys
        assert sorted(seen) == list(range(h))
    assert hd.padded_rows(1080, 1, 8) == 1080 and hd.owned_rows(3, 0, 1) == [(0, 0), (1, 1), (2, 2)]
