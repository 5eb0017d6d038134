"""Multi-GPU behind the C ABI (include/hanamaru_b200.h, "multi-GPU" section): hnm_group_* (one process, N devices, peer-copy
gather to device 0) and hnm_dist_* (one process per device, ncclAllGather inside the C layer).  Each pixel has one owner and
passes are added in order, so every variant must give the bits of a single-GPU render."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def single(hr, dev, scene, w, h, first, count):
    ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    ctx.render_passes(first, count)
    ctx.synchronize()
    acc, img, c = ctx.read_accum(), ctx.resolve(first + count - 1), ctx.counters()
    ctx.close()
    return acc, img, c


@pytest.mark.parametrize("members,tile", [(2, 4), (3, 8), (8, 2), (1, 4)])
def test_group_on_one_device_matches_single_renderer(hr, core, get_scene, get_device_scene, members, tile):
    """The whole group code path (one scene build, N uploads, async passes, peer-copy gather, deinterleave, resolve on member 0)
    with every member on device 0: C ABI only, bit-identical to one renderer."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h = 203, 117   # neither dimension divides the tiling
    want_acc, want_img, want_c = single(hr, dev, scene, w, h, 1, 3)
    g = hr.RenderGroup(scene, scene.camera, w, h, hr.MODE_PATHTRACING, devices=[0] * members, tile_rows=tile)
    g.render_passes(1, 2)
    g.render_passes(3, 1)
    img = g.resolve(3)
    acc = g.read_accum()
    c = g.counters()
    g.close()
    assert np.array_equal(bits(acc), bits(want_acc))
    assert np.array_equal(img, want_img)
    assert (c["paths"], c["segments"], c["shadow_rays"]) == (want_c["paths"], want_c["segments"], want_c["shadow_rays"])


def test_group_debug_mode_and_clear(hr, core, oracle, get_scene):
    scene = get_scene("rtcamp6")
    w, h = 160, 90
    g = hr.RenderGroup(scene, scene.camera, w, h, hr.MODE_DEBUG_NORMAL, devices=[0, 0, 0], tile_rows=4)
    g.render_passes(1, 1)
    img = g.resolve(1)
    want, _ = oracle.render(scene, w, h, hr.MODE_DEBUG_NORMAL, 1, 1, counters=False)
    assert np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, 1))
    g.clear()
    g.render_passes(1, 1)
    assert np.array_equal(g.resolve(1), img)
    g.close()


def test_group_across_devices(hr, core, get_scene, get_device_scene):
    """Two (or more) real devices in one process: shards cross NVLink as peer copies."""
    n = hr.device_count()
    if n < 2:
        pytest.skip("needs >= 2 CUDA devices (run with gpurun --gpus 2)")
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h = 640, 360
    want_acc, want_img, _ = single(hr, dev, scene, w, h, 1, 4)
    g = hr.RenderGroup(scene, scene.camera, w, h, hr.MODE_PATHTRACING, devices=list(range(min(n, 8))), tile_rows=4)
    g.render_passes(1, 4)
    assert np.array_equal(g.resolve(4), want_img)
    assert np.array_equal(bits(g.read_accum()), bits(want_acc))
    g.close()


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import hanamaru_renderer_b200 as hr
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = hr.device_count()
local = rank %% ndev
torch.cuda.set_device(local)
dist.init_process_group("gloo")          # side channel only: the data path is the C ABI's own NCCL communicator
scene = hr.build_scene("rtcamp6", hr.AssetStore.from_pack())
dev = hr.DeviceScene(scene, local)
w, h, passes = 320, 181, 3
ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING, shard=(rank, world, 4))
ids = [hr.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, 0)
ctx.dist_init(ids[0], rank, world)
ctx.render_passes(1, passes)
img = ctx.dist_resolve(passes, want_image=(rank == 0))
acc = ctx.dist_read_accum(want=(rank == 0))
# a process-lifetime communicator (hnm_comm_create) shared by two renderers in turn (hnm_dist_attach)
ids2 = [hr.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids2, 0)
comm = hr.DistComm(local, ids2[0], rank, world)
for k in range(2):
    c2 = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING, shard=(rank, world, 4))
    c2.dist_attach(comm)
    c2.render_passes(1, passes)
    img2 = c2.dist_resolve(passes, want_image=(rank == 0))
    if rank == 0 and not np.array_equal(img2, img):
        print("ATTACH_MISMATCH")
    c2.close()
comm.close()
if rank == 0:
    one = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    one.render_passes(1, passes); one.synchronize()
    ok = np.array_equal(one.read_accum().view(np.uint64), acc.view(np.uint64)) and np.array_equal(one.resolve(passes), img)
    print("DIST_OK" if ok else "DIST_MISMATCH")
dist.barrier()
'''


def test_dist_nccl_allgather_inside_the_abi(hr, core, tmp_path):
    """hnm_dist_*: two processes, two devices, the gather is ncclAllGather inside libhanamaru_b200.so."""
    if hr.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices (NCCL refuses two ranks on one device)")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=600)
    assert "DIST_OK" in out.stdout and "ATTACH_MISMATCH" not in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
