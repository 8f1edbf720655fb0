"""GPU tests (-m gpu) of the spp split over several GPUs.

  * test_two_gpu_reduce_equals_single   (>= 2 devices, one process per GPU + NCCL: lfcuda_nccl_init / lfcuda_reduce) the reduced image
                                        equals the single-GPU render of the same frames up to fp32 summation order
  * test_c5_tile_converged              BASELINE config 5 in miniature: one tile of the 4K synthetic scene (C4, 20.57 M instanced
                                        triangles, depth 8) at 4096 spp IN TOTAL, the frames dealt to N renderers and summed, against the
                                        CPU oracle's 4096-spp tile: north_star check 3 (RMSE below 0.5 % of the mean luminance).
                                        With >= 2 devices: N processes + NCCL (the bench.py / torchrun deployment) and an in-process
                                        device group over NVLink peer access; on a 1-GPU box: an in-process group of two contexts.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from parity_metrics import rmse_over_mean_luminance

pytestmark = pytest.mark.gpu

C5_TILE = (40, 30)           # divides 3840 x 2160
C5_TILE_XY = (48, 36)        # centre of the frame: spheres, inter-reflection, glass
C5_SPP = 4096


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pack_path, nframes, out_dir, tile, tile_xy):
    import torch
    import torch.distributed as dist
    import lavaframe_b200 as lf
    from lavaframe_b200.multigpu import rank_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pt = lf.PathTracer(rank)
    over = dict(tile_width=tile[0], tile_height=tile[1]) if tile else {}
    pt.upload_pack(lf.ScenePack(pack_path), **over)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(pt.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    pt.nccl_init(idt.cpu().numpy().tobytes(), rank, world)
    f0, n, st = rank_frames(2, nframes, rank, world)
    pt.clear()
    pt.render_frames(f0, n, st, *tile_xy)
    pt.reduce()
    img = pt.read_accum()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), img)
    pt.close()
    dist.destroy_process_group()


def test_two_gpu_reduce_equals_single(tmp_path, golden_dir, gpu):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); bench.py's reduce_check proves the same in every multi-GPU bench run")
    import lavaframe_b200 as lf
    pack_path = os.path.join(golden_dir, "c2mini.lfpack")
    nframes = 16
    mp.spawn(_worker, args=(2, _free_port(), pack_path, nframes, str(tmp_path), None, (0, 0)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)                      # all-reduce: every rank holds the same sum
    pt = lf.PathTracer(0)
    pt.upload_pack(lf.ScenePack(pack_path))
    pt.clear(); pt.render_frames(2, nframes)
    single = pt.read_accum()
    pt.close()
    np.testing.assert_allclose(a, single, rtol=2e-5, atol=1e-5)


@pytest.fixture(scope="module")
def c5_reference(gpu, tmp_path_factory, oracle_lib):
    """The C4 / C5 scene and the oracle's 4096-spp render of one tile of its 4K frame (about 5 M pixel-samples on the host cores)."""
    from scenes import gen_scenes
    from oracle_api import Oracle
    out = str(tmp_path_factory.mktemp("c4_stress"))
    pack_path = gen_scenes.build_pack("c4_stress", out)
    o = Oracle(pack_path)
    o.update_params(tile_width=C5_TILE[0], tile_height=C5_TILE[1])
    ref = o.render_frames(2, C5_SPP, 1, *C5_TILE_XY)
    o.close()
    return pack_path, ref


def _crop(img):
    (tw, th), (tx, ty) = C5_TILE, C5_TILE_XY
    return img[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]


def test_c5_tile_converged(c5_reference, tmp_path):
    import torch
    import lavaframe_b200 as lf
    pack_path, ref = c5_reference
    assert _crop(ref).mean() > 0
    ndev = torch.cuda.device_count()
    results = {}
    # in-process device group (CudaRenderer(scene, dir, devices) / lf_render --gpus N): peer-access sum inside the post-process kernel
    for devs in ([0, 0], [0, 1] if ndev >= 2 else None, list(range(ndev)) if ndev > 2 else None):
        if devs is None:
            continue
        g = lf.PathTracerGroup(devs)
        g.upload_pack(lf.ScenePack(pack_path), tile_width=C5_TILE[0], tile_height=C5_TILE[1])
        g.clear(); g.render_frames(2, C5_SPP, 1, *C5_TILE_XY)
        results[f"group{devs}"] = g.read_accum()
        g.close()
    # one process per GPU + NCCL (bench.py under torchrun)
    if ndev >= 2:
        import torch.multiprocessing as mp
        world = 2
        mp.spawn(_worker, args=(world, _free_port(), pack_path, C5_SPP, str(tmp_path), C5_TILE, C5_TILE_XY), nprocs=world, join=True)
        results["nccl x2"] = np.load(tmp_path / "rank0.npy")
    for name, img in results.items():
        outside = img.copy(); _crop(outside)[:] = 0
        assert not outside.any(), name
        e = rmse_over_mean_luminance(_crop(img) / C5_SPP, _crop(ref) / C5_SPP)
        print(f"C5 tile, {name}: RMSE / mean luminance = {e:.3e} (bar 5e-3)")
        assert e < 5e-3, f"{name}: RMSE {e} of the mean luminance"
        np.testing.assert_allclose(_crop(img), _crop(ref), rtol=1e-4, atol=1e-3)   # and the summation-order bound, far inside the bar
