"""GPU test (-m gpu, needs >= 2 devices; skipped otherwise): spp split over two GPUs with the library's own NCCL sum
(lfcuda_nccl_init / lfcuda_reduce) equals the single-GPU render of the same frames up to fp32 summation order."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pack_path, nframes, out_dir):
    import torch
    import torch.distributed as dist
    import lavaframe_b200 as lf
    from lavaframe_b200.multigpu import rank_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pt = lf.PathTracer(rank)
    pt.upload_pack(lf.ScenePack(pack_path))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(pt.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    pt.nccl_init(idt.cpu().numpy().tobytes(), rank, world)
    f0, n, st = rank_frames(2, nframes, rank, world)
    pt.clear()
    pt.render_frames(f0, n, st)
    pt.reduce()
    img = pt.read_accum()
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), img)
    pt.close()
    dist.destroy_process_group()


def test_two_gpu_reduce_equals_single(tmp_path, golden_dir, gpu):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import lavaframe_b200 as lf
    pack_path = os.path.join(golden_dir, "c2mini.lfpack")
    nframes = 16
    mp.spawn(_worker, args=(2, _free_port(), pack_path, nframes, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)                      # all-reduce: every rank holds the same sum
    pt = lf.PathTracer(0)
    pt.upload_pack(lf.ScenePack(pack_path))
    pt.clear(); pt.render_frames(2, nframes)
    single = pt.read_accum()
    pt.close()
    np.testing.assert_allclose(a, single, rtol=2e-5, atol=1e-5)
