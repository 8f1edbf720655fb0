"""CPU tests of the multi-GPU path: frame dealing + the accumulator sum, world_size 2 over gloo.

The renderer stand-in is the CPU oracle on a tiny frame (tests may use it); what is under test is the host logic the
GPU path shares: lavaframe_b200.multigpu.rank_frames / reduce_accumulators."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lavaframe_b200.multigpu import rank_frames, all_frames, reduce_accumulators  # noqa: E402


def test_frames_partition():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 8, 64, 4096):
            lists = all_frames(2, n, world)
            flat = sorted(f for l in lists for f in l)
            assert flat == list(range(2, 2 + n))
            assert max(len(l) for l in lists) - min(len(l) for l in lists) <= 1
    with pytest.raises(ValueError):
        rank_frames(2, 8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pack, nframes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_api import Oracle
    o = Oracle(pack)
    o.update_params(width=32, height=32, tile_width=32, tile_height=32)
    f0, n, st = rank_frames(2, nframes, rank, world)
    acc = o.render_frames(f0, n, st)
    t = torch.from_numpy(acc)
    reduce_accumulators(t, dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    o.close()
    dist.destroy_process_group()


def test_two_rank_spp_split_equals_single(tmp_path, golden_dir, oracle_lib):
    pack = os.path.join(golden_dir, "cornell.lfpack")
    nframes = 6
    mp.spawn(_worker, args=(2, _free_port(), pack, nframes, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    from oracle_api import Oracle
    o = Oracle(pack)
    o.update_params(width=32, height=32, tile_width=32, tile_height=32)
    single = o.render_frames(2, nframes)
    o.close()
    # same samples, different summation order (per-rank partial sums first): equal up to fp32 reassociation
    np.testing.assert_allclose(reduced, single, rtol=1e-5, atol=1e-6)
    assert reduced.sum() > 0
