"""GPU tests (-m gpu): several GPUs behind one renderer in one process (lfcuda_group_*, include/lfcuda.h).

A group deals the frames of a call round-robin to its devices and forms the image inside the post-process kernel of its first device,
which reads the other devices' accumulation buffers directly (peer access).  Two contexts may share a device, so everything except the
peer-access hop itself is exercised on a 1-GPU box; with >= 2 GPUs the same tests run across devices (NVLink)."""
import os

import numpy as np
import pytest

from oracle_api import Oracle, post_process
import lavaframe_b200 as lf

pytestmark = pytest.mark.gpu


def device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1], [1, 0]]
    if n >= 4:
        sets += [[0, 1, 2, 3]]
    if n >= 8:
        sets += [list(range(8))]
    return sets


@pytest.mark.parametrize("name", ["cornell", "c3mini"])
def test_group_image_is_the_sum_of_the_device_shares(gpu, golden_dir, oracle_lib, name):
    pack = lf.ScenePack(os.path.join(golden_dir, f"{name}.lfpack"))
    nframes = 13                                          # not a multiple of any group size: the shares are ragged
    pt = lf.PathTracer(0)
    pt.upload_pack(pack)
    pt.clear(); pt.render_frames(2, nframes)
    single = pt.read_accum()
    o = Oracle(pack.path)
    ref = o.render_frames(2, nframes)
    o.close()
    assert np.array_equal(single, ref)
    for devs in device_sets():
        n = len(devs)
        g = lf.PathTracerGroup(devs)
        g.upload_pack(pack)
        g.clear(); g.render_frames(2, nframes)
        img = g.read_accum()
        # what the group must hold, bit for bit: every device's strided share (frame order within a device), added in device order
        expect = None
        for i in range(n):
            cnt = (nframes - i + n - 1) // n
            pt.clear(); pt.render_frames(2 + i, cnt, n)
            share = pt.read_accum()
            expect = share if expect is None else expect + share
        assert np.array_equal(img, expect), f"devices {devs}: the group image is not the sum of its devices' shares"
        np.testing.assert_allclose(img, single, rtol=2e-5, atol=1e-5)
        if n == 1:
            assert np.array_equal(img, single)
        # rendering continues after a read-out (the local buffers were not disturbed), in several calls
        g.render_frames(2 + nframes, 3); g.render_frames(2 + nframes + 3, 4)
        more = g.read_accum()
        pt.clear(); pt.render_frames(2, nframes + 7)
        np.testing.assert_allclose(more, pt.read_accum(), rtol=2e-5, atol=1e-5)
        # the fused sum + post-process pass == the oracle's post-process of that sum (tonemap, vignette, chromatic aberration)
        post = lf.LfPostParams()
        post.use_vignette = 1; post.vignette_intensity = 0.4; post.vignette_power = 1.5
        post.use_ca = 1; post.use_ca_distortion = 1; post.ca_distance = 0.01; post.ca_p1 = 1.0; post.ca_p2 = 1.0; post.ca_p3 = 0.5
        g.set_post(post)
        inv = np.float32(1.0) / np.float32(nframes + 7)
        for tm in (0, 2):
            out = g.read_output(inv, tm)
            assert np.array_equal(out, post_process(more, inv, tm, post)), f"devices {devs}, tonemap {tm}"
            u8 = g.read_output_u8(inv, tm)
            np.testing.assert_array_equal(u8, np.rint(np.clip(out, 0, 1) * 255).astype(np.uint8))
        g.close()
    pt.close()


def test_group_rejects_bad_arguments(gpu):
    with pytest.raises(lf.LfCudaError):
        lf.PathTracerGroup([])
    with pytest.raises(lf.LfCudaError):
        lf.PathTracerGroup([9999])
    g = lf.PathTracerGroup([0, 0])
    with pytest.raises(lf.LfCudaError):
        g.render_frames(2, 1)                             # no scene yet
    g.close()
