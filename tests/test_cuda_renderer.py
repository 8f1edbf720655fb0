"""GPU tests (-m gpu): the C++ CudaRenderer (drop-in for TiledRenderer) driven through the reference's Renderer
interface by Main.cpp's own call order, on the reference's own scene file loaded by its unchanged loader."""
import os

import numpy as np
import pytest

import lavaframe_b200 as lf
from lavaframe_b200.capi import lib_path
from parity_metrics import radiance_agreement

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cornell_scene(gpu, tmp_path_factory):
    if not os.path.exists(lib_path("liblfhost.so")):
        pytest.fail("liblfhost.so missing: __graft_entry__.build() must run where /root/reference exists")
    from scenes import gen_scenes
    path = gen_scenes.cornell_256(str(tmp_path_factory.mktemp("cornell")))
    s = lf.HostScene(path)
    yield s
    s.close()


def test_main_loop_protocol(cornell_scene, golden_dir):
    g = np.load(os.path.join(golden_dir, "cornell_llvmpipe.npz"))
    r = lf.CudaRenderer(cornell_scene)
    assert r.GetSampleCount() == 1                       # TiledRenderer.cpp:55
    r.Update(0.0); r.Render()
    assert r.GetSampleCount() == 1                       # sample 1 drawn with frame 2, not yet "completed"
    r.Update(0.0)
    assert r.GetSampleCount() == 2
    img = r.GetOutputBufferHDR()                         # image of the last COMPLETED sample = frame 2 alone
    assert radiance_agreement(img, g["spp1"]) >= 0.999
    r.close()

    r = lf.CudaRenderer(cornell_scene)
    n = int(g["nspp"])
    steps = r.Run(n)                                     # Main.cpp loop until maxSamples + 1 == GetSampleCount()
    assert steps == n and r.GetSampleCount() == n + 1
    img = r.GetOutputBufferHDR()
    assert radiance_agreement(img, g["sppN"]) >= 0.99
    u8 = r.GetOutputBuffer()
    assert u8.shape == (256, 256, 3) and u8.dtype == np.uint8
    np.testing.assert_array_equal(u8, np.rint(np.clip(img, 0, 1) * 255).astype(np.uint8))
    r.close()


def test_failed_init_is_reported_not_hung(cornell_scene):
    """A renderer whose Init fails (here: a device that does not exist) says so: Ok() is false, LastError() has the text, the Main.cpp loop
    refuses to spin on a sample counter that will never advance, the output buffers are black, and nothing leaks or crashes on destruction."""
    with pytest.raises(lf.LfCudaError, match="out of range|no CUDA device"):
        lf.CudaRenderer(cornell_scene, device=9999)
    with pytest.raises(lf.LfCudaError):
        lf.CudaRenderer(cornell_scene, devices=[0, 9999])
    lib = cornell_scene.lib
    h = lib.lfhost_renderer_create(cornell_scene.h, 9999)          # the raw object, as a C++ caller would hold it
    assert lib.lfhost_renderer_ok(h) == 0
    assert lib.lfhost_renderer_error(h)
    assert lib.lfhost_renderer_run(h, 4) == -1
    import ctypes as C
    w, hh = C.c_int(), C.c_int()
    out = np.ones((256, 256, 3), np.float32)
    lib.lfhost_renderer_output_hdr(h, out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(hh))
    assert (w.value, hh.value) == (256, 256) and not out.any()
    lib.lfhost_renderer_destroy(h)
    r = lf.CudaRenderer(cornell_scene)                             # and the scene is still usable
    assert r.Run(1) == 1
    r.close()


def test_renderer_equals_c_abi(cornell_scene, golden_dir):
    """The C++ class adds only bookkeeping: same bits as direct lfcuda_render_frames calls."""
    r = lf.CudaRenderer(cornell_scene)
    r.Run(5)
    a = r.GetOutputBufferHDR()
    r.close()
    pt = lf.PathTracer(0)
    pt.upload_pack(lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack")))
    pt.clear(); pt.render_frames(2, 5)
    b = pt.read_output(1.0 / 5, 0)
    pt.close()
    assert np.array_equal(a, b)


def test_camera_move_resets_accumulation(cornell_scene):
    r = lf.CudaRenderer(cornell_scene)
    r.Run(3)
    assert r.GetSampleCount() == 4
    cornell_scene.set_camera_moving(True)
    r.Update(0.0); r.Render()                            # TiledRenderer.cpp:471-484: counters reset, accumulation cleared
    assert r.GetSampleCount() == 1
    cornell_scene.set_camera_moving(False)
    r.Run(2)
    a = r.GetOutputBufferHDR()
    r.close()
    r2 = lf.CudaRenderer(cornell_scene)
    r2.Run(2)
    assert np.array_equal(a, r2.GetOutputBufferHDR())
    r2.close()


def test_instance_edit_path(cornell_scene, golden_dir):
    """Scene::RebuildInstances -> Renderer::Update re-uploads transforms/materials/TLAS (Renderer.cpp:190-205)."""
    r = lf.CudaRenderer(cornell_scene)
    r.Run(2)
    before = r.GetOutputBufferHDR()
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    m = pack.transforms.reshape(-1, 16)[3].copy()        # the small box
    moved = m.copy(); moved[13] += 0.1                   # lift it by 0.1
    cornell_scene.move_instance(3, moved)
    r.Update(0.0); r.Render()
    assert r.GetSampleCount() == 1
    r.Run(2)
    after = r.GetOutputBufferHDR()
    assert not np.array_equal(before, after)
    cornell_scene.move_instance(3, m)                    # and back: identical to the original render
    r.Update(0.0); r.Render()
    r.Run(2)
    assert np.array_equal(r.GetOutputBufferHDR(), before)
    r.close()


def _oracle_image(scene, tmp_path, name, spp, tonemap=0):
    """What the reference would show after `spp` samples of the scene AS IT IS NOW: the host scene re-flattened by the reference's
    own code (lfhost_write_pack -> Scene's arrays after RebuildInstances), rendered by the CPU oracle, through postprocess.glsl."""
    from oracle_api import Oracle, post_process
    path = str(tmp_path / f"{name}.lfpack")
    scene.write_pack(path)
    o = Oracle(path)
    acc = o.render_frames(2, spp)
    o.close()
    return post_process(acc, np.float32(1.0) / np.float32(spp), tonemap)


def test_instance_edit_vs_oracle(cornell_scene, golden_dir, tmp_path, oracle_lib):
    """The instance-edit path against the ORACLE (Renderer.cpp:190-205 after Scene::RebuildInstances, Scene.cpp:165-178): after the
    move, CudaRenderer's image (transforms + materials + rebuilt TLAS re-uploaded by lfcuda_update_instances) equals, bit for bit,
    the oracle's render of the scene re-flattened from scratch; likewise after the move back."""
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    m = pack.transforms.reshape(-1, 16)[3].copy()
    r = lf.CudaRenderer(cornell_scene)
    r.Run(2)
    assert np.array_equal(r.GetOutputBufferHDR(), _oracle_image(cornell_scene, tmp_path, "before", 2))
    moved = m.copy(); moved[12] -= 0.15; moved[13] += 0.1; moved[14] += 0.05
    cornell_scene.move_instance(3, moved)
    r.Update(0.0); r.Render()
    r.Run(3)
    after = r.GetOutputBufferHDR()
    ref = _oracle_image(cornell_scene, tmp_path, "moved", 3)
    assert not np.array_equal(ref, _oracle_image(cornell_scene, tmp_path, "moved2", 2))
    differ = int((after != ref).any(axis=2).sum())
    assert differ == 0, f"after the instance move {differ} pixels differ from the oracle's render of the re-flattened scene"
    cornell_scene.move_instance(3, m)
    r.Update(0.0); r.Render()
    r.Run(2)
    assert np.array_equal(r.GetOutputBufferHDR(), _oracle_image(cornell_scene, tmp_path, "back", 2))
    r.close()


def test_instance_edit_vs_oracle_many_instances(gpu, tmp_path_factory, tmp_path, oracle_lib):
    """The same on a scene with a real TLAS (c4_gold: 196 instances, depth 8, glass / metal / diffuse): three instances are moved and
    non-uniformly rescaled one after the other (three RebuildInstances, three TLAS re-uploads); CUDA == oracle bit for bit."""
    from scenes import gen_scenes
    path = gen_scenes.c4_gold(str(tmp_path_factory.mktemp("c4gold")))
    s = lf.HostScene(path)
    r = lf.CudaRenderer(s)
    r.Run(1)
    assert np.array_equal(r.GetOutputBufferHDR(), _oracle_image(s, tmp_path, "g0", 1))
    v, _, _ = s.views()
    T = np.ctypeslib.as_array(v.transforms, shape=(v.num_instances, 16)).copy()
    rng = np.random.RandomState(5)
    for k, idx in enumerate((7, 100, 195)):
        mtx = T[idx].copy()
        mtx[12:15] += rng.uniform(-1.5, 1.5, 3).astype(np.float32)
        mtx[0] *= 1.5; mtx[5] *= 0.7
        s.move_instance(idx, mtx)
        r.Update(0.0); r.Render()
        assert r.GetSampleCount() == 1
    r.Run(2)
    img = r.GetOutputBufferHDR()
    ref = _oracle_image(s, tmp_path, "g1", 2)
    differ = int((img != ref).any(axis=2).sum())
    assert differ == 0, f"{differ} pixels differ from the oracle after three instance edits"
    r.close()
    s.close()


def test_instance_edit_with_device_tlas(gpu, tmp_path_factory, tmp_path, oracle_lib):
    """CudaRenderer::SetDeviceTlasRebuild: the instance-edit path with the TLAS rebuilt on the device equals the oracle's render of the
    scene the reference's host code re-flattened (c4_gold, 196 instances)."""
    from scenes import gen_scenes
    path = gen_scenes.c4_gold(str(tmp_path_factory.mktemp("c4gold_dev")))
    s = lf.HostScene(path)
    r = lf.CudaRenderer(s)
    r.SetDeviceTlasRebuild(True)
    r.Run(1)
    v, _, _ = s.views()
    T = np.ctypeslib.as_array(v.transforms, shape=(v.num_instances, 16)).copy()
    rng = np.random.RandomState(9)
    for idx in (3, 77, 150, 195):
        mtx = T[idx].copy()
        mtx[12:15] += rng.uniform(-1.5, 1.5, 3).astype(np.float32)
        mtx[0] *= 1.3; mtx[10] *= 0.6
        s.move_instance(idx, mtx)
        r.Update(0.0); r.Render()
    r.Run(2)
    img = r.GetOutputBufferHDR()
    ref = _oracle_image(s, tmp_path, "devtlas", 2)
    differ = int((img != ref).any(axis=2).sum())
    assert differ == 0, f"{differ} pixels differ from the oracle after instance edits with the device-built TLAS"
    r.close()
    s.close()


def test_preview_while_camera_moves(cornell_scene, golden_dir):
    """TiledRenderer::Render draws the preview engine while camera->isMoving (TiledRenderer.cpp:327-333) at
    screenSize * GlobalState.previewScale; the image Present() would show must equal the reference's previewFBO."""
    g = np.load(os.path.join(golden_dir, "cornell_llvmpipe_preview.npz"))
    cornell_scene.set_preview(0.5, False)
    r = lf.CudaRenderer(cornell_scene)
    assert r.GetPreviewBufferHDR() is None
    r.Run(2)
    cornell_scene.set_camera_moving(True)
    r.Update(0.0); r.Render()
    img = r.GetPreviewBufferHDR()
    cornell_scene.set_camera_moving(False)
    assert img.shape == g["half"].shape
    assert radiance_agreement(img, g["half"]) >= 0.999
    r.Run(2)                                             # and the path tracer restarts from a cleared accumulation
    a = r.GetOutputBufferHDR()
    r.close()
    cornell_scene.set_preview(1.0, True)
    r = lf.CudaRenderer(cornell_scene)
    cornell_scene.set_camera_moving(True)
    r.Update(0.0); r.Render()
    img = r.GetPreviewBufferHDR()
    cornell_scene.set_camera_moving(False)
    assert radiance_agreement(img, g["full_dof"]) >= 0.999
    r.Run(2)
    assert np.array_equal(a, r.GetOutputBufferHDR())
    r.close()
    cornell_scene.set_preview(1.0, False)


def _read_exr_bgr(path):
    """Minimal reader of an uncompressed scan-line OpenEXR file with FLOAT channels B, G, R -> (H, W, 3) RGB float32."""
    import struct
    b = open(path, "rb").read()
    assert b[:4] == bytes([0x76, 0x2f, 0x31, 0x01]) and b[4] == 2
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(b"\0", pos); name = b[pos:e].decode(); pos = e + 1
        e = b.index(b"\0", pos); typ = b[pos:e].decode(); pos = e + 1
        size, = struct.unpack_from("<i", b, pos); pos += 4
        attrs[name] = (typ, b[pos:pos + size]); pos += size
    pos += 1
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    ch, names, p = attrs["channels"][1], [], 0
    while ch[p] != 0:
        e = ch.index(b"\0", p); names.append(ch[p:e].decode()); p = e + 1
        assert struct.unpack_from("<i", ch, p)[0] == 2                     # FLOAT
        p += 16
    assert names == ["B", "G", "R"]
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack_from(f"<{h}Q", b, pos)
    img = np.empty((h, w, 3), np.float32)
    for y in range(h):
        yy, size = struct.unpack_from("<2i", b, offs[y])
        assert yy == y and size == 3 * w * 4
        planes = np.frombuffer(b, np.float32, 3 * w, offs[y] + 8).reshape(3, w)
        img[y, :, 2], img[y, :, 1], img[y, :, 0] = planes[0], planes[1], planes[2]
    return img


def test_headless_cxx_driver(golden_dir, gpu, tmp_path):
    """lavaframe_b200/bin/lf_render: Main.cpp's Update -> Render loop on a CudaRenderer in pure C++ (no Python between the
    reference's loader, the drop-in class and the C ABI).  Its GetOutputBufferHDR image equals the llvmpipe golden within the
    north_star bar, and a tiled run (tileWidth/tileHeight from the scene file's renderer block) equals the untiled one."""
    import json
    import subprocess
    exe = lib_path(os.path.join("bin", "lf_render"))
    if not os.path.exists(exe):
        pytest.fail("lavaframe_b200/bin/lf_render missing: __graft_entry__.build() must run where /root/reference exists")
    from scenes import gen_scenes
    g = np.load(os.path.join(golden_dir, "cornell_llvmpipe.npz"))
    n = int(g["nspp"])
    scene = gen_scenes.cornell_256(str(tmp_path / "cornell"))
    out = str(tmp_path / "img.f32")
    png, bmp, exr = str(tmp_path / "img.png"), str(tmp_path / "img.bmp"), str(tmp_path / "img.exr")
    res = subprocess.run([exe, scene, "--spp", str(n), "--out", out, "--png", png, "--bmp", bmp, "--exr", exr], check=True, capture_output=True, text=True, timeout=300)
    info = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert (info["width"], info["height"], info["spp"], info["tile_steps"]) == (256, 256, n, n)
    img = np.fromfile(out, np.float32).reshape(256, 256, 3)
    assert radiance_agreement(img, g["sppN"]) >= 0.99
    # SaveFrame / SaveFrameBMP (Export.h:14-57): GetOutputBuffer's 8-bit image, flipped to top row first
    from PIL import Image
    want = np.rint(np.clip(img, 0, 1) * 255).astype(np.uint8)[::-1]
    for f in (png, bmp):
        got = np.asarray(Image.open(f).convert("RGB"))
        assert got.shape == (256, 256, 3) and np.array_equal(got, want), f
    # SaveFrameEXR (Export.h:57-152): the float image as B, G, R FLOAT planes, rows as GetOutputBufferHDR returns them (no flip)
    assert np.array_equal(_read_exr_bgr(exr), img)
    # the same scene with 64x64 tiles: 16 tile steps per sample and a new `frame` (RNG seed) for every tile step, against the
    # reference's own tiled run on llvmpipe
    gt = np.load(os.path.join(golden_dir, "cornell_tiled_llvmpipe.npz"))
    nt, tile = int(gt["nspp"]), int(gt["tile"])
    text = open(scene).read()
    assert "resolution 256 256" in text
    tiled = os.path.join(os.path.dirname(scene), "cornell_tiled.scene")      # next to the .obj files it names
    open(tiled, "w").write(text.replace("resolution 256 256", f"resolution 256 256\n\ttileWidth {tile}\n\ttileHeight {tile}"))   # Loader.cpp:256-261
    out2 = str(tmp_path / "img2.f32")
    res = subprocess.run([exe, tiled, "--spp", str(nt), "--out", out2], check=True, capture_output=True, text=True, timeout=300)
    info2 = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert info2["tile_steps"] == nt * (256 // tile) ** 2
    img2 = np.fromfile(out2, np.float32).reshape(256, 256, 3)
    assert radiance_agreement(img2, gt["sppN"]) >= 0.999


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0]]                                     # two contexts on one device: everything but the NVLink hop, on any box
    if n >= 2:
        lists.append([0, 1])
    if n >= 8:
        lists.append(list(range(8)))
    return lists


def test_multi_gpu_renderer_behind_the_single_construction(cornell_scene, golden_dir):
    """CudaRenderer(scene, dir, devices): the ONE renderer Main.cpp:91 constructs drives several GPUs; Main.cpp's loop, the sample
    counter and the output buffers behave as on one GPU, and the image equals the 1-GPU image up to fp32 summation order."""
    n = 24
    r1 = lf.CudaRenderer(cornell_scene)
    r1.Run(n)
    single = r1.GetOutputBufferHDR()
    r1.close()
    for devs in _device_lists():
        r = lf.CudaRenderer(cornell_scene, devices=devs)
        steps = r.Run(n)
        assert steps == n and r.GetSampleCount() == n + 1
        img = r.GetOutputBufferHDR()
        np.testing.assert_allclose(img, single, rtol=2e-5, atol=1e-6)
        u8 = r.GetOutputBuffer()
        np.testing.assert_array_equal(u8, np.rint(np.clip(img, 0, 1) * 255).astype(np.uint8))
        # progressive: more samples after a read-out (the same call sequence on one GPU: a loop that is re-entered after its auto-stop
        # has already Update()d once, Main.cpp:197-206), then a camera move resets every device
        r.Run(n + 8)
        r1 = lf.CudaRenderer(cornell_scene); r1.Run(n); r1.GetOutputBufferHDR(); r1.Run(n + 8)
        np.testing.assert_allclose(r.GetOutputBufferHDR(), r1.GetOutputBufferHDR(), rtol=2e-5, atol=1e-6)
        r1.close()
        cornell_scene.set_camera_moving(True)
        r.Update(0.0); r.Render()
        assert r.GetSampleCount() == 1
        cornell_scene.set_camera_moving(False)
        r.Run(4)
        r1 = lf.CudaRenderer(cornell_scene); r1.Run(4)
        np.testing.assert_allclose(r.GetOutputBufferHDR(), r1.GetOutputBufferHDR(), rtol=2e-5, atol=1e-6)
        r1.close()
        r.close()


def test_multi_gpu_renderer_instance_edit(cornell_scene, golden_dir, tmp_path, oracle_lib):
    """The instance-edit path reaches every device of the group (lfcuda_group_update_instances)."""
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    m = pack.transforms.reshape(-1, 16)[3].copy()
    for devs in _device_lists():
        r = lf.CudaRenderer(cornell_scene, devices=devs)
        r.Run(2)
        moved = m.copy(); moved[13] += 0.1
        cornell_scene.move_instance(3, moved)
        r.Update(0.0); r.Render()
        r.Run(6)
        ref = _oracle_image(cornell_scene, tmp_path, "mg_moved", 6)
        np.testing.assert_allclose(r.GetOutputBufferHDR(), ref, rtol=2e-5, atol=1e-6)
        cornell_scene.move_instance(3, m)
        r.Update(0.0); r.Render()
        r.close()
