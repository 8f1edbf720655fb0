"""GPU tests (-m gpu) of lfcuda_build_blas (csrc/lf_blas.cu): a mesh's BVH built on the device equals the reference's host build NODE FOR NODE -
same boxes bit for bit, same child and leaf records, same triangle order.  Ground truth as in test_blas_build.py: the trees
inside the scene packs (built by the reference's unchanged SplitBvh), the reference's builder run on synthetic / degenerate boxes, and - at
BASELINE's full size - C2's 869 880-triangle mesh; then the drop-in route: a scene loaded by the reference's unchanged loader with
LF_DEVICE_BLAS=1 yields a byte-identical pack, and renders identically."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lavaframe_b200 as lf
from lavaframe_b200.capi import lib_path
from blas_cases import pack_meshes, synthetic_cases, signed_zero_cases, reference_blas, have_reference_builder, negative_zero_scene

pytestmark = pytest.mark.gpu


def assert_same_tree(got, want_boxes, want_lr, want_idx, what):
    boxes, lr, idx, info = got
    assert info["num_nodes"] == len(want_boxes), what
    assert boxes.tobytes() == np.ascontiguousarray(want_boxes).tobytes(), f"{what}: boxes differ (bitwise)"
    assert (lr == want_lr).all(), f"{what}: child / leaf records differ"
    assert (idx == want_idx).all(), f"{what}: triangle order differs"
    assert info["num_indices"] == len(want_idx) and info["levels"] == info["height"] + 1 and info["launches"] > 0


@pytest.mark.parametrize("pack_name", ["cornell", "c2mini", "c3mini", "c4gold"])
def test_device_build_equals_the_packs_trees(gpu, golden_dir, pack_name):
    pack = lf.ScenePack(os.path.join(golden_dir, f"{pack_name}.lfpack"))
    for k, m in enumerate(pack_meshes(pack)):
        assert_same_tree(lf.build_blas(m["bounds"], gpu), m["boxes"], m["lr"], m["indices"], f"{pack_name} mesh {k}")


def test_device_build_against_the_reference_builder(gpu):
    if not have_reference_builder():
        pytest.fail("oracle/_ref/liblfrefbvh.so missing: __graft_entry__.build() must run where /root/reference exists")
    for name, b in synthetic_cases():
        rb, rl, ri, rinfo = reference_blas(b)
        got = lf.build_blas(b, gpu)
        assert_same_tree(got, rb, rl, ri, name)
        assert got[3]["height"] == rinfo["height"], name


def test_device_build_sign_of_zero_planes(gpu):
    """+0 / -0 coordinates: the sign of every zero box plane is the one the reference's growth order gives it (lf_blas_build.h acc_zero)."""
    if not have_reference_builder():
        pytest.fail("oracle/_ref/liblfrefbvh.so missing: __graft_entry__.build() must run where /root/reference exists")
    for name, b in signed_zero_cases():
        rb, rl, ri, rinfo = reference_blas(b)
        got = lf.build_blas(b, gpu)
        assert_same_tree(got, rb, rl, ri, name)
        assert got[3]["negative_zero"] == int((np.signbit(b) & (b == 0)).any())


def test_device_build_is_repeatable_and_reports_negative_zero(gpu):
    rng = np.random.default_rng(5)
    c = rng.uniform(-1, 1, (5000, 3)).astype(np.float32)
    b = np.concatenate([c - 0.01, c + 0.01], axis=1).astype(np.float32)
    a1 = lf.build_blas(b, gpu); a2 = lf.build_blas(b, gpu)
    assert a1[0].tobytes() == a2[0].tobytes() and (a1[1] == a2[1]).all() and (a1[2] == a2[2]).all()      # atomics do not make it vary
    assert a1[3]["negative_zero"] == 0
    b[17, 1] = -0.0
    assert lf.build_blas(b, gpu)[3]["negative_zero"] == 1


def test_bad_arguments(gpu):
    b = np.zeros((4, 6), np.float32)
    with pytest.raises(lf.LfCudaError, match="num_bins"):
        lf.build_blas(b, gpu, num_bins=65)
    with pytest.raises(lf.LfCudaError, match="device"):
        lf.build_blas(b, 99)
    with pytest.raises(lf.LfCudaError, match="num_prims"):
        lf.build_blas(np.zeros((0, 6), np.float32), gpu)


@pytest.fixture(scope="module")
def c2_scene(gpu, tmp_path_factory):
    if not os.path.exists(lib_path("liblfhost.so")):
        pytest.fail("liblfhost.so missing: __graft_entry__.build() must run where /root/reference exists")
    from scenes import gen_scenes
    out = str(tmp_path_factory.mktemp("c2_full_blas"))
    scene = gen_scenes.c2_full(out)
    return scene, out


def run_scenepack(scene, pack, env_extra):
    exe = os.path.join(os.path.dirname(lib_path("liblfcuda.so")), "bin", "lf_scenepack")
    r = subprocess.run([exe, scene, pack], capture_output=True, text=True, env=dict(os.environ, **env_extra))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_full_size_drop_in_pack_is_byte_identical(gpu, c2_scene, tmp_path):
    """C2 (869 880 triangles) through the reference's unchanged loader, once with the host build and once with LF_DEVICE_BLAS=1
    (Mesh.h:18's builder constructed as LfDeviceSplitBvh, host/DeviceBvh.h): the flattened scene must not differ by a byte."""
    scene, out = c2_scene
    host_pack, dev_pack = str(tmp_path / "host.lfpack"), str(tmp_path / "dev.lfpack")
    log_h = run_scenepack(scene, host_pack, {"LF_DEVICE_BLAS": "0"})
    log_d = run_scenepack(scene, dev_pack, {"LF_DEVICE_BLAS": "1", "LF_DEVICE_BLAS_MIN": "1"})
    print(log_h.splitlines()[-1]); print(log_d.splitlines()[-1])
    assert "mesh BVH builds: 0 on the device" in log_h
    assert "mesh BVH builds: 2 on the device" in log_d and " 0 on the host" in log_d
    assert open(host_pack, "rb").read() == open(dev_pack, "rb").read()
    # and directly through the C ABI, with the timing the docs quote
    pack = lf.ScenePack(host_pack)
    m = max(pack_meshes(pack), key=lambda q: len(q["bounds"]))
    assert len(m["bounds"]) == 869880
    got = lf.build_blas(m["bounds"], gpu)
    assert_same_tree(got, m["boxes"], m["lr"], m["indices"], "c2_full mesh")
    got = lf.build_blas(m["bounds"], gpu)       # second call: no first-use costs
    print("c2_full device BLAS build:", got[3])
    assert got[3]["build_ms"] < 2000.0


def test_renderer_on_a_device_built_scene(gpu, tmp_path):
    """HostScene (the reference's loader) with the device build switched on through the C API, rendered by CudaRenderer: same image."""
    from scenes import gen_scenes
    lib = lf.load_lfhost()
    scene_path = gen_scenes.c2_mini(str(tmp_path))
    imgs = []
    for on in (0, 1):
        lib.lfhost_set_device_blas(on, 1, gpu)
        s = lf.HostScene(scene_path)
        stats = np.zeros(8)
        lib.lfhost_blas_stats(stats.ctypes.data_as(C.POINTER(C.c_double)), 1)
        assert (stats[0] > 0) == bool(on) and (stats[1] > 0) != bool(on)
        r = lf.CudaRenderer(s, gpu)
        r.Run(2)
        imgs.append(r.GetOutputBufferHDR().copy())
        r.close(); s.close()
    lib.lfhost_set_device_blas(0, -1, -1)
    assert imgs[0].tobytes() == imgs[1].tobytes()


def test_drop_in_pack_with_negative_zero_coordinates(gpu, tmp_path):
    """An OBJ the way exporters write them - `-0.000000` for what rounds to zero, next to `0.000000`: the device-built scene is still the
    host-built scene byte for byte (the sign of zero box planes follows the reference's growth order, lf_blas_build.h acc_zero)."""
    import re
    if not os.path.exists(lib_path("liblfhost.so")):
        pytest.fail("liblfhost.so missing: __graft_entry__.build() must run where /root/reference exists")
    scene = negative_zero_scene(tmp_path)
    host_pack, dev_pack = str(tmp_path / "host.lfpack"), str(tmp_path / "dev.lfpack")
    run_scenepack(scene, host_pack, {"LF_DEVICE_BLAS": "0"})
    log_d = run_scenepack(scene, dev_pack, {"LF_DEVICE_BLAS": "1", "LF_DEVICE_BLAS_MIN": "1"})
    m = re.search(r"(\d+) with -0.0 bounds", log_d)
    assert m and int(m.group(1)) >= 1, log_d
    a, b = open(host_pack, "rb").read(), open(dev_pack, "rb").read()
    assert a == b
    nodes = lf.ScenePack(host_pack).nodes.reshape(-1, 9)[:, :6]
    assert int((np.signbit(nodes) & (nodes == 0)).sum()) > 0          # the reference's own node boxes do contain -0.0 planes here
