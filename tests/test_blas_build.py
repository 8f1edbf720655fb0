"""CPU tests of the device BLAS builder's TEXT (lavaframe_b200/csrc/lf_blas_build.h), compiled for the host by tests/hostcheck: the
level-synchronous data-parallel restatement of RadeonRays::SplitBvh (Mesh.h:18, split_bvh.cpp:11-289) must produce the reference builder's
tree node for node, whatever the order in which the work items of a step run.  Ground truth: (1) the node arrays and triangle order inside
the committed scene packs, which the reference's unchanged builder produced (tests/golden/*.lfpack), and (2) the reference's
builder itself (the unchanged split_bvh.cpp, compiled into oracle/_ref/liblfrefbvh.so) run on synthetic and degenerate inputs.  The CUDA execution of the same text
is tested in test_blas_device_gpu.py; nothing here is a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lavaframe_b200 as lf
from lavaframe_b200.capi import lib_path
from blas_cases import pack_meshes, split_nodes, synthetic_cases, signed_zero_cases, reference_blas, have_reference_builder, negative_zero_scene

HERE = os.path.dirname(os.path.abspath(__file__))
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def hostbuild():
    so = os.path.join(HERE, "hostcheck", "liblfblascheck.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(HERE, "hostcheck"), "liblfblascheck.so"], check=True)
    lib = C.CDLL(so)
    lib.lfblascheck_build.argtypes = [fp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, fp, ip, ip]

    def run(bounds, order=0, bin_cap=4096, tc=2.0, bins=64):
        b = np.ascontiguousarray(bounds, np.float32)
        n = len(b)
        out = np.zeros(9 * 2 * n, np.float32); idx = np.zeros(n, np.int32); info = np.zeros(4, np.int32)
        assert lib.lfblascheck_build(b.ctypes.data_as(fp), n, tc, bins, order, bin_cap, out.ctypes.data_as(fp), idx.ctypes.data_as(ip), info.ctypes.data_as(ip)) == 0
        boxes, lr = split_nodes(out, int(info[0]))
        return boxes, lr, idx, dict(num_nodes=int(info[0]), height=int(info[1]), negative_zero=int(info[2]), levels=int(info[3]))
    return run


def assert_same_tree(got, want_boxes, want_lr, want_idx, what):
    boxes, lr, idx, info = got
    assert info["num_nodes"] == len(want_boxes), what
    assert boxes.tobytes() == np.ascontiguousarray(want_boxes).tobytes(), f"{what}: boxes differ (bitwise)"
    assert (lr == want_lr).all(), f"{what}: child / leaf records differ"
    assert (idx == want_idx).all(), f"{what}: triangle order differs"


@pytest.mark.parametrize("pack_name", ["cornell", "c2mini", "c3mini", "c4gold"])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_text_rebuilds_the_packs_trees(hostbuild, golden_dir, pack_name, order):
    pack = lf.ScenePack(os.path.join(golden_dir, f"{pack_name}.lfpack"))
    meshes = pack_meshes(pack)
    assert meshes and sum(len(m["boxes"]) for m in meshes) == pack.top_index      # every BLAS node of the pack is covered
    for k, m in enumerate(meshes):
        got = hostbuild(m["bounds"], order=order, bin_cap=(3 if order == 1 else 4096))        # 3: the bins of a level go in many batches
        assert_same_tree(got, m["boxes"], m["lr"], m["indices"], f"{pack_name} mesh {k} order {order}")
        assert got[3]["negative_zero"] == 0


@pytest.mark.parametrize("order", [0, 1, 2])
def test_text_against_the_reference_builder(hostbuild, order):
    if not have_reference_builder():
        pytest.skip("oracle/_ref/liblfrefbvh.so not built (make -C oracle ref needs /root/reference)")
    for name, b in synthetic_cases():
        rb, rl, ri, rinfo = reference_blas(b)
        got = hostbuild(b, order=order, bin_cap=(5 if order == 2 else 4096))
        assert_same_tree(got, rb, rl, ri, f"{name} order {order}")
        assert got[3]["height"] == rinfo["height"], name


@pytest.mark.parametrize("order", [0, 1, 2])
def test_sign_of_zero_planes(hostbuild, order):
    """Boxes full of +0 / -0 coordinates: the node boxes carry the sign the reference's growth order gives them, bit for bit, in any execution order."""
    if not have_reference_builder():
        pytest.skip("oracle/_ref/liblfrefbvh.so not built (make -C oracle ref needs /root/reference)")
    planes = 0
    for name, b in signed_zero_cases():
        rb, rl, ri, rinfo = reference_blas(b)
        got = hostbuild(b, order=order, bin_cap=64)
        assert_same_tree(got, rb, rl, ri, f"{name} order {order}")
        planes += int((np.signbit(rb) & (rb == 0)).sum())
    assert planes > 10000          # the cases do exercise it: that many box planes of the reference's trees are -0.0


def test_loader_mesh_with_negative_zero_vertices(hostbuild, tmp_path):
    """A mesh whose OBJ holds `-0.0` and `0.0` vertices, loaded and built by the reference's unchanged loader and builder (lf_scenepack): the
    text rebuilds its tree with the same sign on every zero box plane."""
    exe = os.path.join(os.path.dirname(lib_path("liblfcuda.so")), "bin", "lf_scenepack")
    if not os.path.exists(exe):
        pytest.skip("lf_scenepack not built (needs /root/reference at build time)")
    try:
        scene = negative_zero_scene(tmp_path)
    except Exception as e:   # no reference assets available
        pytest.skip(str(e))
    pack_path = str(tmp_path / "nz.lfpack")
    subprocess.run([exe, scene, pack_path], check=True, capture_output=True, env=dict(os.environ, LF_DEVICE_BLAS="0"))
    pack = lf.ScenePack(pack_path)
    nodes = pack.nodes.reshape(-1, 9)[:pack.top_index, :6]
    assert int((np.signbit(nodes) & (nodes == 0)).sum()) > 500
    for order in (0, 1, 2):
        for k, m in enumerate(pack_meshes(pack)):
            got = hostbuild(m["bounds"], order=order)
            assert_same_tree(got, m["boxes"], m["lr"], m["indices"], f"mesh {k} order {order}")


def test_negative_zero_is_reported(hostbuild):
    b = np.array([[0.0, 0, 0, 1, 1, 1], [-0.0, 1, 1, 2, 2, 2], [1, 1, 1, 3, 3, 3], [2, 0, 0, 3, 1, 1]], np.float32)
    assert hostbuild(b)[3]["negative_zero"] == 1
    b[1, 0] = 0.0
    assert hostbuild(b)[3]["negative_zero"] == 0


def test_leaf_order_is_right_child_first(hostbuild):
    """The reference builds the right child first, so leaves are numbered from the END of the partitioned array (split_bvh.cpp:162-167):
    the first leaf in pre-order holds the last indices."""
    b = np.zeros((8, 6), np.float32)
    b[:, 0] = np.arange(8); b[:, 3] = np.arange(8) + 0.5; b[:, 4:] = 0.5
    boxes, lr, idx, info = hostbuild(b)
    leaves = lr[lr[:, 2] == 1]
    assert sorted(idx.tolist()) == list(range(8)) and leaves[:, 1].sum() == 8
    assert leaves[0, 0] + leaves[0, 1] == 8          # pre-order's first leaf = the array's head = the last indices handed out
    assert (np.diff(leaves[:, 0]) < 0).all()


def test_requested_device_build_fails_loudly_without_a_gpu(tmp_path):
    """LF_DEVICE_BLAS=1 on a machine without a CUDA device: the loader reports the failure - it does not quietly fall back to the host build."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = os.path.join(os.path.dirname(lib_path("liblfcuda.so")), "bin", "lf_scenepack")
    if not os.path.exists(exe):
        pytest.skip("lf_scenepack not built (needs /root/reference at build time)")
    from scenes import gen_scenes
    try:
        scene = gen_scenes.cornell_256(str(tmp_path))
    except Exception as e:   # no reference assets available
        pytest.skip(str(e))
    r = subprocess.run([exe, scene, str(tmp_path / "x.lfpack")], capture_output=True, text=True, env=dict(os.environ, LF_DEVICE_BLAS="1", LF_DEVICE_BLAS_MIN="1"))
    assert r.returncode != 0 and "no CUDA device" in r.stderr and not os.path.exists(tmp_path / "x.lfpack")
    with pytest.raises(lf.LfCudaError, match="no CUDA device"):
        lf.build_blas(np.zeros((4, 6), np.float32), 0)
