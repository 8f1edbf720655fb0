"""CPU tests: the C-ABI library loads and exports every symbol include/lfcuda.h declares; no compute calls."""
import os
import re

import pytest

import lavaframe_b200 as lf
from lavaframe_b200.capi import LFCUDA_SYMBOLS, lib_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lfcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lfcuda_[a-z0-9_]+)\s*\(", text)) - {"lfcuda_ctx"})


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(LFCUDA_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lib_path("liblfcuda.so")), "liblfcuda.so not built: run __graft_entry__.build()"
    lib = lf.load_lfcuda()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.lfcuda_abi_version() == 2


def test_struct_sizes_match_the_header():
    import ctypes as C
    assert C.sizeof(lf.LfParams) == 17 * 4
    assert C.sizeof(lf.LfCamera) == 15 * 4
    assert C.sizeof(lf.LfCounters) == 17 * 8
    assert C.sizeof(lf.LfStageStats) == 7 * 8 + 7 * 8
    assert C.sizeof(lf.LfSceneView) == 160   # checked against gcc sizeof


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lf.LfCudaError, match="no CUDA device"):
        lf.PathTracer(0)


def test_host_library_loads_reference_scene(golden_dir, tmp_path):
    """liblfhost.so = the reference's unchanged loader/BVH code + CudaRenderer; loading needs no GPU."""
    if not os.path.exists(lib_path("liblfhost.so")):
        pytest.skip("liblfhost.so not built (needs /root/reference at build time)")
    from scenes import gen_scenes
    try:
        scene_path = gen_scenes.cornell_256(str(tmp_path))
    except Exception as e:  # no reference assets available
        pytest.skip(str(e))
    s = lf.HostScene(scene_path)
    v, p, c = s.views()
    assert (v.num_nodes, v.top_bvh_index, v.num_tri_refs, p.width, p.max_depth) == (43, 29, 36, 256, 4)
    out = tmp_path / "c.lfpack"
    s.write_pack(out)
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "cornell.lfpack"), "rb").read()
    s.close()
