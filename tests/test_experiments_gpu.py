"""GPU tests of the experiment knobs that are OFF by default: an experiment may change speed, never a pixel.  LF_SORT_RAYS (bit 0: counting
sort of the extend queue, bit 1: of the shadow queue, by origin cell and direction octant, lf_kernels.h SortCtx) was measured in round 2 and
rejected (slower on every workload, profiles/README.md); the kernels stay behind the environment variable, and this test keeps them honest."""
import os
import subprocess
import sys

import numpy as np
import pytest

import lavaframe_b200 as lf

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RENDER = """
import sys, numpy as np
sys.path.insert(0, %r)
import lavaframe_b200 as lf
pt = lf.PathTracer(0)
pt.upload_pack(lf.ScenePack(sys.argv[1]))
pt.clear(); pt.render_frames(2, 3)
np.save(sys.argv[2], pt.read_accum())
pt.close()
""" % ROOT


@pytest.mark.parametrize("name", ["cornell", "c2mini", "c3mini", "c4gold"])
def test_sorted_rays_change_nothing(gpu, golden_dir, tmp_path, name):
    pack = os.path.join(golden_dir, f"{name}.lfpack")
    imgs = []
    for sort in ("0", "3"):
        out = str(tmp_path / f"{name}_{sort}.npy")
        subprocess.run([sys.executable, "-c", RENDER, pack, out], check=True, env=dict(os.environ, LF_SORT_RAYS=sort))
        imgs.append(np.load(out))
    assert np.array_equal(imgs[0], imgs[1]), f"{name}: {int((imgs[0] != imgs[1]).any(axis=2).sum())} pixels differ with LF_SORT_RAYS=3"
