"""GPU tests (-m gpu): the held-out stress scenes of tools/stress_parity.py through the CUDA path.

tools/stress_parity.py generates scenes the restatement was never fitted on (30 non-uniformly scaled bumpy spheres with RANDOM
Disney parameters, all four texture kinds incl. emission textures (pathtrace.glsl:112-113), HDR env + quad + sphere light, thin
lens, depth 3 / 5 / 8; seeds >= 100 also vary the shader's #defines: lights / env on or off, Russian roulette off or from
depth 0 / 1 / 3, constant background) and showed the ORACLE bit-identical to the unmodified reference on llvmpipe on all 18 seeds
(profiles/r1_stress_parity.txt; CPU only, authoring container).  Here the same 18 scenes go through liblfcuda (wavefront and
megakernel) and must equal the oracle bit for bit, which closes the chain reference == oracle == CUDA on them.
The scenes are flattened on the spot by the reference's unchanged loader + BVH builder (lavaframe_b200/bin/lf_scenepack)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle_api import Oracle
import lavaframe_b200 as lf

pytestmark = pytest.mark.gpu

SEEDS = list(range(1, 13)) + list(range(100, 106))


@pytest.fixture(scope="module")
def tracer(gpu):
    pt = lf.PathTracer(gpu)
    yield pt
    pt.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_stress_seed_vs_oracle(tracer, oracle_lib, tmp_path, seed):
    import stress_parity as sp
    if not os.path.exists(sp.PACKBIN):
        pytest.fail("lavaframe_b200/bin/lf_scenepack missing: __graft_entry__.build() must run where /root/reference exists")
    scene = sp.scene(str(tmp_path / "assets"), seed)
    pack_path = str(tmp_path / "s.lfpack")
    subprocess.run([sp.PACKBIN, scene, pack_path], check=True, capture_output=True)
    var = sp.variant(seed)
    pack = lf.ScenePack(pack_path)
    o = Oracle(pack_path)
    over = {}
    if var["bg"]:
        over = dict(use_constant_bg=1)
    tracer.upload_pack(pack, **over)
    if var["bg"]:
        for q in (tracer.params, o.params):
            q.use_constant_bg = 1
            q.bg_color[0], q.bg_color[1], q.bg_color[2] = var["bg"]
        tracer.set_params(tracer.params)
        o.update_params()
    hits, ohits = tracer.primary_hits(2), o.primary_hits(2)
    for a, b in zip(hits, ohits):
        assert np.array_equal(a, b)
    ref1 = o.render_frames(2, 1)
    ref4 = o.render_frames(2, 4)
    o.close()
    assert ref1.any() and np.isfinite(ref4).all()
    for mode in (0, 1):
        tracer.update_params(kernel_mode=mode)
        for n, ref in ((1, ref1), (4, ref4)):
            tracer.clear(); tracer.render_frames(2, n)
            img = tracer.read_accum()
            differ = int((img != ref).any(axis=2).sum())
            assert differ == 0, f"seed {seed} ({var}), mode {mode}, {n} spp: {differ} pixels differ from the oracle"
