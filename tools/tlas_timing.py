#!/usr/bin/env python
"""Development aid: time the instance-edit path on C4 (1296 instances): host TLAS upload + re-pack (lfcuda_update_instances, after the
reference's own Scene::RebuildInstances) against the device-side rebuild (lfcuda_update_instances_device)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lavaframe_b200 as lf  # noqa: E402
from scenes import gen_scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_stress"
path = gen_scenes.SCENES[name](os.path.join(ROOT, "scenes", "_gen", name))
s = lf.HostScene(path)
v, p, c = s.views()
p.width, p.height, p.tile_width, p.tile_height = 480, 270, 480, 270
pt = lf.PathTracer(0)
pt.upload_view(v, p, c)
n, top = v.num_instances, v.top_bvh_index
T = np.ctypeslib.as_array(v.transforms, shape=(n, 16)).copy()
M = np.ctypeslib.as_array(v.materials, shape=(v.num_materials, 28)).copy()
rng = np.random.RandomState(3)
th, td, tr = [], [], []
for it in range(12):
    idx = int(rng.randint(n))
    T[idx, 12:15] += rng.uniform(-1, 1, 3).astype(np.float32)
    t0 = time.perf_counter(); s.move_instance(idx, T[idx]); t1 = time.perf_counter()
    v2, _, _ = s.views()
    nodes = np.ctypeslib.as_array(v2.bvh_nodes, shape=(v2.num_nodes, 9))
    t2 = time.perf_counter(); pt.update_instances(T, M, nodes[top:], top); t3 = time.perf_counter()
    pt.update_instances_device(T, M); t4 = time.perf_counter()
    tr.append(t1 - t0); th.append(t3 - t2); td.append(t4 - t3)
    assert np.array_equal(pt.read_tlas_nodes(n).view(np.uint32), nodes[top:top + 2 * n - 1].view(np.uint32))
print(f"{name}: {n} instances; Scene::RebuildInstances on the host {np.median(tr) * 1e3:.2f} ms; lfcuda_update_instances (upload TLAS + host re-pack) "
      f"{np.median(th) * 1e3:.2f} ms; lfcuda_update_instances_device (device rebuild, matrices only) {np.median(td) * 1e3:.2f} ms; nodes equal on every edit")
pt.close(); s.close()
