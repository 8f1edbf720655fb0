"""The per-mesh inputs and expected outputs of the BLAS build, recovered from a scene pack (test helper).

A pack holds what Scene::CreateAccelerationStructures left (LavaFrame/Scene.cpp:180-231): the flattened nodes of every mesh's BVH, built by
RadeonRays::SplitBvh on the host (Mesh.cpp:93-111) and laid out by BvhTranslator::ProcessBLAS (bvh_translator.cpp:91-114), the triangle
references in BVH leaf order and the meshes' vertices.  From these: the triangle boxes Mesh::BuildBVH fed the builder, and the tree and index
order it produced - the ground truth for the device build and for its host-compiled text."""
import ctypes as C
import os

import numpy as np

REFBVH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "liblfrefbvh.so")


def have_reference_builder():
    return os.path.exists(REFBVH)


def reference_blas(prim_bounds):
    """The reference's own builder (RadeonRays::SplitBvh of the unchanged split_bvh.cpp, compiled into oracle/_ref/liblfrefbvh.so by
    oracle/Makefile) on the same boxes, in lfcuda_build_blas' layout: (boxes [k, 6], lr [k, 3] int32, indices [n], info).  Test infrastructure."""
    lib = C.CDLL(REFBVH)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    lib.lfref_build_blas.argtypes = [fp, C.c_int, fp, ip, ip]
    b = np.ascontiguousarray(prim_bounds, dtype=np.float32).reshape(-1, 6)
    n = b.shape[0]
    nodes = np.zeros(9 * max(2 * n - 1, 1), np.float32)
    idx = np.zeros(n, np.int32)
    info = np.zeros(3, np.int32)
    if lib.lfref_build_blas(b.ctypes.data_as(fp), n, nodes.ctypes.data_as(fp), idx.ctypes.data_as(ip), info.ctypes.data_as(ip)) != 0:
        raise ValueError("lfref_build_blas failed")
    a = nodes[:9 * int(info[0])].reshape(int(info[0]), 9)
    return a[:, :6].copy(), a[:, 6:].copy().view(np.int32), idx, {"num_nodes": int(info[0]), "num_indices": int(info[1]), "height": int(info[2])}


def pack_meshes(pack):
    nodes = pack.nodes.reshape(-1, 9)
    top = pack.top_index
    vi = pack.vert_indices.reshape(-1, 3)
    verts = pack.vertices.reshape(-1, 4)
    out = []
    root, tri_base = 0, 0
    while root < top and not (nodes[root] == 0).all():
        last = root
        while nodes[last, 8] == 0:                        # follow the right children down to the last leaf of the pre-order range
            last = int(nodes[last, 7])
        end = last + 1
        sub = nodes[root:end]
        leaf = sub[:, 8] == 1
        ntri = int(sub[leaf, 7].sum())
        v = verts[3 * tri_base:3 * (tri_base + ntri), :3].reshape(ntri, 3, 3)
        # Mesh::BuildBVH grows the box by v1, v2, v3 with std::min / std::max, which keep the first of +0 / -0 (Mesh.cpp:101-107)
        lo, hi = v[:, 0].copy(), v[:, 0].copy()
        for k in (1, 2):
            lo = np.where(v[:, k] < lo, v[:, k], lo)
            hi = np.where(hi < v[:, k], v[:, k], hi)
        bounds = np.concatenate([lo, hi], axis=1).astype(np.float32)
        idx = ((vi[tri_base:tri_base + ntri, 0] - 3 * tri_base) // 3).astype(np.int32)
        lr = np.zeros((end - root, 3), np.int32)
        lr[~leaf, 0] = sub[~leaf, 6].astype(np.int64) - root
        lr[~leaf, 1] = sub[~leaf, 7].astype(np.int64) - root
        lr[leaf, 0] = sub[leaf, 6].astype(np.int64) - tri_base
        lr[leaf, 1] = sub[leaf, 7].astype(np.int64)
        lr[leaf, 2] = 1
        out.append(dict(bounds=np.ascontiguousarray(bounds), boxes=np.ascontiguousarray(sub[:, :6]), lr=lr, indices=idx, root=root, tri_base=tri_base))
        root, tri_base = end, tri_base + ntri
    return out


def split_nodes(raw_nodes, num_nodes):
    """(boxes float32 [k,6], lr int32 [k,3]) from the builder's 9-word node records."""
    a = np.ascontiguousarray(raw_nodes[:9 * num_nodes]).reshape(num_nodes, 9)
    return a[:, :6].copy(), a[:, 6:].copy().view(np.int32)


def synthetic_cases():
    """(name, [n, 6] float32 triangle boxes): sizes around the leaf threshold, random soups, and the degenerate inputs that drive the builder's
    fallbacks (split_bvh.cpp:108 no centroid extent, :140-160 one side empty, :186-190 zero centroid box)."""
    rng = np.random.default_rng(20261018)

    def boxes(c, h):
        c = np.asarray(c, np.float32); h = np.asarray(h, np.float32)
        return np.concatenate([c - h, c + h], axis=1).astype(np.float32)

    cases = []
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 33, 64, 257, 1000):
        cases.append((f"soup{n}", boxes(rng.uniform(-5, 5, (n, 3)), rng.uniform(0.01, 0.5, (n, 3)))))
    cases.append(("soup20k", boxes(rng.normal(0, 3, (20000, 3)), rng.uniform(0.001, 0.2, (20000, 3)))))
    cases.append(("identical", boxes(np.tile([[1.0, 2.0, 3.0]], (37, 1)), np.tile([[0.5, 0.25, 0.125]], (37, 1)))))
    cases.append(("same_centre_other_size", boxes(np.tile([[0.5, 0.5, 0.5]], (50, 1)), rng.uniform(0.1, 2.0, (50, 3)))))
    line = np.zeros((100, 3), np.float32); line[:, 1] = np.linspace(-3, 3, 100)
    cases.append(("on_a_line", boxes(line, np.full((100, 3), 0.05))))
    plane = rng.uniform(-2, 2, (500, 3)); plane[:, 2] = 1.0
    cases.append(("in_a_plane", boxes(plane, np.full((500, 3), 0.02))))
    two = np.repeat(np.array([[-1.0, 0, 0], [1.0, 0, 0]], np.float32), [300, 5], axis=0)
    cases.append(("two_clusters", boxes(two, np.full((305, 3), 0.1))))
    dup = rng.uniform(-1, 1, (40, 3)); dup = np.repeat(dup, 9, axis=0)
    cases.append(("duplicates", boxes(dup, np.full((360, 3), 0.03))))
    cases.append(("large_coordinates", boxes(rng.uniform(-1e6, 1e6, (3000, 3)), rng.uniform(1, 1e4, (3000, 3)))))
    cases.append(("tiny_extent", boxes(1.0 + rng.uniform(0, 1e-6, (300, 3)), np.full((300, 3), 1e-7))))
    grid = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(8), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    cases.append(("regular_grid", boxes(grid, np.full((len(grid), 3), 0.5))))
    return cases


def signed_zero_cases():
    """Boxes on a coarse integer grid whose zero coordinates carry random signs: the builder's decisions never look at the sign of a zero, but the
    boxes it writes out do (std::min / std::max keep the first of +0 / -0 they meet, in the order the reference grows a box)."""
    rng = np.random.default_rng(77)
    cases = []
    for n, span in ((5, 1), (9, 1), (40, 2), (300, 2), (300, 1), (2500, 3), (20000, 4), (4000, 0)):
        for rep in range(3):
            lo = rng.integers(-span, span + 1, (n, 3)).astype(np.float32)
            hi = lo + rng.integers(0, 2, (n, 3)).astype(np.float32)
            b = np.concatenate([lo, hi], axis=1).astype(np.float32)
            z = b == 0
            b[z & (rng.random(b.shape) < 0.5)] = -0.0
            cases.append((f"signed_zero_n{n}_span{span}_{rep}", b))
    return cases


def negative_zero_scene(outdir):
    """c2_mini with a band of its bumpy sphere collapsed onto the coordinate planes, written the way OBJ exporters write such vertices:
    `-0.0` for about half of them, `0.0` for the rest.  Returns the .scene path."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from scenes import gen_scenes
    scene = gen_scenes.c2_mini(str(outdir))
    obj = os.path.join(os.path.dirname(scene), "c2mini_mesh.obj")
    rng = np.random.default_rng(3)
    out, snapped = [], 0
    for line in open(obj):
        if line.startswith("v "):
            y = []
            for c in (float(t) for t in line.split()[1:4]):
                if abs(c) < 0.12:
                    c = -0.0 if rng.random() < 0.5 else 0.0
                    snapped += 1
                y.append(c)
            line = "v %r %r %r\n" % tuple(y)
        out.append(line)
    open(obj, "w").write("".join(out))
    assert snapped > 100
    return scene
