"""Hand-built .lfpack scenes for edge-case tests (test infrastructure).

The arrays follow SURVEY.md Appendix A / lavaframe_b200/host/scenepack.h exactly as the reference's
Scene::CreateAccelerationStructures + BvhTranslator would lay them out (bvh_translator.cpp:35-141), but the BVH here is built
by hand so that shapes the reference's SAH builder never produces (a 40-level chain) can be walked by both the CUDA path and
the CPU oracle.
"""
import numpy as np

MAGIC = b"LFPACK01"


def write_pack(path, nodes, vert_indices, vertices, normals, transforms, materials, lights, top_index, width, height,
               max_depth=3, enable_rr=1, rr_depth=2, use_const_bg=0, bg=(0.0, 0.0, 0.0), cam_pos=(0, 0, 0), cam_right=(1, 0, 0), cam_up=(0, 1, 0),
               cam_fwd=(0, 0, 1), fov=0.6, focal=1.0, aperture=0.0):
    nodes = np.ascontiguousarray(nodes, np.float32).reshape(-1, 9)
    vert_indices = np.ascontiguousarray(vert_indices, np.int32).reshape(-1, 3)
    vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 4)
    normals = np.ascontiguousarray(normals, np.float32).reshape(-1, 4)
    transforms = np.ascontiguousarray(transforms, np.float32).reshape(-1, 16)
    materials = np.ascontiguousarray(materials, np.float32).reshape(-1, 28)
    lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 15)
    ih = np.zeros(32, np.int32)
    ih[0:7] = [len(nodes), top_index, len(vert_indices), len(vertices), len(transforms), len(materials), len(lights)]
    ih[12:21] = [width, height, width, height, max_depth, enable_rr, rr_depth, 0, use_const_bg]
    fh = np.zeros(32, np.float32)
    fh[0:3] = bg
    fh[3] = 1.0
    fh[4:7] = cam_pos; fh[7:10] = cam_right; fh[10:13] = cam_up; fh[13:16] = cam_fwd
    fh[16:19] = [fov, focal, aperture]
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(ih.tobytes()); f.write(fh.tobytes())
        for a in (nodes, vert_indices, vertices, normals, transforms, materials, lights):
            f.write(a.tobytes())
    return path


def material(albedo=(0.8, 0.8, 0.8), emission=(0, 0, 0), metallic=0.0, roughness=0.5, transmission=0.0, ior=1.45):
    m = np.zeros(28, np.float32)
    m[0:3] = albedo; m[3] = 0.5                       # albedo, specular (Material.h:18-45 defaults)
    m[4:7] = emission
    m[8] = metallic; m[9] = roughness
    m[16] = transmission; m[17] = ior; m[18] = 1.0    # transmission, ior, atDistance
    m[20:23] = 1.0                                    # extinction
    m[24:28] = -1.0                                   # texture ids: none
    return m


def quad_light(pos, u, v, emission):
    L = np.zeros(15, np.float32)
    L[0:3] = pos; L[3:6] = emission; L[6:9] = u; L[9:12] = v
    L[12] = 0.0; L[13] = float(np.linalg.norm(np.cross(u, v))); L[14] = 0.0      # radius, area, type quad (Loader.cpp:184-195)
    return L


def distant_light(direction_pos, emission):
    """type 2 (globals.glsl:13 DISTANT_LIGHT): sampleDistantLight (sampling.glsl:208-216) shines from normalize(position); the
    scene grammar has no keyword for it (Loader.cpp:184-195 writes quads and spheres only), the shader branch exists all the same."""
    L = np.zeros(15, np.float32)
    L[0:3] = direction_pos; L[3:6] = emission
    L[13] = 0.0; L[14] = 2.0                          # area 0: the MIS weight stays 1 (pathtrace.glsl:190-191)
    return L


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[3, 0:3] = (x, y, z)                             # translation in data[3][0..2] (Mat4.h:35-51)
    return m.reshape(16)


def chain_scene(path, n_tris=40, width=64, height=48, instances=((0.0, 0.0, 0.0), (3.0, 0.0, 0.0)), lights=None):
    """`n_tris` triangles stacked along z, one BLAS that is a CHAIN: inner node i = (leaf i, inner i+1).  Seen from the far
    end every inner node's right child is the nearer one, so the walk defers one leaf per level: the traversal stack gets
    about n_tris deep (the reference allows 64, closest_hit.glsl:70).  Two instances under a two-leaf TLAS, one quad light."""
    verts, norms, idx = [], [], []
    for k in range(n_tris):
        z = float(k)
        s = 1.0 + 0.01 * k
        verts += [(-s, -s, z, 0.0), (s, -s, z, 1.0), (0.0, s, z, 0.5)]
        norms += [(0, 0, 1, 0.0), (0, 0, 1, 0.0), (0, 0, 1, 1.0)]
        idx.append((3 * k, 3 * k + 1, 3 * k + 2))
    verts = np.array(verts, np.float32)
    lo = verts.reshape(n_tris, 3, 4)[:, :, :3].min(axis=1)
    hi = verts.reshape(n_tris, 3, 4)[:, :, :3].max(axis=1)
    nodes = []
    # preorder: inner_i at 2i, leaf_i at 2i+1, ..., last inner's right child is the last leaf
    n_inner = n_tris - 1
    for i in range(n_inner):
        blo, bhi = lo[i:].min(axis=0), hi[i:].max(axis=0)
        left = 2 * i + 1
        right = 2 * i + 2                                        # the next inner node, or the last leaf
        nodes.append((*blo, *bhi, left, right, 0))
        nodes.append((*lo[i], *hi[i], i, 1, 1))                  # BLAS leaf: (firstTriRef, numprims, 1)
    nodes.append((*lo[n_tris - 1], *hi[n_tris - 1], n_tris - 1, 1, 1))
    blas_lo, blas_hi = lo.min(axis=0), hi.max(axis=0)
    top = len(nodes)
    ninst = len(instances)
    assert ninst in (1, 2)
    boxes = [(blas_lo + np.array(t, np.float32), blas_hi + np.array(t, np.float32)) for t in instances]
    if ninst == 2:
        nodes.append((*np.minimum(boxes[0][0], boxes[1][0]), *np.maximum(boxes[0][1], boxes[1][1]), top + 1, top + 2, 0))
    for k in range(ninst):                                           # (one instance: the TLAS root is its leaf)
        nodes.append((*boxes[k][0], *boxes[k][1], 0, 1 + k, -(k + 1)))   # TLAS leaf: (blasRoot, materialID, -(instance+1))
    nodes.append((0,) * 9)                                           # 2N reserved slots, 2N-1 used (bvh_translator.cpp:95-101)
    mats = [material(), material(albedo=(0.9, 0.3, 0.2)), material(albedo=(0.2, 0.4, 0.9), metallic=1.0, roughness=0.3)]
    zc = n_tris + 6.0
    if lights is None:
        lights = [quad_light((-1.0, -1.0, zc), (0.0, 3.0, 0.0), (6.0, 0.0, 0.0), (20.0, 20.0, 20.0))]   # behind the camera, normal -z
    return write_pack(path, nodes, idx, verts, norms, [translate(*t) for t in instances], mats, lights, top, width, height,
                      max_depth=3, cam_pos=(1.5, 0.2, n_tris + 4.0), cam_right=(1, 0, 0), cam_up=(0, 1, 0), cam_fwd=(0, 0, -1), fov=0.9)


def axis_ray_scene(path, width=8, height=8, zc=0.5):
    """KAT for the NaN semantics of the slab test (intersection.glsl:53-67).  cam_right = 0, so every
    primary ray is d = normalize(dy * up + forward) with d.z == 0 EXACTLY and origin.z == zc.  The BLAS root's LEFT child box is
    flat in z at zc (hand-built, deliberately not tight: its triangle stands across the ray at x = 5), so its z slabs are
    (zc - zc) * (1 / 0) = 0 * inf = NaN.  With GLSL's min / max as llvmpipe evaluates them (MINPS / MAXPS return the second
    operand on NaN) t1 is NaN and the box is MISSED: the reference never tests the near triangle and the ray reaches the
    RIGHT child's triangle at x = 10.  A min / max that drops NaNs (IEEE fminf / the GPU's FMNMX) hits the flat box and returns
    the near triangle instead."""
    verts = np.array([(5, -4, zc - 2, 0), (5, 4, zc - 2, 1), (5, 0, zc + 3, 0.5),           # near triangle, plane x = 5
                      (10, -8, zc - 4, 0), (10, 8, zc - 4, 1), (10, 0, zc + 6, 0.5)], np.float32)   # far triangle, plane x = 10
    norms = np.array([(-1, 0, 0, 0.0)] * 6, np.float32)
    idx = [(0, 1, 2), (3, 4, 5)]
    near_lo, near_hi = (4.0, -4.0, zc), (6.0, 4.0, zc)            # flat in z, on the rays' plane
    far_lo, far_hi = (9.0, -8.0, zc - 4), (11.0, 8.0, zc + 6)
    blo, bhi = np.minimum(near_lo, far_lo), np.maximum(near_hi, far_hi)
    nodes = [(*blo, *bhi, 1, 2, 0),
             (*near_lo, *near_hi, 0, 1, 1),                       # BLAS leaf: (firstTriRef, numprims, 1)
             (*far_lo, *far_hi, 1, 1, 1)]
    top = len(nodes)
    inst = [(0.0, 0.0, 0.0), (0.0, 40.0, 0.0)]                    # the second instance is out of view
    boxes = [(blo + np.array(t, np.float32), bhi + np.array(t, np.float32)) for t in inst]
    nodes.append((*np.minimum(boxes[0][0], boxes[1][0]), *np.maximum(boxes[0][1], boxes[1][1]), top + 1, top + 2, 0))
    for k in range(2):
        nodes.append((*boxes[k][0], *boxes[k][1], 0, 1 + k, -(k + 1)))
    nodes.append((0,) * 9)
    mats = [material(), material(albedo=(0.9, 0.3, 0.2)), material(albedo=(0.2, 0.4, 0.9))]
    # a quad light behind the camera facing +x: the direct light at the hit falls off with its distance, so the image tells x = 5 from x = 10
    lights = [quad_light((-1.0, -1.5, zc - 1.5), (0.0, 3.0, 0.0), (0.0, 0.0, 3.0), (30.0, 20.0, 10.0))]
    return write_pack(path, nodes, idx, verts, norms, [translate(*t) for t in inst], mats, lights, top, width, height,
                      max_depth=2, cam_pos=(0.0, 0.0, zc), cam_right=(0, 0, 0), cam_up=(0, 1, 0), cam_fwd=(1, 0, 0), fov=0.5)
