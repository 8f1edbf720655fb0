/*
 * lfcuda.h — C ABI of the B200-native path-tracing core (liblfcuda.so).
 *
 * This is the drop-in boundary for ONE path of LavaFrame: the per-pixel progressive
 * path-tracing loop that the reference runs as a GLSL fragment shader
 * (shaders/renderer.glsl:25-69 + the files of shaders/common/), launched by
 * TiledRenderer::Render (LavaFrame/TiledRenderer.cpp:336-339).
 *
 * Everything here is `extern "C"`, plain pointers and sizes.  All pointer arguments are
 * caller-owned HOST memory unless a parameter is explicitly called a device pointer; host
 * arrays are copied during the call.  Every function returns 0 on success and a negative
 * LFCUDA_E* code on failure; lfcuda_last_error() gives the text.  A context is bound to one
 * CUDA device and is not thread-safe (the reference's Renderer is single-threaded too,
 * LavaFrame/Main.cpp:313-755).
 *
 * Array layouts are exactly the ones the reference uploads as GL textures in
 * Renderer::Init (LavaFrame/Renderer.cpp:76-188) -- the element layouts are listed beside each
 * field.  Nothing is re-ordered by the caller: repacking for the GPU happens inside
 * lfcuda_upload_scene.
 */
#ifndef LFCUDA_H
#define LFCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFCUDA_ABI_VERSION 2

enum {
    LFCUDA_OK = 0,
    LFCUDA_EINVAL = -1,   /* bad argument / scene not uploaded / params not set */
    LFCUDA_ECUDA = -2,    /* a CUDA runtime call failed (text in last_error) */
    LFCUDA_ENOMEM = -3,
    LFCUDA_ENCCL = -4,    /* NCCL could not be loaded or a collective failed */
    LFCUDA_ELIMIT = -5    /* scene exceeds a structural limit (stack depth, index range) */
};

typedef struct lfcuda_ctx lfcuda_ctx;

/* Read-only view of the flattened scene arrays of LavaFrame::Scene (LavaFrame/Scene.h:68-101),
 * as produced by Scene::CreateAccelerationStructures (LavaFrame/Scene.cpp:180-231). */
typedef struct LfSceneView {
    /* scene->bvhTranslator.nodes: {vec3 bboxmin, vec3 bboxmax, vec3 LRLeaf} = 9 floats / node
     * (thirdparty/RadeonRays/bvh_translator.h:48-53; uploaded Renderer.cpp:91-97). */
    const float*   bvh_nodes;
    int32_t        num_nodes;
    int32_t        top_bvh_index;      /* bvhTranslator.topLevelIndex (TiledRenderer.cpp:223) */
    /* scene->vertIndices: {int x,y,z} per BVH triangle reference (Scene.h:42-45; Renderer.cpp:99-105) */
    const int32_t* vert_indices;
    int32_t        num_tri_refs;
    /* scene->verticesUVX / normalsUVY: vec4 (xyz + u) / (nxyz + v) (Scene.h:75-76; Renderer.cpp:107-121) */
    const float*   vertices_uvx;
    const float*   normals_uvy;
    int32_t        num_vertices;
    /* scene->transforms: Mat4 float[4][4], data[r] = GLSL column r (Mat4.h:12-25; Renderer.cpp:131-137) */
    const float*   transforms;
    int32_t        num_instances;
    /* scene->materials: 7 x vec4 = 28 floats (Material.h:48-75; Renderer.cpp:123-129) */
    const float*   materials;
    int32_t        num_materials;
    /* scene->lights: 5 x vec3 = 15 floats {position, emission, u, v, (radius, area, type)}
     * (Scene.h:31-40; Renderer.cpp:139-149) */
    const float*   lights;
    int32_t        num_lights;
    /* scene->textureMapsArray: RGBA8 layers of tex_width x tex_height (Scene.h:99-101; Renderer.cpp:151-160) */
    const uint8_t* texture_maps;
    int32_t        tex_width, tex_height, num_textures;
    /* scene->hdrData: cols = RGB fp32 W x H; marginal = vec2 x H; conditional = vec2 x W x H
     * (hdrloader.h:19-28; Renderer.cpp:163-185).  All NULL when the scene has no env map. */
    const float*   hdr_cols;
    const float*   hdr_marginal;
    const float*   hdr_conditional;
    int32_t        hdr_width, hdr_height;
} LfSceneView;

/* Mirrors RenderOptions (LavaFrame/Renderer.h:18-67) restricted to what the path reads, plus the
 * shader #defines TiledRenderer::Init injects (TiledRenderer.cpp:78-91). */
typedef struct LfParams {
    int32_t width, height;           /* renderOptions.resolution */
    int32_t tile_width, tile_height; /* renderOptions.tileWidth/Height */
    int32_t max_depth;               /* uniform maxDepth (TiledRenderer.cpp:516) */
    int32_t enable_rr;               /* #define RR */
    int32_t rr_depth;                /* #define RR_DEPTH n */
    int32_t use_envmap;              /* #define ENVMAP  (useEnvMap && hdrData) */
    int32_t use_constant_bg;         /* #define CONSTANT_BG */
    float   bg_color[3];             /* uniform bgColor */
    float   hdr_multiplier;          /* uniform hdrMultiplier */
    /* implementation knobs (0 = default) */
    int32_t kernel_mode;             /* 0 = wavefront (default), 1 = single megakernel (cross-check) */
    int32_t no_cull;                 /* 1 = visit every pierced box like the reference (closest_hit.glsl:169-198) */
    int32_t count_work;              /* 1 = maintain LfCounters (instrumented kernels, slower) */
    int32_t frames_in_flight;        /* pixel-sample frames batched per wavefront pass; 0 = auto */
} LfParams;

/* Uniforms of the post-process pass beyond invSampleCounter / tonemapIndex (shaders/postprocess.glsl:11-24;
 * TiledRenderer.cpp:539-553; RenderOptions, Renderer.h:35-43).  All zero = plain divide + tonemap. */
typedef struct LfPostParams {
    int32_t use_ca;              /* useCA: chromatic aberration (postprocess.glsl:96-118) */
    int32_t use_ca_distortion;   /* useCADistortion */
    float   ca_distance;         /* caDistance */
    float   ca_p1, ca_p2, ca_p3; /* caP1 (angularity), caP2 (directionality), caP3 (centre) */
    int32_t use_vignette;        /* useVignette (postprocess.glsl:121-124,168-170) */
    float   vignette_intensity, vignette_power;
} LfPostParams;

/* uniform Camera camera (shaders/common/globals.glsl:48-58; TiledRenderer.cpp:507-513) */
typedef struct LfCamera {
    float position[3];
    float right[3];
    float up[3];
    float forward[3];
    float fov;         /* radians */
    float focal_dist;
    float aperture;
} LfCamera;

/* Work counters of the render calls since the last lfcuda_reset_counters (count_work = 1).
 * N_* are the quantities of SURVEY.md §8(d)'s algorithmic-bytes formula. */
typedef struct LfCounters {
    uint64_t samples;        /* pixel-samples generated */
    uint64_t rays_closest;   /* ClosestHit calls */
    uint64_t rays_shadow;    /* AnyHit calls */
    uint64_t inner_visits;   /* N_inner */
    uint64_t leaf_visits;    /* N_leaf (BLAS leaves) */
    uint64_t tri_tests;      /* N_tri */
    uint64_t tlas_visits;    /* N_tlas (TLAS leaves entered) */
    uint64_t light_tests;    /* N_lightTests */
    uint64_t shaded_hits;    /* surface hits that fetched normals + material */
    uint64_t env_nee;        /* EnvSample calls */
    uint64_t env_miss;       /* env lookups on miss */
    uint64_t tex_samples;    /* bilinear RGBA8 material-texture samples */
    /* the share of the traversal counts above that belongs to AnyHit (shadow) rays */
    uint64_t inner_visits_shadow, leaf_visits_shadow, tri_tests_shadow, tlas_visits_shadow, light_tests_shadow;
} LfCounters;

/* Device time per wavefront stage, measured with CUDA events on the context's stream while
 * profiling is on (lfcuda_set_profiling). */
enum { LF_STAGE_GENERATE = 0, LF_STAGE_EXTEND, LF_STAGE_SHADE, LF_STAGE_SHADOW, LF_STAGE_ACCUMULATE,
       LF_STAGE_MEGAKERNEL, LF_STAGE_SAMPLE, LF_STAGE_COUNT };
typedef struct LfStageStats {
    uint64_t launches[LF_STAGE_COUNT];
    double   ms[LF_STAGE_COUNT];
} LfStageStats;

/* ---- life cycle -------------------------------------------------------------------------- */
int  lfcuda_abi_version(void);
/* Replaces `new TiledRenderer(scene, shadersDir)` + Renderer::Init's GL object creation
 * (LavaFrame/Main.cpp:88-97, Renderer.cpp:76-188). */
int  lfcuda_create(lfcuda_ctx** out, int device);
void lfcuda_destroy(lfcuda_ctx* ctx);                /* Renderer::Finish (Renderer.cpp:56-74) */
const char* lfcuda_last_error(const lfcuda_ctx* ctx); /* ctx may be NULL: error of the last failed create */

/* Launch all work of this context on an existing CUDA stream (a cudaStream_t passed as void*);
 * NULL restores the context's own stream. */
int  lfcuda_set_stream(lfcuda_ctx* ctx, void* cuda_stream);
int  lfcuda_synchronize(lfcuda_ctx* ctx);

/* ---- scene upload (Renderer::Init, Renderer.cpp:87-185) -------------------------------------- */
int  lfcuda_upload_scene(lfcuda_ctx* ctx, const LfSceneView* scene);
/* Renderer::Update when scene->instancesModified (Renderer.cpp:190-205): re-upload transforms,
 * materials and the TLAS node range [first_node, first_node + num_tlas_nodes). */
int  lfcuda_update_instances(lfcuda_ctx* ctx, const float* transforms, int32_t num_instances,
                             const float* materials, int32_t num_materials,
                             const float* tlas_nodes, int32_t first_node, int32_t num_tlas_nodes);

/* The same update with the TLAS REBUILT ON THE DEVICE from the instance matrices (csrc/lf_tlas.cu): what Scene::RebuildInstances does on the
 * host (createTLAS, Scene.cpp:106-146; Bvh::Build without SAH, thirdparty/RadeonRays/bvh.cpp:40-243; BvhTranslator::UpdateTLAS,
 * bvh_translator.cpp:61-89,134-140), node for node, plus the instance records.  Only the matrices (and materials) are uploaded.
 * instance_material_ids: NULL, or the materialID of every instance if those changed.  On LFCUDA_ELIMIT the scene must be uploaded again. */
int  lfcuda_update_instances_device(lfcuda_ctx* ctx, const float* transforms, int32_t num_instances,
                                    const float* materials, int32_t num_materials, const int32_t* instance_material_ids);
/* Parity probe: the TLAS nodes the device holds, in the reference's layout (9 floats per node, 2 * num_instances - 1 nodes in the
 * order of bvhTranslator.nodes from topLevelIndex on). */
int  lfcuda_read_tlas_nodes(lfcuda_ctx* ctx, float* nodes_out, int32_t max_nodes, int32_t* num_nodes_out);

/* ---- per-frame uniforms (TiledRenderer::Init :222-227, ::Update :505-521) ----------------------- */
/* Only a change of width / height re-creates (and thereby clears) the accumulation buffer and moves lfcuda_accum_device_ptr's address; tile
 * size, depth, batch size and every other field leave the accumulated image alone (the reference updates them as plain uniforms). */
int  lfcuda_set_params(lfcuda_ctx* ctx, const LfParams* params);
int  lfcuda_set_camera(lfcuda_ctx* ctx, const LfCamera* camera);
int  lfcuda_set_post(lfcuda_ctx* ctx, const LfPostParams* post);   /* NULL = defaults (no CA, no vignette) */

/* ---- the hot path ------------------------------------------------------------------------------- */
/* glClear of accumFBO (TiledRenderer.cpp:478-480). */
int  lfcuda_clear(lfcuda_ctx* ctx);
/* One call = `nframes` draws of pathTraceShader + the tile copy into the accumulation texture
 * (TiledRenderer.cpp:336-344) for tile (tile_x, tile_y), with uniform `frame` taking the values
 * first_frame, first_frame + frame_stride, ...  Samples are added to the accumulation buffer in
 * frame order.  Asynchronous on the context's stream. */
int  lfcuda_render_frames(lfcuda_ctx* ctx, int32_t first_frame, int32_t nframes, int32_t frame_stride,
                          int32_t tile_x, int32_t tile_y);
/* The preview engine "Flareon" (shaders/preview_flareon.glsl:22-61), i.e. what TiledRenderer::Render draws instead of the
 * path-trace pass while camera->isMoving || instancesModified (TiledRenderer.cpp:327-333): one sample per pixel of a
 * pv_width x pv_height viewport (screenSize * previewScale, :330), RNG frame fixed to 1, uniform maxDepth = max_depth
 * (2 while moving, :532), thin lens only if use_dof (#define USE_DOF, :90-91).  Written to the preview target, never
 * accumulated.  Always runs the wavefront kernels.  Asynchronous on the context's stream. */
int  lfcuda_render_preview(lfcuda_ctx* ctx, int32_t pv_width, int32_t pv_height, int32_t max_depth, int32_t use_dof);
/* The preview target through the post-process pass with invSampleCounter = 1, as Present()/SetViewport() display it
 * (TiledRenderer.cpp:361-364,558-562): pv_width * pv_height * 3 floats, rows bottom-up.  is_in_preview = the postShader
 * uniform isInPreview (TiledRenderer.cpp:541: camera->isMoving): non-zero skips the chromatic-aberration branch
 * (postprocess.glsl:128-132).  Synchronises the stream. */
int  lfcuda_read_preview(lfcuda_ctx* ctx, int32_t tonemap_index, int32_t is_in_preview, float* rgb_out);
/* Copy the accumulation buffer (running SUM, W*H*3 floats, rows bottom-up like glGetTexImage,
 * TiledRenderer.cpp:399-414) to host memory.  Synchronises the stream. */
int  lfcuda_read_accum(lfcuda_ctx* ctx, float* rgb_out);
/* The post-process pass (shaders/postprocess.glsl:126-172): accum * (1/sample_count) [or the chromatic-aberration
 * fetches], tonemap `tonemap_index` (0 = identity), vignette; rows bottom-up; W*H*3 floats.  GetOutputBufferHDR's payload. */
int  lfcuda_read_output(lfcuda_ctx* ctx, float inv_sample_counter, int32_t tonemap_index, float* rgb_out);
/* Same, converted to 8-bit like glGetTexImage(GL_RGB, GL_UNSIGNED_BYTE) (TiledRenderer.cpp:382-397). */
int  lfcuda_read_output_u8(lfcuda_ctx* ctx, float inv_sample_counter, int32_t tonemap_index, uint8_t* rgb_out);

/* Device pointer of the accumulation buffer (W*H*3 floats, rows bottom-up) and its element count,
 * so that a caller-side collective (torch.distributed / NCCL) can reduce it in place. */
int  lfcuda_accum_device_ptr(lfcuda_ctx* ctx, void** dev_ptr, size_t* num_floats);

/* Parity probe 1 (north_star: primary-hit IDs and t): trace only the camera ray of `frame` for every
 * pixel of the full frame and return, per pixel (rows bottom-up): t, triID.x (first vertex index of
 * the hit triangle, -1 if none), matID (-1 if none), is_emitter (1 when an analytic light is nearest). */
int  lfcuda_read_primary_hits(lfcuda_ctx* ctx, int32_t frame, float* t_out, int32_t* tri_out,
                              int32_t* mat_out, int32_t* emitter_out);

/* ---- multi-GPU: spp split, accumulation buffers summed with NCCL ------------------- */
/* 128-byte ncclUniqueId for rank 0 to broadcast. */
int  lfcuda_nccl_unique_id(void* id128_out);
int  lfcuda_nccl_init(lfcuda_ctx* ctx, const void* id128, int32_t rank, int32_t nranks);
/* ncclAllReduce(sum, float32) of the accumulation buffer, in place, on the context's stream. */
int  lfcuda_reduce(lfcuda_ctx* ctx);

/* ---- multi-GPU in one process: a group of contexts behind one renderer --------------------------------------------------
 * The reference constructs ONE renderer (LavaFrame/Main.cpp:88-97); CudaRenderer(scene, dir, devices) therefore drives several GPUs
 * from one process through a group: one context + one host thread per device, the scene replicated, the frames of every render
 * call dealt round-robin to the devices (same spp split as above), and the devices' accumulation buffers summed INSIDE the
 * post-process kernel of the group's first device, which reads the peers' buffers over NVLink through CUDA peer access
 * (no NCCL, no staging buffer, the local sums stay local so rendering continues after a read-out).
 * Replaces what SURVEY 8(b) sketched as lfcuda_create(ctx**, const int* devices, int ndev). */
typedef struct lfcuda_group lfcuda_group;
int  lfcuda_group_create(lfcuda_group** out, const int32_t* devices, int32_t ndev);   /* 1 <= ndev <= 16 */
void lfcuda_group_destroy(lfcuda_group* g);
const char* lfcuda_group_last_error(const lfcuda_group* g);                           /* g may be NULL: error of the last failed create */
int  lfcuda_group_size(const lfcuda_group* g);
lfcuda_ctx* lfcuda_group_ctx(lfcuda_group* g, int32_t index);                         /* the context of device `index` (preview, probes, counters) */
int  lfcuda_group_upload_scene(lfcuda_group* g, const LfSceneView* scene);            /* every device, in parallel */
int  lfcuda_group_update_instances(lfcuda_group* g, const float* transforms, int32_t num_instances,
                                   const float* materials, int32_t num_materials,
                                   const float* tlas_nodes, int32_t first_node, int32_t num_tlas_nodes);
int  lfcuda_group_update_instances_device(lfcuda_group* g, const float* transforms, int32_t num_instances,
                                          const float* materials, int32_t num_materials, const int32_t* instance_material_ids);
int  lfcuda_group_set_params(lfcuda_group* g, const LfParams* params);
int  lfcuda_group_set_camera(lfcuda_group* g, const LfCamera* camera);
int  lfcuda_group_set_post(lfcuda_group* g, const LfPostParams* post);
int  lfcuda_group_clear(lfcuda_group* g);
int  lfcuda_group_synchronize(lfcuda_group* g);
/* lfcuda_render_frames for the group: device i of n renders frames first + i*stride, first + (i+n)*stride, ...  Asynchronous. */
int  lfcuda_group_render_frames(lfcuda_group* g, int32_t first_frame, int32_t nframes, int32_t frame_stride,
                                int32_t tile_x, int32_t tile_y);
/* lfcuda_read_output[_u8] / lfcuda_read_accum of the SUM of the devices' accumulation buffers (added in device order). */
int  lfcuda_group_read_output(lfcuda_group* g, float inv_sample_counter, int32_t tonemap_index, float* rgb_out);
int  lfcuda_group_read_output_u8(lfcuda_group* g, float inv_sample_counter, int32_t tonemap_index, uint8_t* rgb_out);
int  lfcuda_group_read_accum(lfcuda_group* g, float* rgb_out);

/* ---- instrumentation -------------------------------------------------------------------------------- */
int  lfcuda_reset_counters(lfcuda_ctx* ctx);
int  lfcuda_get_counters(lfcuda_ctx* ctx, LfCounters* out);
int  lfcuda_set_profiling(lfcuda_ctx* ctx, int32_t on);
int  lfcuda_get_stage_stats(lfcuda_ctx* ctx, LfStageStats* out);   /* synchronises */
/* Number of CUDA kernels launched by this context since creation. */
int  lfcuda_get_launch_count(lfcuda_ctx* ctx, uint64_t* out);
/* Read-bandwidth probe for the roofline denominators MEASURED_PEAKS.json lacks (L2): `iters` passes of 16-byte
 * read-only loads over a `bytes`-sized buffer by a full grid; GB/s of the best pass.  A working set well below the
 * L2 size measures L2 -> SM bandwidth, one far above it measures HBM reads. */
int  lfcuda_measure_read_bandwidth(lfcuda_ctx* ctx, size_t bytes, int32_t iters, double* gbps_out);

/* Node-fetch probe: the rate (GB/s of 64-byte records) at which this GPU delivers one-record-per-lane fetches issued exactly like the
 * traversal's inner-node fetch (2 x 256-bit read-only loads, every lane a different record, dependent chain) from a table of
 * `table_bytes` (8 MiB: L2 hits; 32 KiB: L1 hits).  This is the ceiling of the bound ncu names for the extend kernel (L1 data-pipe
 * wavefronts), measured in the run instead of read from a file. */
int  lfcuda_measure_node_fetch(lfcuda_ctx* ctx, size_t table_bytes, int32_t iters, double* gbps_out);

/* ---- bottom-level BVH build on the device (csrc/lf_blas.cu, csrc/lf_blas_build.h) ----------------
 * The BVH of ONE mesh, node for node what the reference builds on the host at scene load: Mesh::BuildBVH (LavaFrame/Mesh.cpp:93-111) ->
 * RadeonRays::SplitBvh(2.0f, 64, 0, 0.001f, 0)::Build (Mesh.h:18; thirdparty/RadeonRays/split_bvh.cpp:11-289: binned-SAH object splits - with
 * max_split_depth 0 the spatial-split branch is never entered -, two-pointer partition, halving fallback, right child first), flattened in
 * pre-order as BvhTranslator::ProcessBLASNodes does (bvh_translator.cpp:35-60).  No context is needed; the call is thread-safe.
 *   prim_bounds   num_prims x 6 floats: pmin.xyz, pmax.xyz of every triangle (what Mesh::BuildBVH passes to Bvh::Build)
 *   out_nodes     capacity (2 * num_prims - 1) x 9 words.  Node k: pmin.xyz, pmax.xyz (fp32) and three INT32: inner node = (left, right, 0)
 *                 with left == k + 1; leaf = (startidx into out_indices, numprims 1..3, 1).  Node 0's box is Bvh::Bounds().
 *   out_indices   num_prims primitive indices in leaf order (Bvh::GetIndices; Scene.cpp:196-209 forms vertIndices from them)
 * Identical means identical bits, the sign of a zero box plane included: std::min / std::max keep the first of +0 / -0 they meet, so that sign
 * depends on the order in which the reference grows a box; the order is restated (lf_blas_build.h, acc_zero).  info->negative_zero only reports
 * that some bound is -0.0. */
typedef struct LfBlasInfo {
    int32_t num_nodes;       /* Bvh::m_nodecnt */
    int32_t num_indices;     /* Bvh::GetNumIndices() == num_prims (no reference is duplicated without spatial splits) */
    int32_t height;          /* Bvh::GetHeight() */
    int32_t negative_zero;
    int32_t levels;          /* height + 1 */
    int32_t launches;        /* kernels launched */
    float   build_ms;        /* device time of the build steps (CUDA events) */
    float   total_ms;        /* wall time of the call: allocation, upload, build, read-back */
} LfBlasInfo;
int  lfcuda_build_blas(int32_t device, const float* prim_bounds, int32_t num_prims, float traversal_cost, int32_t num_bins,
                       float* out_nodes, int32_t* out_indices, LfBlasInfo* info);

#ifdef __cplusplus
}
#endif
#endif /* LFCUDA_H */
