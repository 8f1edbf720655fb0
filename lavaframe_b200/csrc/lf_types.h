// lf_types.h — device-side scene layout and per-path state of the CUDA path tracer (host + device).
//
// The arrays the reference uploads as GL textures (LavaFrame/Renderer.cpp:87-185) are re-packed once at
// upload (lf_repack.cpp) into 16-byte-aligned float4 records so that every traversal step is a short run
// of vectorised read-only loads (LDG.E.128.CONSTANT):
//
//   inner node (64 B)  : both child boxes + both child references.  A child reference says what the child
//                        IS, so the child's own LRLeaf texel (closest_hit.glsl:102) is never fetched:
//                          ref >= 0                      inner node, index into `nodes` (x4 float4)
//                          ref <  0, bit 30 clear        BLAS leaf: bits 0-23 first triangle ref, bits 24-29 count-1
//                          ref <  0, bit 30 set          TLAS leaf: bits 0-23 instance index (-1 = stack sentinel)
//   triangle (64 B)    : v0|u0, e0=v1-v0|u1, e1=v2-v0|u2, pad, in BVH leaf order (the vertIndices indirection of
//                        closest_hit.glsl:113-117 is resolved at upload; e0/e1 are the same fp32 subtractions
//                        the shader does per test, :119-120)
//   triangle normals   : n0|v0, n1|v1, n2|v2 in the same order (pathtrace.glsl:19-21)
//   instance (160 B)   : rows of inverse(M) (closest_hit.glsl:159-160, hoisted), meta (BLAS root ref, matID),
//                        rows of M (:143), rows of transpose(inverse(mat3(M))) (pathtrace.glsl:31)
//   light (112 B)      : the 5 texels of lightsTex + the per-call derived values of closest_hit.glsl:29-35
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "lfcuda.h"

namespace lf {

constexpr int   kRefLeafBit   = int(0x80000000u);
constexpr int   kRefTlasBit   = 0x40000000;
constexpr int   kRefSentinel  = -1;             // the reference's "-1" stack marker (closest_hit.glsl:72,162); TLAS bit set
constexpr int   kMaxLeafTris  = 64;
constexpr int   kTriStride    = 4;              // float4 per triangle record (3 used; 64-byte records keep 256-bit loads aligned)
constexpr int   kInstStride   = 10;             // float4 per instance record
constexpr int   kLightStride  = 7;              // float4 per light record
constexpr int   kBlockThreads = 128;            // threads per CTA of the traversal kernels

__host__ __device__ inline bool ref_is_inner(int r) { return r >= 0; }
__host__ __device__ inline bool ref_is_tlas_leaf(int r) { return r < 0 && (r & kRefTlasBit); }
__host__ __device__ inline int  ref_leaf_first(int r) { return r & 0x00ffffff; }
__host__ __device__ inline int  ref_leaf_count(int r) { return ((r >> 24) & 0x3f) + 1; }
__host__ __device__ inline int  ref_instance(int r) { return r & 0x00ffffff; }
inline int make_blas_leaf_ref(int first, int count) { return kRefLeafBit | ((count - 1) << 24) | first; }
inline int make_tlas_leaf_ref(int inst) { return kRefLeafBit | kRefTlasBit | inst; }   // never -1: inst < 2^24

struct DevScene {
    const float4* nodes;        // 4 per inner node
    const float4* tris;         // kTriStride per triangle ref
    const float4* trinrm;       // 3 per triangle ref
    const int*    tri_vx;       // vertIndices[ref].x (primary-hit probe only)
    const float4* inst;         // kInstStride per instance
    const float4* materials;    // 7 per material
    const float4* lights;       // kLightStride per light
    const float2* marginal;     // hdr_h
    const float2* conditional;  // hdr_w * hdr_h
    cudaTextureObject_t tex_maps;   // RGBA8 layered 2D, point sampled, raw uchar4 (filtered in fp32 by the kernel)
    cudaTextureObject_t hdr_tex;    // RGBA32F 2D, point sampled (filtered in fp32 by the kernel)
    cudaTextureObject_t tris_tex;   // the triangle records as a linear float4 texture (LF_TRI_TEX experiment)
    cudaTextureObject_t nodes_tex;  // the inner-node array as a linear float4 texture (LF_NODE_TEX experiment: node fetches through the TEX pipe)
    int top_ref;
    int num_lights, num_materials, num_instances;
    int tex_w, tex_h, num_tex;
    int hdr_w, hdr_h;
};

// Uniforms of the path-trace program (uniforms.glsl:6-38, globals.glsl:108) + injected #defines.
struct DevParams {
    int   width, height, tile_w, tile_h;
    int   max_depth, enable_rr, rr_depth, use_envmap, use_constant_bg;
    float bg[3];
    float hdr_multiplier, hdr_resolution;
    float inv_tiles_x, inv_tiles_y;       // 1/((float)W/tileW) (TiledRenderer.cpp:226-227)
    float cam_pos[3], cam_right[3], cam_up[3], cam_fwd[3];
    float cam_scale;                      // tan(fov * 0.5) (renderer.glsl:51), evaluated once on the host
    float focal_dist, aperture;
    int   tile_x, tile_y;
    int   first_frame, frame_stride, num_frames;   // frames of the current batch
    int   pix_w8, pix_h4;                 // tile size rounded up to 8x4 pixel blocks (warp-coherent primary rays)
    int   slots_per_frame;                // pix_w8 * pix_h4
    // preview engine (shaders/preview_flareon.glsl): the batch is rows [pv_y0, pv_y0 + tile_h) of a pv_w x pv_h viewport
    int   preview, pv_w, pv_h, pv_y0, use_dof;
};

// Per-path state; one slot = one pixel-sample of the batch.  Every field is a 16-byte element, and fields that are read or written
// TOGETHER share a 32-byte record (Pair: element i of a field lives at p[2 i]), i.e. one DRAM sector: after the first bounce the live
// slots are scattered, and a 16-byte access to a plain array drags the dead neighbour's half of the sector along.
template <class T>
struct Pair {
    T* p;
    __host__ __device__ T& operator[](size_t i) const { return p[2 * i]; }
};
// LF_SAMPLE_REMAT = 1 (default): k_sample re-reads the hit's material from the material table (a few records, always cached) instead of
// being handed 4 float4 of it per path through the path state; only what a TEXTURE changed (albedo, metallic, roughness) is still handed
// over, and only for hits on textured materials.  0 = the round-2 layout before that (sf0..sf4 always), kept for A/B runs.
#ifndef LF_SAMPLE_REMAT
#define LF_SAMPLE_REMAT 1
#endif
struct PathSoA {
    Pair<float4> ray_o;     // origin.xyz                                                           } one record
    Pair<float4> ray_d;     // direction.xyz                                                        }
    Pair<float4> hit_f;     // t, bary u, bary v                                                    } one record
    Pair<int4>   hit_i;     // triangle ref, instance, emitter light index (-1: surface), matID    }
    Pair<float4> thr;       // throughput.xyz, bsdfSampleRec.pdf                                    } one record
    Pair<float4> rad;       // radiance.xyz                                                         }
#if LF_SAMPLE_REMAT
    // what k_shade hands to k_sample for a path that goes on: two full records, [sf0 | absn] and [rng | hit_p]
    Pair<float4> sf0;       // normal.xyz, eta                                                      } one record
    Pair<float4> absn;      // absorption.xyz                                                       }
    Pair<uint4>  rng;       // pcg4d state                                                          } one record
    Pair<float4> hit_p;     // first hit point (world), closest_hit.glsl:139,143; .w = material index (bits) }
    Pair<float4> sf1;       // albedo.xyz, specular                          } one record, written and read only for hits on materials with an
    Pair<float4> sf2;       // metallic, roughness, specularTint, sheenTint  } albedo or metallic-roughness texture
    Pair<float4> stale;     // state.mat.emission of the last shaded surface (pathtrace.glsl:246-253 on emitter hits; scenes with lights only)  } one record
    Pair<float4> sh_T;      // throughput the NEE sum is multiplied with (:266; fused kernel only)                                               }
#else
    Pair<float4> absn;      // absorption.xyz                                                       } one record
    Pair<float4> stale;     // state.mat.emission of the last shaded surface (pathtrace.glsl:246-253 on emitter hits; scenes with lights only)
    Pair<uint4>  rng;       // pcg4d state                                                          } one record
    Pair<float4> hit_p;     // first hit point (world), closest_hit.glsl:139,143 (paths that go on) }
    // what DisneySample needs of `State`, handed from the hit/NEE kernel to the sample kernel
    Pair<float4> sf0;       // normal.xyz, eta                                                      } one record
    Pair<float4> sf1;       // albedo.xyz, specular                                                 }
    Pair<float4> sf2;       // metallic, roughness, specularTint, sheenTint                         } one record
    Pair<float4> sf3;       // sheen, clearcoat, clearcoatRoughness, specTrans                      }
    Pair<float4> sf4;       // -log(extinction)/atDistance .xyz, subsurface                         } one record
    Pair<float4> sh_T;      // throughput the NEE sum is multiplied with (:266; fused kernel only)  }
#endif
    // next-event estimation requests of the current bounce.  Candidate 0 is the environment ray when there is one, else the analytic
    // light's; candidate 1 the analytic light's when both exist (so a request with one ray touches two records, not three)
    Pair<float4> sh_o;      // surfacePos.xyz, candidate mask in .w bits (bit 0: candidate 0, bit 1: candidate 1)   } one record
    Pair<float4> sh_d0;     // candidate 0: direction.xyz, max distance                                             }
    Pair<float4> sh_c0;     // candidate 0: weighted contribution (pathtrace.glsl:155 / :197)       } one record
    Pair<float4> sh_c1;     // candidate 1: contribution                                            }
    float4* sh_d1;          // candidate 1: direction.xyz, max distance
};

struct Queues {
    int* active[2];    // ping-pong queues of live path slots
    int* shadow;       // slots with at least one shadow ray this bounce
    int* sample;       // surface hits that go on to the BSDF-sample kernel
    int* counts;       // [5][max_depth + 2]: active count, shadow count, extend cursor, shadow cursor, sample count per bounce
    int  stride;       // max_depth + 2
};

struct DevCounters {   // mirrors LfCounters, device side
    unsigned long long v[17];
};
enum { C_SAMPLES = 0, C_RAYS_CLOSEST, C_RAYS_SHADOW, C_INNER, C_LEAF, C_TRI, C_TLAS, C_LIGHT, C_SHADED, C_ENV_NEE, C_ENV_MISS, C_TEX,
       C_INNER_SH, C_LEAF_SH, C_TRI_SH, C_TLAS_SH, C_LIGHT_SH };

}  // namespace lf
