/*
 * oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  This library is a single-threaded (optionally
 * tile-parallel for the CPU baseline) C++ restatement of the reference's HLSL build passes,
 * traversal and shaders.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; nothing under dxrexperiments_b200/ does.
 *
 * The reference (philcn/DXRExperiments) is Windows / D3D12 / DXC only and its ray-tracing
 * scheduler lives in a closed binary (dxrfallbackcompiler.dll v1.5-dxr, not under
 * /root/reference), so it cannot be compiled or run here: there is no oracle/_ref.
 * Pinning status per stage:
 *   - scene AABB, Morton codes, sort order, BVH structure, hit/miss behaviour (culling, masks,
 *     instance transforms): PINNED by re-running the logic of the reference's own unit tests
 *     (externals/D3D12RaytracingFallback/src/FallbackLayerUnitTests/fallbacklayerunittests.cpp
 *      :2569-2615, 2627-2701, 2831-2945, 3630-4078; BVHValidator.cpp:58-176) — see
 *     tests/test_oracle_reference_kats.py.
 *   - treelet optimisation (FL/FindTreelets.hlsl, FL/TreeletReorder.hlsl): restated in oracle_treelet.cpp;
 *     the reference's own tests pin only topology sanity for it (UT:2957-3068, re-run in
 *     tests/test_oracle_reference_kats.py), the arithmetic is pinned by code inspection.
 *   - shading, RNG, accumulation, denoiser: PARITY UNPINNED by any reference test (the app has
 *     none); pinned by code inspection only, plus frozen known-answer vectors of our own.
 *
 * All pointers are HOST pointers.  Struct types come from include/rt_types.h.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include "../include/rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_blas orc_blas;
typedef struct orc_tlas orc_tlas;

/* ---- stage-level functions (white-box parity with the CUDA build kernels) ---- */

/* FL/CalculateSceneAABBFromPrimitives.hlsl:16-40 — exact min/max; out = {min xyz, max xyz}. */
void orc_scene_aabb(const rt_primitive *prims, uint32_t n, float out[6]);
/* FL/CalculateMortonCodes.hlsli:73-118 + CalculateMortonCodesForPrimitives.hlsl:16-31. */
void orc_morton_codes(const rt_primitive *prims, uint32_t n, const float aabb[6], uint32_t *codes);
/* Morton code of one centroid (used for TLAS: CalculateMortonCodesForAABBs.hlsl:25-28). */
uint32_t orc_morton_code_from_centroid(const float c[3], const float aabb[6]);
/* Order defined by FL/BitonicSortCommon.hlsli:37-47: ascending key, ties by ascending index. */
void orc_sort_pairs(const uint32_t *codes, uint32_t n, uint32_t *sorted_codes, uint32_t *perm);
/* FL/BuildBVHSplits.hlsli:35-143; nodes[2n-1] = {parent,left,right}; leaves have left=right=0. */
void orc_build_hierarchy(const uint32_t *sorted_codes, uint32_t n, rt_hierarchy_node *nodes);
/* FL/TreeletReorder.cpp:38-109 on a hierarchy over n sorted triangles, in place. */
void orc_treelet_optimise(uint32_t n, rt_hierarchy_node *hier, const rt_primitive *sorted_prims, uint32_t build_flags);
/* initRand / nextRand: assets/shaders/RaytracingUtils.hlsli:26-45. */
uint32_t orc_init_rand(uint32_t v0, uint32_t v1);
float orc_next_rand(uint32_t *state);

/* ---- acceleration structures ---- */

/* BLAS over triangle geometries (host pointers in the descs).  Follows FL/GpuBVH2Builder.cpp:137-328 including the
 * treelet pass: build_flags PREFER_FAST_BUILD -> 0 passes, PREFER_FAST_TRACE -> 3, otherwise 1
 * (FL/TreeletReorder.cpp:66-80).  orc_blas_hierarchy returns the hierarchy AFTER that pass. */
orc_blas *orc_blas_build(const rt_geometry_desc *geoms, uint32_t n_geoms, uint32_t build_flags);
void orc_blas_free(orc_blas *b);
uint32_t orc_blas_num_prims(const orc_blas *b);
const rt_primitive *orc_blas_unsorted_prims(const orc_blas *b);
const float *orc_blas_scene_aabb(const orc_blas *b);
const uint32_t *orc_blas_morton(const orc_blas *b);        /* per input primitive */
const uint32_t *orc_blas_sorted_morton(const orc_blas *b);
const uint32_t *orc_blas_perm(const orc_blas *b);          /* sorted slot -> input primitive */
const rt_hierarchy_node *orc_blas_hierarchy(const orc_blas *b);
const uint8_t *orc_blas_blob(const orc_blas *b, uint64_t *bytes);

/* PERFORM_UPDATE (FL/GpuBVH2Builder.cpp:152-204, FL/ComputeAABBs.hlsli:38-67): re-load the triangles (same count, same
 * order) into the cached sorted slots and re-fit the boxes on the stored topology, in place.  -1 if the count differs. */
int orc_blas_update(orc_blas *b, const rt_geometry_desc *geoms, uint32_t n_geoms);
const uint32_t *orc_blas_sort_cache(const orc_blas *b); /* n: load order -> sorted slot  (FL/RearrangeTriangles.hlsl:25-28) */
const uint32_t *orc_blas_parents(const orc_blas *b);    /* 2n-1: parent of every node    (FL/ComputeAABBs.hlsli:160-164) */

/* TLAS over instances; inst[i].blas must hold an orc_blas* cast to uint64_t. */
orc_tlas *orc_tlas_build(const rt_instance_desc *inst, uint32_t n, uint32_t build_flags);
int orc_tlas_update(orc_tlas *t, const rt_instance_desc *inst, uint32_t n);
const uint32_t *orc_tlas_sort_cache(const orc_tlas *t);
const uint32_t *orc_tlas_parents(const orc_tlas *t);
void orc_tlas_free(orc_tlas *t);
const uint8_t *orc_tlas_blob(const orc_tlas *t, uint64_t *bytes);
const uint32_t *orc_tlas_sorted_morton(const orc_tlas *t);
const uint32_t *orc_tlas_perm(const orc_tlas *t);

/* ---- traversal: FL/TraverseFunction.hlsli:520-799 + TraverseShader.hlsli:21-73 ---- */
/* Traversal with the hit groups' any-hit / intersection programs (rt_types.h RT_ANYHIT_*, RT_INTERSECTION_*);
 * leaf_slot carries HitKind() in bits 31:24. */
void orc_trace_hit_groups(const orc_tlas *t, const rt_ray *rays, uint64_t n, uint32_t ray_flags, uint32_t instance_mask,
                          uint32_t ray_contribution, uint32_t geometry_multiplier, const rt_hit_group_programs *programs,
                          uint32_t n_programs, rt_hit *hits, int threads);
void orc_trace(const orc_tlas *t, const rt_ray *rays, uint64_t n, uint32_t ray_flags, uint32_t instance_mask,
               rt_hit *hits, rt_trace_stats *stats /* nullable, accumulated */, int threads);

/* ---- pipelines ---- */

/* Two things that are fixed in the reference and options here (process-wide; set before rendering):
 *   max_radiance_ray_depth: MAX_RADIANCE_RAY_DEPTH (assets/shaders/RaytracingCommon.hlsli:11), 1 = the reference, 2 = the
 *     Phong-lobe bounce continues one more level (BASELINE config 3 "2-bounce"; parity is pinned at 1 only);
 *   half_render_targets: emulate the R16G16B16A16_FLOAT render targets (src/DXRExperimentsApp.cpp:28): every value stored to
 *     an output (accumulation, AOVs, denoiser passes) is rounded to fp16 and back. */
void orc_set_render_options(uint32_t max_radiance_ray_depth, int half_render_targets);

/* One DispatchRays of ProgressiveRaytracing.hlsl over width x height; `accum` (RGBA fp32,
 * width*height*4) is gOutput: read-modify-written with the running mean of RayGen:36-38.
 * recs[i] is the hit record of instance i.  threads<=1: single-threaded. */
void orc_render_progressive(const orc_tlas *t, const rt_hit_record *recs, uint32_t n_recs, const rt_env_cube *env,
                            const rt_per_frame_constants *frame, uint32_t width, uint32_t height, float *accum,
                            int threads, rt_ray_counts *counts /* nullable, accumulated */,
                            rt_trace_stats *secondary_stats /* nullable: stats of the incoherent closest-hit rays */);

/* One DispatchRays of RealtimeRaytracing.hlsl: writes the two AOVs (RGBA fp32). */
void orc_render_realtime(const orc_tlas *t, const rt_hit_record *recs, uint32_t n_recs, const rt_env_cube *env,
                         const rt_per_frame_constants *frame, uint32_t width, uint32_t height, float *direct,
                         float *indirect_specular, int threads, rt_ray_counts *counts);

/* DenoiseCompositor::dispatch (src/DenoiseCompositor.cpp:109-148): pass H then pass V. */
void orc_denoise(const float *direct, const float *indirect_specular, float *tmp, float *out, uint32_t width,
                 uint32_t height, const rt_denoiser_params *params, int threads);

/* Primary ray of pixel (x,y) exactly as RayGen builds it (ProgressiveRaytracing.hlsl:17-31);
 * jitter_scale = 30 (progressive) or 10 (realtime). */
void orc_primary_ray(const rt_per_frame_constants *frame, uint32_t width, uint32_t height, uint32_t x, uint32_t y,
                     float jitter_scale, rt_ray *out);

void orc_primary_rays(const rt_per_frame_constants *frame, uint32_t width, uint32_t height, float jitter_scale,
                      rt_ray *out /* width*height, row major */);

/* Environment lookup used by both the oracle and (restated) by the CUDA miss shader. */
void orc_sample_env(const rt_env_cube *env, const float dir[3], float rgb[3]);

#ifdef __cplusplus
}
#endif
#endif
