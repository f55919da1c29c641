/*
 * vxo.h -- CPU ORACLE for the voxel-rt per-pixel hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm (Berry2460/voxel-rt):
 *   src/fshader.glsl   (castRay / main: the per-pixel path)
 *   src/render.cpp     (getVoxelIndex, fixDepthField, computeDepthIndices, updatePartialGeometry)
 *   src/level.cpp      (initVoxels, placeBush, placeTrunk, removeSphere)
 *   src/controls.cpp   (doDestroy)
 * Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libvxrt.so) never links, loads or calls it.
 *
 * PINNING.  The reference ships no tests / golden vectors, so the oracle is
 * pinned against the reference ITSELF, run in the build container (oracle/ref_build/ -> oracle/_ref/):
 *   - host half : the reference's level.cpp / render.cpp / controls.cpp / main.cpp compiled
 *                 unmodified against a no-op GL shim (libref_host.so)
 *   - shader half: the reference's fshader.glsl compiled UNMODIFIED as C++ through the reference's
 *                 own vendored GLM (libref_shader.so); GLSL leaves float evaluation details to the
 *                 implementation, this build fixes them to "IEEE binary32, no contraction, GLM's
 *                 normalize/length/mat*vec operation order" and the oracle restates exactly that.
 * tests/test_oracle_vs_ref.py compares the two bit-for-bit; tests/golden/ holds fingerprints made by
 * tests/golden/make_golden.py from the _ref build so the pin also holds where /root/reference is
 * absent (the GPU box).
 *
 * Conventions the GLSL spec leaves open, fixed here (and mirrored by the CUDA path):
 *   - float->int conversion: truncation; NaN / out-of-range -> INT_MIN (x86 cvttss2si behaviour,
 *     which is what the compiled reference shader does)
 *   - int arithmetic wraps (two's complement)
 *   - sign(NaN) = 0; max(0,NaN) = 0
 *   - pixel centre -> vPos: ((px+0.5)/W)*2-1, ((py+0.5)/H)*2-1, row 0 = bottom (GL window origin)
 *   - UNORM8 store: clamp to [0,1] (NaN -> 0), floor(c*255+0.5)
 */
#ifndef VXO_H
#define VXO_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VXO_MAX_LOCAL_LIGHTS 16

typedef struct { int32_t w, h, d; } vxo_dims;          /* x extent, y extent (up), z extent */

/* frame parameters == the shader's uniforms, fshader.glsl:20-26, render.cpp:289-296 */
typedef struct {
    float cam_pos[3];
    float cam_rotation[2];      /* uploaded by the reference, never read by the shader */
    float light_pos[3];
    float aspect;
    float rotate[16];           /* column-major mat4 (GL_FALSE transpose, render.cpp:294) */
    int32_t view_depth_field;
    float lights[VXO_MAX_LOCAL_LIGHTS][4];   /* xyz, w = diffuse weight; inactive if any of xyz < 0 */
} vxo_frame;

/* per-ray side outputs of castRay (the shader's globals hitPos / hitNormal / stepCount) */
typedef struct {
    float hit_pos[3];
    float hit_normal[3];
    float steps;                /* iterations executed by THIS call */
    int32_t hit_set;            /* 1 iff this call wrote hitPos (i.e. returned >= 0) */
} vxo_ray_out;

uint64_t vxo_fnv1a64(const void* data, size_t nbytes);

/* ---- level (level.cpp) ------------------------------------------------------------------ */
void vxo_init_default_level(int32_t* vox);              /* 512x96x512; render.cpp:349-352 + level.cpp:82-138 */
int  vxo_host_index(vxo_dims g, int x, int y, int z);   /* render.cpp:189-196 */
void vxo_place_voxel(int32_t* vox, vxo_dims g, int x, int y, int z, int32_t v);   /* render.cpp:256-262 */
void vxo_destroy_voxel(int32_t* vox, vxo_dims g, int x, int y, int z);            /* render.cpp:265-271 */

/* ---- depth field (render.cpp) ------------------------------------------------------------ */
int  vxo_depth_offsets(float* dist, int32_t* xyz, int max_entries);   /* render.cpp:66-100; returns count (1419) */
void vxo_fix_depth_field(int32_t* vox, vxo_dims g, int x, int y, int z);          /* render.cpp:226-253 */
void vxo_compute_depth_field(int32_t* vox, vxo_dims g, int nthreads);             /* render.cpp:273-286 over all z */

/* ---- edits (level.cpp / controls.cpp / render.cpp) ---------------------------------------- */
void vxo_remove_sphere(int32_t* vox, vxo_dims g, int cx, int cy, int cz, int radius);   /* level.cpp:30-56 */
/* controls.cpp:100-110: centre = camPos + 15*camDir; removeSphere(ivec3(centre), 7); returns centre (float) */
void vxo_do_destroy(int32_t* vox, vxo_dims g, const float cam_pos[3], const float cam_dir[3], float centre_out[3]);
/* render.cpp:204-223: the (first element, element count) list of glBufferSubData calls; returns #calls
 * (counts all calls even beyond max_calls) */
int  vxo_partial_ranges(vxo_dims g, const float start[3], const float end[3],
                        int64_t* first, int32_t* count, int max_calls);

/* ---- shader (fshader.glsl) ---------------------------------------------------------------- */
int32_t vxo_shader_index(vxo_dims g, int32_t x, int32_t y, int32_t z);            /* fshader.glsl:33-52 */
int32_t vxo_cast_ray(const int32_t* vox, vxo_dims g, const float start[3], const float dir[3],
                     int32_t dist, vxo_ray_out* out);                            /* fshader.glsl:59-129 */

/* Render rows [y0,y1) of a width x height frame (row 0 = bottom).  Any output pointer may be NULL.
 * All arrays are FULL-FRAME sized (width*height entries), only rows [y0,y1) are written.
 *   rgba_f32   : 4 floats / pixel, the shader's fColor before the UNORM8 store
 *   rgba8      : 4 bytes / pixel
 *   hit_index  : primary castRay return value (-1 = miss)
 *   steps      : primary-ray stepCount (iterations), as uint16
 *   occl_mask  : bit0 = global-light ray occluded, bit(1+i) = local light i occluded
 *   cast_mask  : bit0 = global-light ray cast, bit(1+i) = local light i ray cast (reference rule)
 *   counters   : [0] rays primary, [1] rays global, [2] rays local, [3] voxel fetches (= sum of
 *                castRay iterations over every cast ray), [4] hit pixels; accumulated (+=)
 * Follows fshader.glsl:131-190.  nthreads <= 0: all OpenMP threads. */
void vxo_render(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height,
                int y0, int y1, float* rgba_f32, uint8_t* rgba8, int32_t* hit_index,
                uint16_t* steps, uint32_t* occl_mask, uint32_t* cast_mask,
                uint64_t counters[5], int nthreads);

/* one pixel; same outputs for a single (px,py); hit_pos/hit_normal = primary ray's */
void vxo_shade_pixel(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height,
                     int px, int py, float rgba[4], int32_t* hit_index, float* steps,
                     uint32_t* occl_mask, uint32_t* cast_mask, float hit_pos[3], float hit_normal[3],
                     uint64_t counters[5]);

int vxo_num_threads(void);

/* ---- analysis only (oracle/vxo_analysis.c, scripts/where_iterations_go.py; not part of the restatement) ----------
 * cell[kind][outcome]: kind 0 primary / 1 global-light / 2 local-light ray; outcome 0 hit / 1 left the grid /
 * 2 budget exhausted.  "after_cull": iterations that came after the ray's cell first lay beyond every grid row holding
 * a solid, in its direction of travel (the CUDA path's occupancy-summary culling ends the ray there).  "dark": rays
 * from a surface that faces away from their light (N.L <= 0; the CUDA path does not trace them). */
typedef struct {
    uint64_t rays, iterations, jumps;
    uint64_t rays_culled, iterations_after_cull;
    uint64_t rays_dark, iterations_dark, iterations_dark_after_cull;
} vxo_profile_cell;
typedef struct {
    vxo_profile_cell cell[3][3];
    uint64_t longest[3];            /* most iterations of a single ray, per kind */
    int32_t ymin, ymax;             /* rows that hold a solid voxel */
} vxo_profile;
void vxo_profile_frame(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height, vxo_profile* out);


/* ---- traversal grid of the CUDA path, restated on the host (oracle/vxo_trav.c; not part of the reference) ------------
 * word of one cell / of the whole grid (returns the number of values that cannot be encoded), the reference value of a
 * word, castRay and vxo_render on a traversal grid (stats9: fast steps, checked steps, jumps per ray kind 0/1/2). */
int32_t vxo_trav_word(const int32_t* vox, vxo_dims g, int x, int y, int z);
int64_t vxo_trav_build(const int32_t* vox, vxo_dims g, int32_t* trav);
int32_t vxo_trav_canonical(int32_t w);
int32_t vxo_trav_cast_ray(const int32_t* trav, vxo_dims g, const float start[3], const float dir[3], int32_t dist, vxo_ray_out* out);
void vxo_trav_render(const int32_t* trav, vxo_dims g, const vxo_frame* f, int width, int height, int y0, int y1,
                     uint8_t* rgba8, int32_t* hit_index, uint16_t* steps, uint32_t* occl_mask, uint32_t* cast_mask,
                     uint64_t counters[5], uint64_t stats9[9]);

#ifdef __cplusplus
}
#endif
/* ---- host statement of how the CUDA path's production kernels organise the lighting (vxo_wide.c; test infrastructure) ----
 * skip_dark: rays toward lights the surface faces away from are not traced; wide: the active lights are evaluated in two halves,
 * the second without the early-out, and combined in slot order.  Must equal vxo_render's float frame bit for bit. */
void vxo_wide_render(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height, int skip_dark, int wide, float* rgba_f32);

#endif
