/*
 * vxo_internal.h -- shared by the oracle's translation units (TEST INFRASTRUCTURE ONLY, see vxo.h): the shader's mutable
 * globals, the conventions GLSL leaves open (vxo.h), and the hook through which the analysis units run fshader.glsl's
 * main() with their own castRay.
 */
#ifndef VXO_INTERNAL_H
#define VXO_INTERNAL_H
#include "vxo.h"
#include <limits.h>

/* constants: fshader.glsl:3-10 */
#define VXO_RENDER_DIST 384

/* the shader's mutable globals, fshader.glsl:28-31 */
typedef struct {
    float hitPos[3];
    float hitNormal[3];
    float stepCount;
    uint64_t fetches;           /* instrumentation: iterations of every castRay call */
} vxo_shader_state;

/* float -> int: truncation; NaN / out of range -> INT_MIN (see vxo.h conventions) */
static inline int vxo_f2i(float x) {
    return (x >= -2147483648.0f && x < 2147483648.0f) ? (int)x : INT_MIN;
}
static inline float vxo_fsign(float x) { return (float)((0.0f < x) - (x < 0.0f)); }     /* GLM sign(): (0<x) - (x<0) */
static inline float vxo_fmax0(float b) { return (0.0f < b) ? b : 0.0f; }                /* max(0, b) */

/* a castRay (fshader.glsl:59-129) as main() calls it; kind: 0 primary, 1 global-light, 2 local-light ray; dark: the
 * surface faces away from the light (N.L <= 0).  kind / dark / user are for the analysis units only. */
typedef int32_t (*vxo_cast_fn)(void* user, const int32_t* vox, vxo_dims g, vxo_shader_state* st, float sx, float sy, float sz,
                               float rx, float ry, float rz, int32_t dist, int kind, int dark);

/* fshader.glsl:131-190 for one pixel with the given castRay (vxo_shade_pixel = this with the plain restatement) */
void vxo_shade_pixel_with(vxo_cast_fn castRay, void* user, const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height,
                          int px, int py, float rgba[4], int32_t* hit_index, float* steps,
                          uint32_t* occl_mask, uint32_t* cast_mask, float hit_pos[3], float hit_normal[3],
                          uint64_t counters[5]);
#endif
