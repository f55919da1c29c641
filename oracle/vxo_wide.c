/*
 * vxo_wide.c -- CPU statement of how the CUDA path's PRODUCTION kernels organise fshader.glsl's lighting (voxel-rt_b200/csrc/
 * kernels.cuh shade_pixel, shade_wide_first / shade_wide_second).  TEST INFRASTRUCTURE ONLY (see vxo.h).  Not part of the
 * restatement of the reference: the reference walks slots 0..15 in one thread.  It exists so that the three re-organisations
 * are CHECKED against the oracle on the CPU, bit for bit on whole frames (tests/test_wide_oracle.py), before any GPU runs them:
 *   1. only the ACTIVE light slots are walked (compacted, slot order kept); the overbright clamp the reference tests at the
 *      top of every iteration -- inactive slots included -- is applied once more after the loop when the last active slot is
 *      not slot 15;
 *   2. skip_dark: a ray toward a light the surface faces away from (N.L <= 0) is not traced -- its term is
 *      weight * max(0, N.L) * falloff = +-0 whether occluded or not (fshader.glsl:155,177) -- unless the weight is not finite;
 *   3. wide: the active lights are split in two halves; the first half is walked with the early-out (fshader.glsl:161-164),
 *      the second half is EVALUATED without it (what a light adds does not depend on the multiplier) and its terms are applied
 *      afterwards in slot order with the clamp test in front of each.
 * Primary ray, castRay and the colour are the oracle's own code paths (castRay through vxo_castray_body.inc).
 */
#include "vxo.h"
#include "vxo_internal.h"
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOCAL_LIGHT_DIST 64
#define AMBIENT 0.4f
#define DIFFUSE 0.8f
#define MAX_OVERBRIGHT 1.25f

#define VXO_CAST_RAY_NAME cast_ray_wide_copy
#define VXO_HOOK_ARGS
#define VXO_HOOK_START()
#define VXO_HOOK_ITER()
#define VXO_HOOK_JUMP()
#define VXO_HOOK_END(outcome)
#include "vxo_castray_body.inc"

static float dot3(const float a[3], const float b[3]) {             /* GLM: (a.x*b.x + a.y*b.y) + a.z*b.z */
    float t0 = a[0] * b[0], t1 = a[1] * b[1], t2 = a[2] * b[2];
    return (t0 + t1) + t2;
}
static void normalize3(const float v[3], float out[3]) {            /* GLM: v * (1 / sqrt(dot(v, v))) */
    float inv = 1.0f / sqrtf(dot3(v, v));
    out[0] = v[0] * inv; out[1] = v[1] * inv; out[2] = v[2] * inv;
}

/* fshader.glsl:166-179 for one light: 1 when the light adds *c to the multiplier (kernels.cuh light_term) */
static int light_term(const int32_t* vox, vxo_dims g, vxo_shader_state* st, const float L[4], const float fhp[3], const float fhn[3],
                      int skip_dark, float* c) {
    float d[3] = {L[0] - fhp[0], L[1] - fhp[1], L[2] - fhp[2]};
    const int lit = dot3(fhn, d) > 0.0f || !(fabsf(L[3]) <= 3.0e38f);
    if (skip_dark && !lit) return 0;
    float lld = sqrtf(dot3(d, d));                                  /* :168 */
    if (!(lld <= (float)LOCAL_LIGHT_DIST)) return 0;                /* :171 */
    float inv = 1.0f / lld;                                         /* :173, the same 1 / sqrt(dot) */
    float tll[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
    if (cast_ray_wide_copy(vox, g, st, fhp[0] + tll[0] * 0.001f, fhp[1] + tll[1] * 0.001f, fhp[2] + tll[2] * 0.001f,
                           tll[0], tll[1], tll[2], vxo_f2i(lld + 1.0f)) != -1) return 0;                /* :175 */
    float fall = ((float)LOCAL_LIGHT_DIST - lld) * (1.0f / (float)LOCAL_LIGHT_DIST);                    /* an exact scaling == / 64 */
    *c = (L[3] * vxo_fmax0(dot3(fhn, tll))) * fall;                 /* :177 */
    return 1;
}

/* fshader.glsl:131-190 for one pixel, lighting organised like the production kernels */
static void shade_pixel_cuda_style(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height, int px, int py,
                                   int skip_dark, int wide, float rgba[4]) {
    vxo_shader_state st; memset(&st, 0, sizeof st);
    float vx = ((float)px + 0.5f) / (float)width * 2.0f - 1.0f;
    float vy = ((float)py + 0.5f) / (float)height * 2.0f - 1.0f;
    float col[4] = {(float)0.6, (float)0.7, (float)0.8, 1.0f};      /* :133 */
    float v[3] = {vx * f->aspect, vy, 1.0f}, rd[3];
    normalize3(v, rd);                                              /* :136 */
    const float* m = f->rotate;
    float rot[3];
    for (int r = 0; r < 3; r++) {                                   /* :137 */
        float mul0 = m[0 + r] * rd[0], mul1 = m[4 + r] * rd[1], mul2 = m[8 + r] * rd[2], mul3 = m[12 + r] * 0.0f;
        rot[r] = (mul0 + mul1) + (mul2 + mul3);
    }
    int32_t idx = cast_ray_wide_copy(vox, g, &st, f->cam_pos[0], f->cam_pos[1], f->cam_pos[2], rot[0], rot[1], rot[2], VXO_RENDER_DIST);   /* :139 */
    float fhp[3] = {st.hitPos[0], st.hitPos[1], st.hitPos[2]};
    float fhn[3] = {st.hitNormal[0], st.hitNormal[1], st.hitNormal[2]};
    if (f->view_depth_field == 1) {                                 /* :143-145 */
        float gr = st.stepCount / 100.0f;
        col[0] = gr; col[1] = gr; col[2] = gr; col[3] = 1.0f;
    } else if (idx != -1 && vox[idx] >= 0) {                        /* :152 */
        /* the active slots, compacted, order kept (kernels.cuh compact_lights) */
        float L[VXO_MAX_LOCAL_LIGHTS][4]; int slot[VXO_MAX_LOCAL_LIGHTS]; int nact = 0;
        for (int i = 0; i < VXO_MAX_LOCAL_LIGHTS; i++)
            if (f->lights[i][0] >= 0 && f->lights[i][1] >= 0 && f->lights[i][2] >= 0) {                /* :167 */
                for (int k = 0; k < 4; k++) L[nact][k] = f->lights[i][k];
                slot[nact++] = i;
            }
        float multiplier = AMBIENT;                                 /* :149 */
        {   /* :147, :154-155 global light; the sign test uses the unnormalised direction */
            float tl[3] = {f->light_pos[0] - fhp[0], f->light_pos[1] - fhp[1], f->light_pos[2] - fhp[2]};
            if (dot3(fhn, tl) > 0.0f || !skip_dark) {
                float toLight[3];
                normalize3(tl, toLight);
                if (cast_ray_wide_copy(vox, g, &st, fhp[0] + toLight[0] * 0.001f, fhp[1] + toLight[1] * 0.001f, fhp[2] + toLight[2] * 0.001f,
                                       toLight[0], toLight[1], toLight[2], VXO_RENDER_DIST) == -1)
                    multiplier = multiplier + DIFFUSE * vxo_fmax0(dot3(fhn, toLight));
            }
        }
        const int n0 = wide ? (nact >> 1) : nact;
        int last_slot = -1, broke = 0;
        for (int k = 0; k < n0; k++) {                              /* first half (or all): with the early-out */
            if (multiplier >= MAX_OVERBRIGHT) { multiplier = MAX_OVERBRIGHT; broke = 1; break; }       /* :161-164 */
            last_slot = slot[k];
            float c;
            if (light_term(vox, g, &st, L[k], fhp, fhn, skip_dark, &c)) multiplier = multiplier + c;
        }
        if (wide) {
            float c2[VXO_MAX_LOCAL_LIGHTS]; int has[VXO_MAX_LOCAL_LIGHTS];
            for (int k = n0; k < nact; k++) {                       /* second half: every light, no early-out (the other thread) */
                c2[k] = 0.0f;
                has[k] = light_term(vox, g, &st, L[k], fhp, fhn, skip_dark, &c2[k]);
            }
            if (!broke)
                for (int k = n0; k < nact; k++) {                   /* applied in slot order, the clamp test in front of each */
                    if (multiplier >= MAX_OVERBRIGHT) { multiplier = MAX_OVERBRIGHT; broke = 1; break; }
                    last_slot = slot[k];
                    if (has[k]) multiplier = multiplier + c2[k];
                }
        }
        /* the reference also tests the clamp at the top of the inactive iterations behind the last active slot */
        if (!broke && last_slot < VXO_MAX_LOCAL_LIGHTS - 1 && multiplier >= MAX_OVERBRIGHT) multiplier = MAX_OVERBRIGHT;
        int32_t c = vox[idx];                                       /* :184-187 */
        col[0] = (float)((c & 0x00FF0000) >> 16) / 255.0f * multiplier;
        col[1] = (float)((c & 0x0000FF00) >> 8) / 255.0f * multiplier;
        col[2] = (float)(c & 0x000000FF) / 255.0f * multiplier;
        col[3] = 1.0f;
    }
    for (int k = 0; k < 4; k++) rgba[k] = col[k];
}

/* whole frame, float RGBA (compared bit for bit with vxo_render's rgba_f32) */
void vxo_wide_render(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height, int skip_dark, int wide, float* rgba_f32) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int py = 0; py < height; py++)
        for (int px = 0; px < width; px++)
            shade_pixel_cuda_style(vox, g, f, width, height, px, py, skip_dark, wide, rgba_f32 + 4 * ((size_t)py * (size_t)width + (size_t)px));
}
