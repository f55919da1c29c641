/*
 * vxo_trav.c -- CPU statement of the TRAVERSAL GRID of the CUDA path (voxel-rt_b200/csrc/trav.cuh) and of castRay on it.
 * TEST INFRASTRUCTURE ONLY (see vxo.h).  Not part of the restatement of the reference: the reference has no such structure.
 * It exists so that (1) the device-built traversal grid can be compared word for word with a host rebuild, and (2) the
 * argument that runs of steps without index arithmetic / range test / load preserve castRay's result exactly is CHECKED on
 * whole frames on the CPU (tests/test_trav_oracle.py: vxo_trav_render == vxo_render, bit for bit) before any GPU is involved.
 *
 * Traversal word of a cell with reference value v (render.cpp:189-196 layout, same index):
 *   v >= 0 (solid) or v a depth-field jump (negative, != -1)  ->  v itself
 *   v == -1 (empty, no jump: the band the depth field leaves next to surfaces, render.cpp:240-251)
 *          ->  0x80000000 | for each travel quadrant q = (stepx > 0) | (stepz > 0) << 1, 7 bits at bit 7q:
 *                 K (4 bits): every cell of the same y layer at quadrant offsets (a, b), a, b >= 0, 1 <= a + b <= K, is inside
 *                             the grid and holds -1                                                     (0 .. 15)
 *                 U (3 bits): additionally every cell of the layer ABOVE at offsets a + b <= U - 1 is inside the grid and
 *                             holds -1 (U <= K)                                                          (0 .. 7)
 * Band words are the negative ints with bit 30 clear; depth-field jumps are -sqrt(k) <= -2.0f, whose bit 30 is set.  A grid
 * that holds some OTHER negative value with bit 30 clear cannot be encoded (vxo_trav_build returns their count; the CUDA
 * path then renders from the reference-layout grid with its plain kernels).
 *
 * castRay on it: a step moves ONE cell along ONE axis whatever the float state says (ties included), so while K >= 1 an x / z
 * step needs no index arithmetic, range test or load -- its target is a -1 cell inside the grid, for which fshader.glsl:105-125
 * does nothing -- and one upward y step is covered by U.  Every float operation of fshader.glsl:83-104 still runs, in order.
 */
#include "vxo.h"
#include "vxo_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TRAV_K_CAP 15
#define TRAV_U_CAP 7
#define TRAV_BIG (1 << 20)

static int trav_is_band(int32_t w) { return w < 0 && !(w & 0x40000000); }

static int is_free(const int32_t* vox, vxo_dims g, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= g.w || y >= g.h || z >= g.d) return 0;
    return vox[(size_t)x + (size_t)g.w * y + (size_t)g.w * g.h * z] == -1;
}

/* min(cap, quadrant-L1 distance from (x,y,z) to the nearest cell of layer y that is not a -1 cell inside the grid), the cell
 * itself included */
static int quad_dist(const int32_t* vox, vxo_dims g, int x, int y, int z, int sx, int sz, int cap) {
    int best = cap;
    for (int b = 0; b < best; b++)
        for (int a = 0; a + b < best; a++)
            if (!is_free(vox, g, x + sx * a, y, z + sz * b)) { best = a + b; break; }
    return best;
}

int32_t vxo_trav_word(const int32_t* vox, vxo_dims g, int x, int y, int z) {
    const int32_t v = vox[(size_t)x + (size_t)g.w * y + (size_t)g.w * g.h * z];
    if (v != -1) return v;
    uint32_t w = 0x80000000u;
    for (int q = 0; q < 4; q++) {
        const int sx = (q & 1) ? 1 : -1, sz = (q & 2) ? 1 : -1;
        const int K = quad_dist(vox, g, x, y, z, sx, sz, TRAV_K_CAP + 1) - 1;
        int U = quad_dist(vox, g, x, y + 1, z, sx, sz, TRAV_U_CAP);
        if (U > K) U = K;
        w |= (uint32_t)(K | (U << 4)) << (7 * q);
    }
    return (int32_t)w;
}

/* whole grid; returns the number of cells whose value cannot be encoded (negative, != -1, bit 30 clear) */
int64_t vxo_trav_build(const int32_t* vox, vxo_dims g, int32_t* trav) {
    int64_t bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
#endif
    for (int z = 0; z < g.d; z++)
        for (int y = 0; y < g.h; y++)
            for (int x = 0; x < g.w; x++) {
                const size_t i = (size_t)x + (size_t)g.w * y + (size_t)g.w * g.h * z;
                const int32_t v = vox[i];
                if (v < 0 && v != -1 && !(v & 0x40000000)) bad++;
                trav[i] = vxo_trav_word(vox, g, x, y, z);
            }
    return bad;
}

/* reference value of a traversal word */
int32_t vxo_trav_canonical(int32_t w) { return trav_is_band(w) ? -1 : w; }

typedef struct { uint64_t steps_fast[3], steps_checked[3], jumps[3]; } trav_stats;

/* fshader.glsl:59-129 on the traversal grid (the algorithm of ray.cuh's traversal loops, without their scheduling) */
static int32_t cast_ray_trav(const int32_t* trav, vxo_dims g, vxo_shader_state* st,
                             float sx, float sy, float sz, float rx, float ry, float rz, int32_t dist, trav_stats* ts, int kind) {
    int32_t cx = vxo_f2i(sx), cy = vxo_f2i(sy), cz = vxo_f2i(sz);   /* :64 */
    int32_t fColorIndex = -1, tempIndex = -1;
    int32_t stepx = vxo_f2i(vxo_fsign(rx)), stepy = vxo_f2i(vxo_fsign(ry)), stepz = vxo_f2i(vxo_fsign(rz));   /* :71 */
    int32_t fwx = (stepx > 0), fwy = (stepy > 0), fwz = (stepz > 0);                  /* :72 */
    float dx = 1.0f / fabsf(rx + 0.000001f);                        /* :74-76 */
    float dy = 1.0f / fabsf(ry + 0.000001f);
    float dz = 1.0f / fabsf(rz + 0.000001f);
    float ix = ((float)(int32_t)((uint32_t)cx + (uint32_t)fwx) - sx) / rx;          /* :79 */
    float iy = ((float)(int32_t)((uint32_t)cy + (uint32_t)fwy) - sy) / ry;
    float iz = ((float)(int32_t)((uint32_t)cz + (uint32_t)fwz) - sz) / rz;
    float currDist = 0.0f, distTravelled = 0.0f;
    const int shift = 7 * ((stepx > 0) | ((stepz > 0) << 1));       /* the travel quadrant's 7 bits */
    const int up = stepy > 0;
    int K = 0, E = TRAV_BIG;                                        /* K: x / z steps known to land on -1 cells; K - E: U */
    while (distTravelled < (float)dist && distTravelled < (float)VXO_RENDER_DIST) {  /* :83 */
        st->stepCount = st->stepCount + 1.0f;                       /* :84 */
        st->fetches++;
        const int bx = ix < iy && ix < iz, by = !bx && (iy < ix && iy < iz);         /* :87, :93 */
        if (K >= 1 && (!by || K - E >= 1)) {
            /* a step of a run: :85-104 exactly, then nothing -- the target is a -1 cell inside the grid (:105-125 do nothing) */
            distTravelled = distTravelled + 1.0f;
            if (bx) {
                currDist = ix; cx = (int32_t)((uint32_t)cx + (uint32_t)stepx); ix = ix + dx;
                st->hitNormal[0] = (float)(-stepx); st->hitNormal[1] = 0.0f; st->hitNormal[2] = 0.0f;
                K--;
            } else if (by) {
                currDist = iy; cy = (int32_t)((uint32_t)cy + (uint32_t)stepy); iy = iy + dy;
                st->hitNormal[0] = 0.0f; st->hitNormal[1] = (float)(-stepy); st->hitNormal[2] = 0.0f;
                K = K - E - 1; E = TRAV_BIG;                        /* the layer above: U - 1 steps are left, nothing known beyond */
            } else {
                currDist = iz; cz = (int32_t)((uint32_t)cz + (uint32_t)stepz); iz = iz + dz;
                st->hitNormal[0] = 0.0f; st->hitNormal[1] = 0.0f; st->hitNormal[2] = (float)(-stepz);
                K--;
            }
            if (ts) ts->steps_fast[kind]++;
            continue;
        }
        if (ts) ts->steps_checked[kind]++;
        distTravelled = distTravelled + 1.0f;                       /* :85 */
        if (bx) {                                                   /* :87-92 */
            currDist = ix; cx = (int32_t)((uint32_t)cx + (uint32_t)stepx); ix = ix + dx;
            st->hitNormal[0] = (float)(-stepx); st->hitNormal[1] = 0.0f; st->hitNormal[2] = 0.0f;
        } else if (by) {                                            /* :93-98 */
            currDist = iy; cy = (int32_t)((uint32_t)cy + (uint32_t)stepy); iy = iy + dy;
            st->hitNormal[0] = 0.0f; st->hitNormal[1] = (float)(-stepy); st->hitNormal[2] = 0.0f;
        } else {                                                    /* :99-104 (ties land here) */
            currDist = iz; cz = (int32_t)((uint32_t)cz + (uint32_t)stepz); iz = iz + dz;
            st->hitNormal[0] = 0.0f; st->hitNormal[1] = 0.0f; st->hitNormal[2] = (float)(-stepz);
        }
        tempIndex = vxo_shader_index(g, cx, cy, cz);                /* :105 */
        K = 0; E = TRAV_BIG;
        if (tempIndex < 0) break;                                   /* :123-125 */
        const int32_t w = trav[tempIndex];
        if (w >= 0) {                                               /* :108-112 */
            st->hitPos[0] = rx * currDist + sx;
            st->hitPos[1] = ry * currDist + sy;
            st->hitPos[2] = rz * currDist + sz;
            fColorIndex = tempIndex;
            break;
        } else if (trav_is_band(w)) {                               /* the reference's -1: nothing happens; the word says what lies ahead */
            const int f = (int)(((uint32_t)w >> shift) & 127u);
            K = f & 15;
            const int U = up ? (f >> 4) : 0;
            E = U ? K - U : TRAV_BIG;
        } else {                                                    /* :114-121 */
            float bits; memcpy(&bits, &w, 4);
            float toJump = -bits;
            distTravelled = distTravelled + toJump;
            currDist = currDist + toJump;
            sx = rx * currDist + sx; sy = ry * currDist + sy; sz = rz * currDist + sz;
            cx = vxo_f2i(sx); cy = vxo_f2i(sy); cz = vxo_f2i(sz);
            ix = ((float)(int32_t)((uint32_t)cx + (uint32_t)fwx) - sx) / rx;
            iy = ((float)(int32_t)((uint32_t)cy + (uint32_t)fwy) - sy) / ry;
            iz = ((float)(int32_t)((uint32_t)cz + (uint32_t)fwz) - sz) / rz;
            if (ts) ts->jumps[kind]++;
        }
    }
    return fColorIndex;
}

static int32_t cast_adapter(void* user, const int32_t* vox, vxo_dims g, vxo_shader_state* st, float sx, float sy, float sz,
                            float rx, float ry, float rz, int32_t dist, int kind, int dark) {
    (void)dark;
    return cast_ray_trav(vox, g, st, sx, sy, sz, rx, ry, rz, dist, (trav_stats*)user, kind);
}

int32_t vxo_trav_cast_ray(const int32_t* trav, vxo_dims g, const float start[3], const float dir[3], int32_t dist, vxo_ray_out* out) {
    vxo_shader_state st; memset(&st, 0, sizeof st);
    int32_t r = cast_ray_trav(trav, g, &st, start[0], start[1], start[2], dir[0], dir[1], dir[2], dist, NULL, 0);
    if (out) {
        for (int k = 0; k < 3; k++) { out->hit_pos[k] = st.hitPos[k]; out->hit_normal[k] = st.hitNormal[k]; }
        out->steps = st.stepCount; out->hit_set = (r >= 0);
    }
    return r;
}

static uint8_t unorm8(float c) {
    if (!(c > 0.0f)) c = 0.0f;
    if (c > 1.0f) c = 1.0f;
    return (uint8_t)(int)(c * 255.0f + 0.5f);
}

/* vxo_render on the traversal grid; stats9: fast steps, checked steps, jumps for ray kinds 0 / 1 / 2 (accumulated) */
void vxo_trav_render(const int32_t* trav, vxo_dims g, const vxo_frame* f, int width, int height, int y0, int y1,
                     uint8_t* rgba8, int32_t* hit_index, uint16_t* steps, uint32_t* occl_mask, uint32_t* cast_mask,
                     uint64_t counters[5], uint64_t stats9[9]) {
    uint64_t tot[5] = {0, 0, 0, 0, 0};
    trav_stats tsum; memset(&tsum, 0, sizeof tsum);
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        uint64_t loc[5] = {0, 0, 0, 0, 0};
        trav_stats ts; memset(&ts, 0, sizeof ts);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int py = y0; py < y1; py++)
            for (int px = 0; px < width; px++) {
                float rgba[4]; int32_t hi; float sc; uint32_t om, cm;
                vxo_shade_pixel_with(cast_adapter, &ts, trav, g, f, width, height, px, py, rgba, &hi, &sc, &om, &cm, NULL, NULL, loc);
                size_t p = (size_t)py * (size_t)width + (size_t)px;
                if (rgba8) for (int k = 0; k < 4; k++) rgba8[4 * p + k] = unorm8(rgba[k]);
                if (hit_index) hit_index[p] = hi;
                if (steps) steps[p] = (uint16_t)(sc > 65535.0f ? 65535 : (int)sc);
                if (occl_mask) occl_mask[p] = om;
                if (cast_mask) cast_mask[p] = cm;
            }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (int k = 0; k < 5; k++) tot[k] += loc[k];
            for (int k = 0; k < 3; k++) { tsum.steps_fast[k] += ts.steps_fast[k]; tsum.steps_checked[k] += ts.steps_checked[k]; tsum.jumps[k] += ts.jumps[k]; }
        }
    }
    if (counters) for (int k = 0; k < 5; k++) counters[k] += tot[k];
    if (stats9) for (int k = 0; k < 3; k++) { stats9[k] += tsum.steps_fast[k]; stats9[3 + k] += tsum.steps_checked[k]; stats9[6 + k] += tsum.jumps[k]; }
}

/* ---- analysis only: how long are the runs under other caps / encodings?  (scripts/where_iterations_go.py --trav-study) -------
 * The promise of every -1 cell is kept UNPACKED (one byte per quadrant for K, one for U), so that caps beyond what fits the 30
 * payload bits of a band word can be tried.  ucode7: U is stored in 3 bits as {0..6, 7 = "U equals K"} (anything between 7 and K - 1
 * is stored as 6).  Rays from surfaces that face away from their light are left out (the production kernels do not trace them). */
typedef struct { const uint8_t* K; const uint8_t* U; trav_stats ts; uint64_t runs[3]; } study_ctx;

static int32_t cast_ray_study(const int32_t* vox, vxo_dims g, vxo_shader_state* st, float sx, float sy, float sz, float rx, float ry, float rz,
                              int32_t dist, study_ctx* S, int kind) {
    int32_t cx = vxo_f2i(sx), cy = vxo_f2i(sy), cz = vxo_f2i(sz);
    int32_t fColorIndex = -1, tempIndex = -1;
    int32_t stepx = vxo_f2i(vxo_fsign(rx)), stepy = vxo_f2i(vxo_fsign(ry)), stepz = vxo_f2i(vxo_fsign(rz));
    int32_t fwx = (stepx > 0), fwy = (stepy > 0), fwz = (stepz > 0);
    float dx = 1.0f / fabsf(rx + 0.000001f), dy = 1.0f / fabsf(ry + 0.000001f), dz = 1.0f / fabsf(rz + 0.000001f);
    float ix = ((float)(int32_t)((uint32_t)cx + (uint32_t)fwx) - sx) / rx;
    float iy = ((float)(int32_t)((uint32_t)cy + (uint32_t)fwy) - sy) / ry;
    float iz = ((float)(int32_t)((uint32_t)cz + (uint32_t)fwz) - sz) / rz;
    float currDist = 0.0f, distTravelled = 0.0f;
    const int q = (stepx > 0) | ((stepz > 0) << 1), up = stepy > 0;
    int K = 0, E = TRAV_BIG;
    while (distTravelled < (float)dist && distTravelled < (float)VXO_RENDER_DIST) {
        st->stepCount = st->stepCount + 1.0f;
        st->fetches++;
        const int bx = ix < iy && ix < iz, by = !bx && (iy < ix && iy < iz);
        const int in_run = K >= 1 && (!by || K - E >= 1);
        distTravelled = distTravelled + 1.0f;
        if (bx) { currDist = ix; cx = (int32_t)((uint32_t)cx + (uint32_t)stepx); ix = ix + dx; }
        else if (by) { currDist = iy; cy = (int32_t)((uint32_t)cy + (uint32_t)stepy); iy = iy + dy; }
        else { currDist = iz; cz = (int32_t)((uint32_t)cz + (uint32_t)stepz); iz = iz + dz; }
        st->hitNormal[0] = bx ? (float)(-stepx) : 0.0f; st->hitNormal[1] = by ? (float)(-stepy) : 0.0f; st->hitNormal[2] = (!bx && !by) ? (float)(-stepz) : 0.0f;
        if (in_run) {
            if (by) { K = K - E - 1; E = TRAV_BIG; } else K--;
            S->ts.steps_fast[kind]++;
            continue;
        }
        S->ts.steps_checked[kind]++;
        tempIndex = vxo_shader_index(g, cx, cy, cz);
        K = 0; E = TRAV_BIG;
        if (tempIndex < 0) break;
        const int32_t w = vox[tempIndex];
        if (w >= 0) {
            st->hitPos[0] = rx * currDist + sx; st->hitPos[1] = ry * currDist + sy; st->hitPos[2] = rz * currDist + sz;
            fColorIndex = tempIndex;
            break;
        } else if (w == -1) {
            K = S->K[(size_t)tempIndex * 4 + q];
            const int U = up ? S->U[(size_t)tempIndex * 4 + q] : 0;
            E = K - U;
            if (K >= 1) S->runs[kind]++;
        } else {
            float bits; memcpy(&bits, &w, 4);
            float toJump = -bits;
            distTravelled = distTravelled + toJump;
            currDist = currDist + toJump;
            sx = rx * currDist + sx; sy = ry * currDist + sy; sz = rz * currDist + sz;
            cx = vxo_f2i(sx); cy = vxo_f2i(sy); cz = vxo_f2i(sz);
            ix = ((float)(int32_t)((uint32_t)cx + (uint32_t)fwx) - sx) / rx;
            iy = ((float)(int32_t)((uint32_t)cy + (uint32_t)fwy) - sy) / ry;
            iz = ((float)(int32_t)((uint32_t)cz + (uint32_t)fwz) - sz) / rz;
            S->ts.jumps[kind]++;
        }
    }
    return fColorIndex;
}

/* the plain restatement for the rays the study leaves out */
#define VXO_CAST_RAY_NAME cast_ray_plain_copy
#define VXO_HOOK_ARGS
#define VXO_HOOK_START()
#define VXO_HOOK_ITER()
#define VXO_HOOK_JUMP()
#define VXO_HOOK_END(outcome)
#include "vxo_castray_body.inc"

static int32_t study_adapter(void* user, const int32_t* vox, vxo_dims g, vxo_shader_state* st, float sx, float sy, float sz,
                             float rx, float ry, float rz, int32_t dist, int kind, int dark) {
    if (dark) return cast_ray_plain_copy(vox, g, st, sx, sy, sz, rx, ry, rz, dist);
    return cast_ray_study(vox, g, st, sx, sy, sz, rx, ry, rz, dist, (study_ctx*)user, kind);
}

/* out12: run steps, checked steps, jumps, runs started -- per ray kind 0 / 1 / 2; returns the number of pixels that differ from
 * vxo_render (must be 0) */
int64_t vxo_trav_study(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height, int kcap, int ucap, int ucode7, uint64_t out12[12]) {
    const size_t n = (size_t)g.w * g.h * g.d;
    uint8_t* Kb = (uint8_t*)calloc(n * 4, 1);
    uint8_t* Ub = (uint8_t*)calloc(n * 4, 1);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int z = 0; z < g.d; z++)
        for (int y = 0; y < g.h; y++)
            for (int x = 0; x < g.w; x++) {
                const size_t i = (size_t)x + (size_t)g.w * y + (size_t)g.w * g.h * z;
                if (vox[i] != -1) continue;
                for (int q = 0; q < 4; q++) {
                    const int sx = (q & 1) ? 1 : -1, sz = (q & 2) ? 1 : -1;
                    const int K = quad_dist(vox, g, x, y, z, sx, sz, kcap + 1) - 1;
                    int U = quad_dist(vox, g, x, y + 1, z, sx, sz, ucode7 ? kcap : ucap);
                    if (U > K) U = K;
                    if (ucode7 && U < K && U > 6) U = 6;
                    Kb[i * 4 + q] = (uint8_t)K; Ub[i * 4 + q] = (uint8_t)U;
                }
            }
    int64_t bad = 0;
    uint64_t tot[12] = {0};
#ifdef _OPENMP
#pragma omp parallel reduction(+ : bad)
#endif
    {
        study_ctx S; memset(&S, 0, sizeof S); S.K = Kb; S.U = Ub;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int py = 0; py < height; py++)
            for (int px = 0; px < width; px++) {
                float a[4], b[4];
                vxo_shade_pixel_with(study_adapter, &S, vox, g, f, width, height, px, py, a, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
                vxo_shade_pixel(vox, g, f, width, height, px, py, b, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
                if (memcmp(a, b, sizeof a)) bad++;
            }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int k = 0; k < 3; k++) { tot[k] += S.ts.steps_fast[k]; tot[3 + k] += S.ts.steps_checked[k]; tot[6 + k] += S.ts.jumps[k]; tot[9 + k] += S.runs[k]; }
    }
    for (int k = 0; k < 12; k++) out12[k] = tot[k];
    free(Kb); free(Ub);
    return bad;
}
