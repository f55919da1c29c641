/*
 * vxo_analysis.c -- ANALYSIS ONLY (scripts/where_iterations_go.py), TEST INFRASTRUCTURE (see vxo.h): where do castRay's
 * iterations go?  Runs the oracle's main() (vxo_shade_pixel_with) with an instrumented instance of the shared castRay body
 * (vxo_castray_body.inc): the hooks only count.  For every ray: its kind, how it ended, its iterations, how many of them were
 * depth-field jumps, and how many came AFTER the first moment at which its cell lay beyond every grid row that holds a solid
 * voxel in its direction of travel (tested at the start and after each jump, like the CUDA path's occupancy-summary culling).
 */
#include "vxo.h"
#include "vxo_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { vxo_profile* prof; int ymin, ymax; } ana_ctx;

static void prof_ray(vxo_profile* p, int kind, int dark, int outcome, uint64_t it, uint64_t jumps, uint64_t cull_it) {
    vxo_profile_cell* c = &p->cell[kind][outcome];
    c->rays++; c->iterations += it; c->jumps += jumps;
    if (cull_it < it) { c->rays_culled++; c->iterations_after_cull += it - cull_it; }
    if (dark) { c->rays_dark++; c->iterations_dark += it; if (cull_it < it) c->iterations_dark_after_cull += it - cull_it; }
    if (it > p->longest[kind]) p->longest[kind] = it;
}

#define VXO_BEYOND() ((stepy > 0 && cy > A->ymax) || (stepy < 0 && cy < A->ymin))
#define VXO_CAST_RAY_NAME cast_ray_profiled
#define VXO_HOOK_ARGS , const ana_ctx* A, int kind, int dark
#define VXO_HOOK_START() uint64_t p_it = 0, p_jumps = 0, p_cull = UINT64_MAX; if (VXO_BEYOND()) p_cull = 0
#define VXO_HOOK_ITER() p_it++
#define VXO_HOOK_JUMP() do { p_jumps++; if (p_cull == UINT64_MAX && VXO_BEYOND()) p_cull = p_it; } while (0)
#define VXO_HOOK_END(outcome) prof_ray(A->prof, kind, dark, outcome, p_it, p_jumps, p_cull)
#include "vxo_castray_body.inc"

static int32_t cast_adapter(void* user, const int32_t* vox, vxo_dims g, vxo_shader_state* st, float sx, float sy, float sz,
                            float rx, float ry, float rz, int32_t dist, int kind, int dark) {
    return cast_ray_profiled(vox, g, st, sx, sy, sz, rx, ry, rz, dist, (const ana_ctx*)user, kind, dark);
}

/* vxo_render's loop with the per-ray profile switched on; out is accumulated over all threads */
void vxo_profile_frame(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height, vxo_profile* out) {
    int ymin = INT_MAX, ymax = INT_MIN;
    for (int z = 0; z < g.d; z++)
        for (int y = 0; y < g.h; y++) {
            if (y >= ymin && y <= ymax) continue;
            const int32_t* row = vox + ((size_t)z * g.h + y) * g.w;
            for (int x = 0; x < g.w; x++) if (row[x] >= 0) { if (y < ymin) ymin = y; if (y > ymax) ymax = y; break; }
        }
    memset(out, 0, sizeof *out);
    out->ymin = ymin; out->ymax = ymax;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        vxo_profile loc; memset(&loc, 0, sizeof loc);
        ana_ctx A = {&loc, ymin, ymax};
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int py = 0; py < height; py++)
            for (int px = 0; px < width; px++) {
                float rgba[4];
                vxo_shade_pixel_with(cast_adapter, &A, vox, g, f, width, height, px, py, rgba, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
            }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int k = 0; k < 3; k++) {
            for (int o = 0; o < 3; o++) {
                vxo_profile_cell* a = &out->cell[k][o]; const vxo_profile_cell* b = &loc.cell[k][o];
                a->rays += b->rays; a->iterations += b->iterations; a->jumps += b->jumps; a->rays_culled += b->rays_culled;
                a->iterations_after_cull += b->iterations_after_cull; a->rays_dark += b->rays_dark;
                a->iterations_dark += b->iterations_dark; a->iterations_dark_after_cull += b->iterations_dark_after_cull;
            }
            if (loc.longest[k] > out->longest[k]) out->longest[k] = loc.longest[k];
        }
    }
}
