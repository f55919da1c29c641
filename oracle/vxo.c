/*
 * vxo.c -- CPU ORACLE (test infrastructure only, see vxo.h).  Plain-C restatement of the
 * reference's per-pixel path and of the host code that feeds it.  Compile with
 *   gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fwrapv -fopenmp
 * so that every float operation is one correctly rounded IEEE-754 binary32 operation in source
 * order and int arithmetic wraps.  Citations are file:line under the reference's src/.
 */
#include "vxo.h"
#include "vxo_internal.h"
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* constants: fshader.glsl:3-10, render.hpp:4-11                                               */
#define RENDER_DIST VXO_RENDER_DIST
#define LOCAL_LIGHT_DIST 64
#define AMBIENT 0.4f
#define DIFFUSE 0.8f
#define MAX_OVERBRIGHT 1.25f
#define DEPTH_FIELD_RADIUS 7

int vxo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

uint64_t vxo_fnv1a64(const void* data, size_t nbytes) {
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < nbytes; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

/* ------------------------------------------------------------------------------------------ */
/* host-side voxel accessors                                                                   */

/* render.cpp:189-196 */
int vxo_host_index(vxo_dims g, int x, int y, int z) {
    int index = -1;
    if (x >= 0 && y >= 0 && z >= 0 && x < g.w && y < g.h && z < g.d)
        index = x + (g.w * y) + (g.w * g.h * z);
    return index;
}

/* render.cpp:256-262 */
void vxo_place_voxel(int32_t* vox, vxo_dims g, int x, int y, int z, int32_t v) {
    int index = vxo_host_index(g, x, y, z);
    if (index >= 0) vox[index] = v;
}

/* render.cpp:265-271.  The reference tests x,y,z >= 0 and index < N; with index == -1 (coordinate
 * beyond the high faces) it would write voxels[-1]; no caller in the reference reaches that case
 * (level.cpp guards the high faces), the oracle makes it a no-op. */
void vxo_destroy_voxel(int32_t* vox, vxo_dims g, int x, int y, int z) {
    int index = vxo_host_index(g, x, y, z);
    if (x >= 0 && y >= 0 && z >= 0 && index < g.w * g.h * g.d && index >= 0) vox[index] = -1;
}

/* ------------------------------------------------------------------------------------------ */
/* default level: level.cpp                                                                    */

/* level.cpp:4-27 */
static void place_bush(int32_t* vox, vxo_dims g, int px, int py, int pz, int cr, int cg, int cb, int radius) {
    for (int z = -radius; z < radius; z++)
        for (int y = -radius; y < radius; y++)
            for (int x = -radius; x < radius; x++)
                if (x + px < g.w && y + py < g.h && z + pz < g.d && x + px >= 0 && y + py >= 0 && z + pz >= 0)
                    if (x * x + y * y + z * z < radius * radius) {
                        vxo_destroy_voxel(vox, g, x, y, z);        /* level.cpp:11 -- RELATIVE coords (reference quirk) */
                        int voxel = 0;
                        voxel += cr;
                        voxel = voxel << 8;
                        voxel += cg - ((x + y + z) % 3) * 20;      /* level.cpp:17, C remainder keeps the sign */
                        voxel = voxel << 8;
                        voxel += cb;
                        vxo_place_voxel(vox, g, x + px, y + py, z + pz, voxel);
                    }
}

/* level.cpp:59-79.  "-length>>2" parses as (-length)>>2 (arithmetic shift) */
static void place_trunk(int32_t* vox, vxo_dims g, int px, int py, int pz, int cr, int cg, int cb, int length) {
    for (int z = -length >> 2; z < length >> 2; z++)
        for (int y = 0; y < length; y++)
            for (int x = -length >> 2; x < length >> 2; x++)
                if (x + px < g.w && y + py < g.h && z + pz < g.d) {
                    vxo_destroy_voxel(vox, g, x, y, z);            /* level.cpp:64 -- RELATIVE coords */
                    int voxel = 0;
                    voxel += cr - ((x + z) % 2) * 10;
                    voxel = voxel << 8;
                    voxel += cg - ((x + z) % 2) * 10;
                    voxel = voxel << 8;
                    voxel += cb;
                    vxo_place_voxel(vox, g, x + px, y + py, z + pz, voxel);
                }
}

/* render.cpp:349-352 (fill -1) + level.cpp:82-138 (initVoxels) */
void vxo_init_default_level(int32_t* vox) {
    vxo_dims g = {512, 96, 512};
    for (int i = 0; i < g.w * g.h * g.d; i++) vox[i] = -1;
    for (int z = 0; z < g.d; z++)
        for (int y = 0; y < g.h; y++)
            for (int x = 0; x < g.w; x++) {
                int cv = 5 * ((x + y + z) % 3);
                int voxel;
                vxo_destroy_voxel(vox, g, x, y, z);
                if (y <= 25) {                                       /* stone, level.cpp:94-103 */
                    voxel = 0; voxel += 90 + cv; voxel = voxel << 8; voxel += 90 + cv; voxel = voxel << 8; voxel += 90 + cv;
                    vxo_place_voxel(vox, g, x, y, z, voxel);
                } else if (y <= 33) {                                /* dirt, level.cpp:105-114 */
                    voxel = 0; voxel += 120 + cv; voxel = voxel << 8; voxel += 100 + cv; voxel = voxel << 8; voxel += 0;
                    vxo_place_voxel(vox, g, x, y, z, voxel);
                } else if (y <= 36) {                                /* grass, level.cpp:116-125 */
                    voxel = 0; voxel += 10; voxel = voxel << 8; voxel += 130 + cv; voxel = voxel << 8; voxel += 10;
                    vxo_place_voxel(vox, g, x, y, z, voxel);
                }
            }
    for (int z = 10; z < g.d - 10; z++)                              /* level.cpp:130-137 */
        for (int x = 10; x < g.w - 10; x++)
            if (x % 30 == 0 && z % 25 == 0) {
                place_trunk(vox, g, x + 1 + z % 7, 36, z, 128, 100, 15, 6);
                place_bush(vox, g, x + z % 7, 36 + 10, z, 15, 128, 15, 6);
            }
}

/* ------------------------------------------------------------------------------------------ */
/* depth field: render.cpp:66-100, 226-253, 273-286                                            */

/* render.cpp:66-100.  dist is computed in double (std::sqrt(int)) and narrowed to float. */
int vxo_depth_offsets(float* dist_out, int32_t* xyz_out, int max_entries) {
    int entries = 0;
    for (int zc = -DEPTH_FIELD_RADIUS; zc <= DEPTH_FIELD_RADIUS; zc++)
        for (int yc = -DEPTH_FIELD_RADIUS; yc <= DEPTH_FIELD_RADIUS; yc++)
            for (int xc = -DEPTH_FIELD_RADIUS; xc <= DEPTH_FIELD_RADIUS; xc++)
                if (xc * xc + yc * yc + zc * zc <= DEPTH_FIELD_RADIUS * DEPTH_FIELD_RADIUS) {
                    int xd = xc - (xc > 0) + (xc < 0);
                    int yd = yc - (yc > 0) + (yc < 0);
                    int zd = zc - (zc > 0) + (zc < 0);
                    float dist = (float)(-sqrt((double)(xd * xd + yd * yd + zd * zd)));
                    if (dist <= DEPTH_FIELD_RADIUS) {
                        if (entries < max_entries) {
                            if (dist_out) dist_out[entries] = dist;
                            if (xyz_out) { xyz_out[3 * entries] = xc; xyz_out[3 * entries + 1] = yc; xyz_out[3 * entries + 2] = zc; }
                        }
                        entries++;
                    }
                }
    return entries;
}

#define VXO_MAX_OFFSETS 2744   /* 7*7*7*2*2*2, the reference's scratch bound render.cpp:69 */
static float g_off_dist[VXO_MAX_OFFSETS];
static int32_t g_off_xyz[3 * VXO_MAX_OFFSETS];
static int g_off_count = 0;
static void ensure_offsets(void) {
    /* always called from serial code before any parallel region */
    if (g_off_count == 0) g_off_count = vxo_depth_offsets(g_off_dist, g_off_xyz, VXO_MAX_OFFSETS);
}

/* render.cpp:226-253.  Rule made explicit: the reference reads voxels[-1] for out-of-grid
 * neighbours (render.cpp:237-240); in the reference's Linux/g++ build that word is zero padding,
 * i.e. OUT-OF-GRID NEIGHBOURS COUNT AS SOLID (pinned by the fingerprints). */
static void fix_depth_field_n(int32_t* vox, vxo_dims g, int x, int y, int z, int count) {
    int index = vxo_host_index(g, x, y, z);
    float dist = -DEPTH_FIELD_RADIUS + 1;
    float nearest = dist;
    if (index >= 0 && vox[index] < 0) {
        for (int i = 0; i < count; i++) {
            dist = g_off_dist[i];
            int ic = vxo_host_index(g, x + g_off_xyz[3 * i], y + g_off_xyz[3 * i + 1], z + g_off_xyz[3 * i + 2]);
            int32_t v = (ic >= 0) ? vox[ic] : 0;                    /* OOB = solid */
            if (v >= 0 && dist > nearest) {
                if (dist <= -2.0f) nearest = dist;
                else nearest = 0;
            }
        }
        if (nearest < 0) { int32_t bits; memcpy(&bits, &nearest, 4); vox[index] = bits; }
    }
}

void vxo_fix_depth_field(int32_t* vox, vxo_dims g, int x, int y, int z) {
    ensure_offsets();
    fix_depth_field_n(vox, g, x, y, z, g_off_count);
}

/* render.cpp:273-286 + :357-363.  The reference splits z into 4 pthread slabs that race benignly
 * (writers only turn negative values into other negative values; readers only test the sign), so
 * the result is independent of the schedule; here z-slices are spread over OpenMP threads. */
void vxo_compute_depth_field(int32_t* vox, vxo_dims g, int nthreads) {
    ensure_offsets();
    int count = g_off_count;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int z = 0; z < g.d; z++)
        for (int y = 0; y < g.h; y++)
            for (int x = 0; x < g.w; x++)
                fix_depth_field_n(vox, g, x, y, z, count);
}

/* ------------------------------------------------------------------------------------------ */
/* edits                                                                                       */

/* level.cpp:30-56 */
void vxo_remove_sphere(int32_t* vox, vxo_dims g, int px, int py, int pz, int radius) {
    ensure_offsets();
    for (int z = -radius; z < radius; z++)
        for (int y = -radius; y < radius; y++)
            for (int x = -radius; x < radius; x++)
                if (x + px < g.w && y + py < g.h && z + pz < g.d && x + px >= 0 && y + py >= 0 && z + pz >= 0)
                    if (x * x + y * y + z * z < radius * radius)
                        vxo_destroy_voxel(vox, g, x + px, y + py, z + pz);
    radius += DEPTH_FIELD_RADIUS >> 1;                              /* level.cpp:43 */
    for (int z = -radius; z < radius; z++)
        for (int y = -radius; y < radius; y++)
            for (int x = -radius; x < radius; x++)
                if (x + px < g.w && y + py < g.h && z + pz < g.d && x + px >= 0 && y + py >= 0 && z + pz >= 0)
                    if (x * x + y * y + z * z < radius * radius)
                        fix_depth_field_n(vox, g, x + px, y + py, z + pz, g_off_count);
}

static int vxo_f2i(float x);   /* below */

/* controls.cpp:100-110 (the body under keys[RMB]); glm::ivec3(vec3) truncates */
void vxo_do_destroy(int32_t* vox, vxo_dims g, const float cam_pos[3], const float cam_dir[3], float centre_out[3]) {
    float destroyRange = 15.0f;
    float c[3];
    for (int k = 0; k < 3; k++) c[k] = cam_pos[k] + destroyRange * cam_dir[k];
    vxo_remove_sphere(vox, g, vxo_f2i(c[0]), vxo_f2i(c[1]), vxo_f2i(c[2]), (int)(destroyRange / 2));
    if (centre_out) for (int k = 0; k < 3; k++) centre_out[k] = c[k];
}

/* render.cpp:204-223 */
int vxo_partial_ranges(vxo_dims g, const float start_in[3], const float end_in[3],
                       int64_t* first, int32_t* count, int max_calls) {
    float start[3] = {start_in[0], start_in[1], start_in[2]};
    float end[3] = {end_in[0], end_in[1], end_in[2]};
    int startInd = vxo_host_index(g, (int)start[0], (int)start[1], (int)start[2]);
    int endInd = vxo_host_index(g, (int)end[0], (int)end[1], (int)end[2]);
    if (startInd > endInd) {
        for (int k = 0; k < 3; k++) { float t = start[k]; start[k] = end[k]; end[k] = t; }
    }
    int xLength = (int)(end[0] - start[0]) + 1;
    int calls = 0;
    for (float i = start[2]; i < end[2]; i++)
        for (float j = start[1]; j < end[1]; j++) {
            int offset = vxo_host_index(g, (int)start[0], (int)j, (int)i);
            if (offset != -1) {
                if (calls < max_calls) { if (first) first[calls] = offset; if (count) count[calls] = xLength; }
                calls++;
            }
        }
    return calls;
}

/* ------------------------------------------------------------------------------------------ */
/* shader: fshader.glsl                                                                        */

/* fshader.glsl:33-52, generalised from (512,96,512) to (w,h,d).  The shader multiplies first and
 * range-checks the products; int arithmetic wraps. */
int32_t vxo_shader_index(vxo_dims g, int32_t x, int32_t y, int32_t z) {
    int32_t hit = -1;
    int32_t n = g.w * g.h * g.d;
    int32_t yy = (int32_t)((uint32_t)y * (uint32_t)g.w);
    int32_t zz = (int32_t)((uint32_t)z * (uint32_t)(g.w * g.h));
    int32_t index = (int32_t)((uint32_t)x + (uint32_t)yy + (uint32_t)zz);
    if (index < n && zz >= 0 && zz < n && yy >= 0 && yy < g.w * g.h && x >= 0 && x < g.w) hit = index;
    return hit;
}

/* fshader.glsl:59-129: the statement-for-statement loop lives in vxo_castray_body.inc, shared with the analysis translation
 * units (oracle/vxo_analysis.c, oracle/vxo_trav.c), which define the VXO_HOOK_* macros; here they are empty, so this is
 * the plain restatement and nothing else. */
#define VXO_CAST_RAY_NAME cast_ray
#define VXO_HOOK_ARGS
#define VXO_HOOK_START()
#define VXO_HOOK_ITER()
#define VXO_HOOK_JUMP()
#define VXO_HOOK_END(outcome)
#include "vxo_castray_body.inc"

int32_t vxo_cast_ray(const int32_t* vox, vxo_dims g, const float start[3], const float dir[3],
                     int32_t dist, vxo_ray_out* out) {
    vxo_shader_state st; memset(&st, 0, sizeof st);
    int32_t r = cast_ray(vox, g, &st, start[0], start[1], start[2], dir[0], dir[1], dir[2], dist);
    if (out) {
        for (int k = 0; k < 3; k++) { out->hit_pos[k] = st.hitPos[k]; out->hit_normal[k] = st.hitNormal[k]; }
        out->steps = st.stepCount; out->hit_set = (r >= 0);
    }
    return r;
}

/* GLM: dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z ; normalize(v) = v * (1/sqrt(dot(v,v))) ;
 * length(v) = sqrt(dot(v,v))   (glm/detail/func_geometric.inl:48-55, :82-90, :8-14) */
static float dot3(const float a[3], const float b[3]) {
    float t0 = a[0] * b[0], t1 = a[1] * b[1], t2 = a[2] * b[2];
    return (t0 + t1) + t2;
}
static void normalize3(const float v[3], float out[3]) {
    float inv = 1.0f / sqrtf(dot3(v, v));
    out[0] = v[0] * inv; out[1] = v[1] * inv; out[2] = v[2] * inv;
}

/* fshader.glsl:131-190.  `castRay` is the castRay to use: cast_ray above for the oracle proper; the analysis translation units
 * pass their instrumented / traversal-grid variants (ray kind: 0 primary, 1 global light, 2 local light; dark: the surface
 * faces away from the light -- both ignored by the plain cast_ray). */
static int32_t cast_ray_plain(void* user, const int32_t* vox, vxo_dims g, vxo_shader_state* st, float sx, float sy, float sz,
                              float rx, float ry, float rz, int32_t dist, int kind, int dark) {
    (void)user; (void)kind; (void)dark;
    return cast_ray(vox, g, st, sx, sy, sz, rx, ry, rz, dist);
}

void vxo_shade_pixel_with(vxo_cast_fn castRay, void* user, const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height,
                          int px, int py, float rgba[4], int32_t* hit_index, float* steps,
                          uint32_t* occl_mask, uint32_t* cast_mask, float hit_pos[3], float hit_normal[3],
                          uint64_t counters[5]) {
    vxo_shader_state st; memset(&st, 0, sizeof st);
    /* vshader.glsl:6-9 + quad render.cpp:36-44: vPos = NDC of the pixel centre */
    float vx = ((float)px + 0.5f) / (float)width * 2.0f - 1.0f;
    float vy = ((float)py + 0.5f) / (float)height * 2.0f - 1.0f;
    float col[4] = {(float)0.6, (float)0.7, (float)0.8, 1.0f};      /* :133 */
    float v[3] = {vx * f->aspect, vy, 1.0f}, rd[3];
    normalize3(v, rd);                                              /* :136 */
    /* :137  mat4 * vec4(rd,0): (m[0]*v0 + m[1]*v1) + (m[2]*v2 + m[3]*v3), glm/detail/type_mat4x4.inl:561-572 */
    const float* m = f->rotate;
    float rot[3];
    for (int r = 0; r < 3; r++) {
        float mul0 = m[0 + r] * rd[0], mul1 = m[4 + r] * rd[1], mul2 = m[8 + r] * rd[2], mul3 = m[12 + r] * 0.0f;
        rot[r] = (mul0 + mul1) + (mul2 + mul3);
    }
    int32_t idx = castRay(user, vox, g, &st, f->cam_pos[0], f->cam_pos[1], f->cam_pos[2], rot[0], rot[1], rot[2], RENDER_DIST, 0, 0);   /* :139 */
    float fhp[3] = {st.hitPos[0], st.hitPos[1], st.hitPos[2]};      /* :140 */
    float fhn[3] = {st.hitNormal[0], st.hitNormal[1], st.hitNormal[2]};   /* :141 */
    float primary_steps = st.stepCount;
    uint32_t occl = 0, cast = 0;
    uint64_t n_global = 0, n_local = 0;
    int hit = 0;
    if (f->view_depth_field == 1) {                                 /* :143-145 */
        float gr = st.stepCount / 100.0f;
        col[0] = gr; col[1] = gr; col[2] = gr; col[3] = 1.0f;
    } else {
        float tl[3] = {f->light_pos[0] - fhp[0], f->light_pos[1] - fhp[1], f->light_pos[2] - fhp[2]}, toLight[3];
        normalize3(tl, toLight);                                    /* :147 */
        float multiplier = AMBIENT;                                 /* :149 */
        if (idx != -1 && vox[idx] >= 0) {                           /* :152 */
            hit = 1;
            cast |= 1u; n_global++;
            if (castRay(user, vox, g, &st, fhp[0] + toLight[0] * 0.001f, fhp[1] + toLight[1] * 0.001f, fhp[2] + toLight[2] * 0.001f,
                     toLight[0], toLight[1], toLight[2], RENDER_DIST, 1, !(0.0f < dot3(fhn, toLight))) == -1) {   /* :154 */
                multiplier = multiplier + DIFFUSE * vxo_fmax0(dot3(fhn, toLight));                /* :155 */
            } else occl |= 1u;
            for (int i = 0; i < VXO_MAX_LOCAL_LIGHTS; i++) {        /* :159 */
                if (multiplier >= MAX_OVERBRIGHT) { multiplier = MAX_OVERBRIGHT; break; }     /* :161-164 */
                else if (f->lights[i][0] >= 0 && f->lights[i][1] >= 0 && f->lights[i][2] >= 0) {   /* :167 */
                    float d[3] = {f->lights[i][0] - fhp[0], f->lights[i][1] - fhp[1], f->lights[i][2] - fhp[2]};
                    float lld = sqrtf(dot3(d, d));                  /* :168 */
                    if (lld <= (float)LOCAL_LIGHT_DIST) {           /* :171 */
                        float tll[3]; normalize3(d, tll);           /* :173 */
                        cast |= 2u << i; n_local++;
                        if (castRay(user, vox, g, &st, fhp[0] + tll[0] * 0.001f, fhp[1] + tll[1] * 0.001f, fhp[2] + tll[2] * 0.001f,
                                 tll[0], tll[1], tll[2], vxo_f2i(lld + 1.0f), 2, !(0.0f < dot3(fhn, tll))) == -1) {   /* :175 */
                            multiplier = multiplier + f->lights[i][3] * vxo_fmax0(dot3(fhn, tll)) *
                                         (((float)LOCAL_LIGHT_DIST - lld) / (float)LOCAL_LIGHT_DIST);   /* :177 */
                        } else occl |= 2u << i;
                    }
                }
            }
            int32_t c = vox[idx];                                   /* :184-187 */
            col[0] = (float)((c & 0x00FF0000) >> 16) / 255.0f * multiplier;
            col[1] = (float)((c & 0x0000FF00) >> 8) / 255.0f * multiplier;
            col[2] = (float)(c & 0x000000FF) / 255.0f * multiplier;
            col[3] = 1.0f;
        }
    }
    if (rgba) for (int k = 0; k < 4; k++) rgba[k] = col[k];
    if (hit_index) *hit_index = idx;
    if (steps) *steps = primary_steps;
    if (occl_mask) *occl_mask = occl;
    if (cast_mask) *cast_mask = cast;
    if (hit_pos) for (int k = 0; k < 3; k++) hit_pos[k] = fhp[k];
    if (hit_normal) for (int k = 0; k < 3; k++) hit_normal[k] = fhn[k];
    if (counters) {
        counters[0] += 1; counters[1] += n_global; counters[2] += n_local; counters[3] += st.fetches; counters[4] += (uint64_t)hit;
    }
}

void vxo_shade_pixel(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height,
                     int px, int py, float rgba[4], int32_t* hit_index, float* steps,
                     uint32_t* occl_mask, uint32_t* cast_mask, float hit_pos[3], float hit_normal[3],
                     uint64_t counters[5]) {
    vxo_shade_pixel_with(cast_ray_plain, NULL, vox, g, f, width, height, px, py, rgba, hit_index, steps, occl_mask, cast_mask,
                         hit_pos, hit_normal, counters);
}

/* default-framebuffer store: RGBA8 UNORM */
static uint8_t unorm8(float c) {
    if (!(c > 0.0f)) c = 0.0f;      /* also NaN -> 0 */
    if (c > 1.0f) c = 1.0f;
    return (uint8_t)(int)(c * 255.0f + 0.5f);
}

void vxo_render(const int32_t* vox, vxo_dims g, const vxo_frame* f, int width, int height,
                int y0, int y1, float* rgba_f32, uint8_t* rgba8, int32_t* hit_index,
                uint16_t* steps, uint32_t* occl_mask, uint32_t* cast_mask,
                uint64_t counters[5], int nthreads) {
    uint64_t tot[5] = {0, 0, 0, 0, 0};
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        uint64_t loc[5] = {0, 0, 0, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int py = y0; py < y1; py++)
            for (int px = 0; px < width; px++) {
                float rgba[4]; int32_t hi; float sc; uint32_t om, cm;
                vxo_shade_pixel(vox, g, f, width, height, px, py, rgba, &hi, &sc, &om, &cm, NULL, NULL, loc);
                size_t p = (size_t)py * (size_t)width + (size_t)px;
                if (rgba_f32) for (int k = 0; k < 4; k++) rgba_f32[4 * p + k] = rgba[k];
                if (rgba8) for (int k = 0; k < 4; k++) rgba8[4 * p + k] = unorm8(rgba[k]);
                if (hit_index) hit_index[p] = hi;
                if (steps) steps[p] = (uint16_t)(sc > 65535.0f ? 65535 : (int)sc);
                if (occl_mask) occl_mask[p] = om;
                if (cast_mask) cast_mask[p] = cm;
            }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int k = 0; k < 5; k++) tot[k] += loc[k];
    }
    if (counters) for (int k = 0; k < 5; k++) counters[k] += tot[k];
}

