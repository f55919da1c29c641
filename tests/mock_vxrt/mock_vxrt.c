/*
 * mock_vxrt.c -- TEST INFRASTRUCTURE (CPU).  A stand-in `libvxrt.so` for tests/test_glshim.py: it implements the part of
 * include/vxrt.h that the link-level GL shim (voxel-rt_b200/csrc/host/vxrt_glshim.cpp) calls, keeps the "device" grid in
 * host memory, renders through the ORACLE when a frame is dumped, and -- because it runs inside the reference's own
 * process -- checks at every draw that what reached it through the shim equals the reference's globals
 * (camPos, lightPos, rotateMatrix, localLights, ... main.cpp:7-44) and that its grid equals the reference's voxels[].
 * Every call is logged as one JSON line to $MOCK_VXRT_LOG.  Never shipped, never loaded by the product: the test puts
 * its directory in LD_LIBRARY_PATH so that the dynamic linker resolves the shim's `libvxrt.so` to it.
 */
#define _GNU_SOURCE
#include "../../include/vxrt.h"
#include "../../oracle/vxo.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct vxrt_ctx {
    vxrt_config cfg;
    int32_t* grid;
    size_t nvox;
    vxrt_frame frame;
    long draws, full_uploads, logged_draws;
    long ready_uploads;                 /* log / check draws from this many whole-grid uploads on */
};

static FILE* g_log;
static char g_err[256];

static FILE* logf_(void) {
    if (!g_log) {
        const char* p = getenv("MOCK_VXRT_LOG");
        g_log = p ? fopen(p, "w") : stderr;
        if (!g_log) g_log = stderr;
    }
    return g_log;
}

static const void* host_global(const char* name) { return dlsym(RTLD_DEFAULT, name); }

const char* vxrt_last_error(void) { return g_err; }
int vxrt_device_available(void) { return 0; }

int vxrt_create(const vxrt_config* cfg, vxrt_ctx** out) {
    vxrt_ctx* c = (vxrt_ctx*)calloc(1, sizeof *c);
    c->cfg = *cfg;
    c->nvox = (size_t)cfg->grid_w * cfg->grid_h * cfg->grid_d;
    c->grid = (int32_t*)malloc(c->nvox * 4);
    memset(c->grid, 0xFF, c->nvox * 4);
    const char* r = getenv("VXRT_GLSHIM_READY_UPLOADS");
    c->ready_uploads = r ? atol(r) : 0;
    fprintf(logf_(), "{\"call\": \"create\", \"grid\": [%d, %d, %d], \"width\": %d, \"height\": %d, \"world\": %d}\n",
            cfg->grid_w, cfg->grid_h, cfg->grid_d, cfg->width, cfg->height, cfg->world);
    *out = c;
    return VXRT_OK;
}

void vxrt_destroy(vxrt_ctx* c) {
    if (!c) return;
    fprintf(logf_(), "{\"call\": \"destroy\", \"draws\": %ld, \"full_uploads\": %ld, \"grid_fnv\": \"%016llx\"}\n", c->draws,
            c->full_uploads, (unsigned long long)vxo_fnv1a64(c->grid, c->nvox * 4));
    fflush(logf_());
    free(c->grid);
    free(c);
}

int vxrt_upload_grid(vxrt_ctx* c, const int32_t* voxels, size_t count) {
    if (count != c->nvox) { snprintf(g_err, sizeof g_err, "upload_grid: count"); return VXRT_ERR_INVALID; }
    memcpy(c->grid, voxels, count * 4);
    c->full_uploads++;
    fprintf(logf_(), "{\"call\": \"upload_grid\", \"count\": %zu, \"nth\": %ld}\n", count, c->full_uploads);
    return VXRT_OK;
}

int vxrt_upload_rows(vxrt_ctx* c, size_t rows, size_t row_len, const int64_t* firsts, const int32_t* packed) {
    for (size_t r = 0; r < rows; r++)
        if (firsts[r] < 0 || (size_t)firsts[r] + row_len > c->nvox) { snprintf(g_err, sizeof g_err, "upload_rows: range"); return VXRT_ERR_INVALID; }
    for (size_t r = 0; r < rows; r++) memcpy(c->grid + firsts[r], packed + r * row_len, row_len * 4);
    /* the camera the reference's doDestroy (controls.cpp:100-110) just used, so that the test can replay the edit */
    const float zero[3] = {0, 0, 0};
    const float* cp = (const float*)host_global("camPos");
    const float* cd = (const float*)host_global("camDir");
    if (!cp) cp = zero;
    if (!cd) cd = zero;
    fprintf(logf_(), "{\"call\": \"upload_rows\", \"rows\": %zu, \"row_len\": %zu, \"first\": %lld, \"grid_fnv\": \"%016llx\", "
                     "\"cam_pos\": [%.9g, %.9g, %.9g], \"cam_dir\": [%.9g, %.9g, %.9g]}\n", rows, row_len,
            rows ? (long long)firsts[0] : -1LL, (unsigned long long)vxo_fnv1a64(c->grid, c->nvox * 4),
            cp[0], cp[1], cp[2], cd[0], cd[1], cd[2]);
    return VXRT_OK;
}

int vxrt_set_frame(vxrt_ctx* c, const vxrt_frame* f) { c->frame = *f; return VXRT_OK; }

int vxrt_resize(vxrt_ctx* c, int w, int h) {
    c->cfg.width = w; c->cfg.height = h;
    c->frame.aspect = (float)w / (float)h;
    fprintf(logf_(), "{\"call\": \"resize\", \"width\": %d, \"height\": %d}\n", w, h);
    return VXRT_OK;
}

static int differs(const void* a, const char* global, size_t n) {
    const void* b = host_global(global);
    return b ? memcmp(a, b, n) != 0 : 1;
}

int vxrt_render(vxrt_ctx* c) {
    c->draws++;
    if (c->full_uploads < c->ready_uploads) return VXRT_OK;
    const vxrt_frame* f = &c->frame;
    /* what the shim handed over == the reference's own globals at draw time (main.cpp:59 draws right after
       updateUniforms render.cpp:289-296; only window events lie in between) */
    int bad = 0;
    bad += differs(f->cam_pos, "camPos", 12);
    bad += differs(f->cam_rotation, "camRotation", 8);
    bad += differs(f->light_pos, "lightPos", 12);
    bad += differs(&f->aspect, "aspectRatio", 4);
    bad += differs(f->rotate, "rotateMatrix", 64);
    bad += differs(&f->view_depth_field, "viewDepthField", 4);
    bad += differs(f->lights, "localLights", 256);
    const int* sw = (const int*)host_global("screenWidth");
    const int* sh = (const int*)host_global("screenHeight");
    const int size_ok = sw && sh && *sw == c->cfg.width && *sh == c->cfg.height;
    long long grid_diff = -1;
    const int32_t* hv = (const int32_t*)host_global("voxels");
    if (hv) {
        grid_diff = 0;
        if (memcmp(hv, c->grid, c->nvox * 4) != 0)
            for (size_t i = 0; i < c->nvox; i++) grid_diff += hv[i] != c->grid[i];
    }
    int active = 0;
    for (int i = 0; i < VXRT_MAX_LOCAL_LIGHTS; i++) active += !(f->lights[i][0] < 0 || f->lights[i][1] < 0 || f->lights[i][2] < 0);
    fprintf(logf_(), "{\"call\": \"render\", \"frame\": %ld, \"width\": %d, \"height\": %d, \"uniform_mismatch\": %d, \"size_ok\": %d, "
                     "\"grid_diff\": %lld, \"cam_pos\": [%.9g, %.9g, %.9g], \"light_pos\": [%.9g, %.9g, %.9g], \"aspect\": %.9g, "
                     "\"view\": %d, \"lights_active\": %d, \"rotate0\": %.9g}\n",
            c->logged_draws, c->cfg.width, c->cfg.height, bad, size_ok, grid_diff, f->cam_pos[0], f->cam_pos[1], f->cam_pos[2],
            f->light_pos[0], f->light_pos[1], f->light_pos[2], f->aspect, f->view_depth_field, active, f->rotate[0]);
    c->logged_draws++;
    return VXRT_OK;
}

int vxrt_sync(vxrt_ctx* c) { (void)c; return VXRT_OK; }

/* the oracle renders the frame the shim asked for */
int vxrt_write_ppm(vxrt_ctx* c, const char* path) {
    const int w = c->cfg.width, h = c->cfg.height;
    uint8_t* rgba = (uint8_t*)malloc((size_t)w * h * 4);
    vxo_dims g = {c->cfg.grid_w, c->cfg.grid_h, c->cfg.grid_d};
    vxo_frame vf;
    memcpy(&vf, &c->frame, sizeof vf);
    vxo_render(c->grid, g, &vf, w, h, 0, h, NULL, rgba, NULL, NULL, NULL, NULL, NULL, 0);
    FILE* fp = fopen(path, "wb");
    if (!fp) { free(rgba); snprintf(g_err, sizeof g_err, "write_ppm: cannot open %s", path); return VXRT_ERR_IO; }
    fprintf(fp, "P6\n%d %d\n255\n", w, h);
    for (int y = h - 1; y >= 0; y--)
        for (int x = 0; x < w; x++) fwrite(rgba + ((size_t)y * w + x) * 4, 1, 3, fp);
    fclose(fp);
    free(rgba);
    fprintf(logf_(), "{\"call\": \"write_ppm\", \"path\": \"%s\", \"after_draw\": %ld, \"grid_fnv\": \"%016llx\"}\n", path, c->logged_draws - 1,
            (unsigned long long)vxo_fnv1a64(c->grid, c->nvox * 4));
    return VXRT_OK;
}

/* VXRTGRD1 (include/vxrt.h): 64-byte header + the voxels */
int vxrt_save_grid(vxrt_ctx* c, const char* path) {
    FILE* fp = fopen(path, "wb");
    if (!fp) { snprintf(g_err, sizeof g_err, "save_grid: cannot open %s", path); return VXRT_ERR_IO; }
    unsigned char hd[64];
    memset(hd, 0, sizeof hd);
    memcpy(hd, "VXRTGRD1", 8);
    const uint32_t ext[4] = {(uint32_t)c->cfg.grid_w, (uint32_t)c->cfg.grid_h, (uint32_t)c->cfg.grid_d, 0};
    const uint64_t tail[2] = {c->nvox, vxo_fnv1a64(c->grid, c->nvox * 4)};
    memcpy(hd + 8, ext, 16);
    memcpy(hd + 24, tail, 16);
    fwrite(hd, 1, sizeof hd, fp);
    fwrite(c->grid, 4, c->nvox, fp);
    fclose(fp);
    fprintf(logf_(), "{\"call\": \"save_grid\", \"path\": \"%s\"}\n", path);
    return VXRT_OK;
}
