// vxrt_render.cpp -- see vxrt_render.hpp.  Host logic mirrors src/render.cpp / src/controls.cpp of the reference;
// everything data-parallel goes through the C ABI to the sm_100a kernels.
#include "vxrt_render.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vxrt_host {

void Render::check(int rc) {
    if (rc < 0) {
        err_ = vxrt_last_error();
        // the reference aborts on GL set-up failure (render.cpp:140-143,152-162) and otherwise ignores errors
        fprintf(stderr, "vxrt: %s\n", err_.c_str());
        if (rc == VXRT_ERR_NO_DEVICE || rc == VXRT_ERR_CUDA) exit(EXIT_FAILURE);
    }
}

Render::~Render() { vxrt_destroy(ctx_); }

int Render::getVoxelIndex(int x, int y, int z) const {
    int index = -1;
    if (x >= 0 && y >= 0 && z >= 0 && x < VOXELS_WIDTH && y < VOXELS_HEIGHT && z < VOXELS_DEPTH)
        index = x + (VOXELS_WIDTH * y) + (VOXELS_WIDTH * VOXELS_HEIGHT * z);
    return index;
}

void Render::initRender() {
    vxrt_config cfg{};
    cfg.grid_w = VOXELS_WIDTH; cfg.grid_h = VOXELS_HEIGHT; cfg.grid_d = VOXELS_DEPTH;
    cfg.width = screenWidth; cfg.height = screenHeight;
    cfg.device = device; cfg.rank = 0; cfg.world = 1; cfg.flags = 0;
    check(vxrt_create(&cfg, &ctx_));
    if (!ctx_) exit(EXIT_FAILURE);
    aspectRatio = (float)screenWidth / screenHeight;
    check(vxrt_init_local_lights(ctx_));                     // initLocalLights render.cpp:343
    for (auto& l : localLights) { l[0] = -1.0f; l[1] = -1.0f; l[2] = -1.0f; l[3] = 0.0f; }
    updateUniforms();                                        // render.cpp:346
    // render.cpp:349-363: fill -1, initVoxels(), depth field.  The reference sweeps the depth field on 4 CPU
    // threads for ~16 s and renders without it meanwhile; here level and depth field are built on the device
    // before the first frame and the host mirror is downloaded once.
    check(vxrt_generate_default_level(ctx_));
    check(vxrt_build_depth_field(ctx_));
    voxels.resize((size_t)VOXELS_WIDTH * VOXELS_HEIGHT * VOXELS_DEPTH);
    check(vxrt_download_grid(ctx_, voxels.data(), voxels.size()));
}

void Render::updateGeometry() { check(vxrt_upload_grid(ctx_, voxels.data(), voxels.size())); }

void Render::updatePartialGeometry(vec3 start, vec3 end) {
    const float s[3] = {start.x, start.y, start.z}, e[3] = {end.x, end.y, end.z};
    check(vxrt_update_partial(ctx_, s, e, voxels.data(), nullptr));
}

void Render::placeVoxel(int x, int y, int z, int voxel) {
    const int index = getVoxelIndex(x, y, z);
    if (index >= 0) { voxels[index] = voxel; check(vxrt_place_voxel(ctx_, x, y, z, voxel)); }
}

void Render::destroyVoxel(int x, int y, int z) {
    const int index = getVoxelIndex(x, y, z);
    if (x >= 0 && y >= 0 && z >= 0 && index >= 0) { voxels[index] = -1; check(vxrt_destroy_voxel(ctx_, x, y, z)); }
}

// host-mirror version of render.cpp:226-253 (out-of-grid neighbours count as solid); the caller uploads the
// touched cells with updatePartialGeometry like the reference does.  Bulk edits should use removeSphere().
void Render::fixDepthField(int x, int y, int z) {
    const int index = getVoxelIndex(x, y, z);
    float nearest = -DEPTH_FIELD_RADIUS + 1;
    if (index >= 0 && voxels[index] < 0) {
        const int R = DEPTH_FIELD_RADIUS;
        for (int zc = -R; zc <= R; zc++)
            for (int yc = -R; yc <= R; yc++)
                for (int xc = -R; xc <= R; xc++) {
                    if (xc * xc + yc * yc + zc * zc > R * R) continue;
                    const int xd = xc - (xc > 0) + (xc < 0), yd = yc - (yc > 0) + (yc < 0), zd = zc - (zc > 0) + (zc < 0);
                    const float dist = (float)(-std::sqrt((double)(xd * xd + yd * yd + zd * zd)));
                    const int ic = getVoxelIndex(x + xc, y + yc, z + zc);
                    const int v = ic >= 0 ? voxels[ic] : 0;
                    if (v >= 0 && dist > nearest) nearest = (dist <= -2.0f) ? dist : 0.0f;
                }
        if (nearest < 0) { int32_t bits; memcpy(&bits, &nearest, 4); voxels[index] = bits; }
    }
}

void Render::removeSphere(ivec3 pos, int radius) {
    check(vxrt_edit_remove_sphere(ctx_, pos.x, pos.y, pos.z, radius));
    const int r = radius + (DEPTH_FIELD_RADIUS >> 1);
    const int32_t lo[3] = {pos.x - r, pos.y - r, pos.z - r}, hi[3] = {pos.x + r, pos.y + r, pos.z + r};
    check(vxrt_download_box(ctx_, lo, hi, voxels.data()));   // keeps collision / gravity (controls.cpp:10-19) coherent
}

void Render::doDestroy() {
    const float destroyRange = 15.0f;                        // controls.cpp:101
    const vec3 c{camPos.x + destroyRange * camDir.x, camPos.y + destroyRange * camDir.y, camPos.z + destroyRange * camDir.z};
    removeSphere(ivec3{(int)c.x, (int)c.y, (int)c.z}, (int)(destroyRange / 2));
    // no updatePartialGeometry: the edit already happened on the device grid
}

void Render::placeLocalLight(float x, float y, float z, float diffuse) {
    for (int i = 0; i < MAX_LOCAL_LIGHTS; i++)
        if (localLights[i][0] < 0 || localLights[i][1] < 0 || localLights[i][2] < 0) {
            localLights[i][0] = x; localLights[i][1] = y; localLights[i][2] = z; localLights[i][3] = diffuse;
            break;
        }
}

void Render::lightUpdate() {                                  // render.cpp:388-402 (vxrt_controls.cpp: light_update)
    const float start[3] = {startLightPos.x, startLightPos.y, startLightPos.z};
    float pos[3];
    light_update(fps, lightRotation, start, pos);
    lightPos = vec3{pos[0], pos[1], pos[2]};
}

void Render::setMouseLook(float rx, float ry) {               // controls.cpp:137-142
    camRotation = vec2{rx, ry};
    Mat4 m;
    float dir[3];
    mouse_look_matrix(rx, ry, m, dir);
    memcpy(rotateMatrix, m.m, sizeof rotateMatrix);
    camDir = vec3{dir[0], dir[1], dir[2]};
}

// ---- gameplay: the Player of vxrt_controls.cpp works on this object's "globals" ----
void Render::toPlayer() {
    player_.cam_pos[0] = camPos.x; player_.cam_pos[1] = camPos.y; player_.cam_pos[2] = camPos.z;
    player_.cam_dir[0] = camDir.x; player_.cam_dir[1] = camDir.y; player_.cam_dir[2] = camDir.z;
    player_.cam_rotation[0] = camRotation.x; player_.cam_rotation[1] = camRotation.y;
    memcpy(player_.rotate_matrix.m, rotateMatrix, sizeof rotateMatrix);
    player_.view_depth_field = viewDepthField;
    for (int k = 0; k < KEYS; k++) player_.keys[k] = keys[k];
    player_.fps = fps; player_.mouse_x = mouseX; player_.mouse_y = mouseY; player_.screen_w = screenWidth; player_.screen_h = screenHeight;
    player_.voxels = voxels.data(); player_.w = VOXELS_WIDTH; player_.h = VOXELS_HEIGHT; player_.d = VOXELS_DEPTH;
}

void Render::fromPlayer() {
    camPos = vec3{player_.cam_pos[0], player_.cam_pos[1], player_.cam_pos[2]};
    camDir = vec3{player_.cam_dir[0], player_.cam_dir[1], player_.cam_dir[2]};
    camRotation = vec2{player_.cam_rotation[0], player_.cam_rotation[1]};
    memcpy(rotateMatrix, player_.rotate_matrix.m, sizeof rotateMatrix);
    viewDepthField = player_.view_depth_field;
    for (int k = 0; k < KEYS; k++) keys[k] = player_.keys[k];
    if (player_.light_requested) {                            // controls.cpp:46-49
        player_.light_requested = false;
        placeLocalLight(player_.light_pos[0], player_.light_pos[1], player_.light_pos[2], 0.5f);
    }
}

void Render::movementUpdate() { toPlayer(); player_.movementUpdate(); fromPlayer(); }
void Render::doGravity() { toPlayer(); player_.doGravity(); fromPlayer(); }
void Render::doMouseLook() { toPlayer(); player_.doMouseLook(); fromPlayer(); }

void Render::updateUniforms() {
    vxrt_frame f{};
    f.cam_pos[0] = camPos.x; f.cam_pos[1] = camPos.y; f.cam_pos[2] = camPos.z;
    f.cam_rotation[0] = camRotation.x; f.cam_rotation[1] = camRotation.y;
    f.light_pos[0] = lightPos.x; f.light_pos[1] = lightPos.y; f.light_pos[2] = lightPos.z;
    f.aspect = aspectRatio;
    memcpy(f.rotate, rotateMatrix, sizeof f.rotate);
    f.view_depth_field = viewDepthField;
    memcpy(f.lights, localLights, sizeof f.lights);
    check(vxrt_set_frame(ctx_, &f));
}

void Render::reshape(int width, int height) {
    screenHeight = height; screenWidth = width;
    aspectRatio = (float)width / height;
    check(vxrt_resize(ctx_, width, height));
}

void Render::draw() { check(vxrt_render(ctx_)); }

bool Render::saveGrid(const std::string& path) {
    const int rc = vxrt_save_grid(ctx_, path.c_str());
    if (rc != VXRT_OK) err_ = vxrt_last_error();
    return rc == VXRT_OK;
}

bool Render::loadGrid(const std::string& path) {
    int rc = vxrt_load_grid(ctx_, path.c_str());
    if (rc == VXRT_OK) rc = vxrt_download_grid(ctx_, voxels.data(), voxels.size());    // keep the collision mirror coherent
    if (rc != VXRT_OK) err_ = vxrt_last_error();
    return rc == VXRT_OK;
}

bool Render::writePPM(const std::string& path) { const int rc = vxrt_write_ppm(ctx_, path.c_str()); check(rc); return rc == 0; }

bool Render::readPixels(std::vector<uint8_t>& rgba) {
    rgba.resize((size_t)screenWidth * screenHeight * 4);
    const int rc = vxrt_read_rgba8(ctx_, rgba.data());
    check(rc);
    return rc == 0;
}

vxrt_stats Render::stats() { vxrt_stats s{}; check(vxrt_get_stats(ctx_, &s)); return s; }

}  // namespace vxrt_host
