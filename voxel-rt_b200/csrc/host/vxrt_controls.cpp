// vxrt_controls.cpp -- see vxrt_controls.hpp.  Statement order follows src/controls.cpp; vector helpers follow the
// operation order of the reference's GLM so that results are bit-identical on the same libm.
#include "vxrt_controls.hpp"

#include <cmath>

namespace vxrt_host {

static const float PI = 3.14159f;                      // controls.cpp:4
static const float MAP_EDGE_OFFSET = 1.0f;             // controls.cpp:6

Mat4 mat4_identity() {
    Mat4 r{};
    for (int i = 0; i < 4; i++) r.m[5 * i] = 1.0f;
    return r;
}

// glm: dot = (x*x + y*y) + z*z ; normalize = v * (1 / sqrt(dot))
static void normalize3(const float v[3], float out[3]) {
    const float t0 = v[0] * v[0], t1 = v[1] * v[1], t2 = v[2] * v[2];
    const float inv = 1.0f / std::sqrt((t0 + t1) + t2);
    out[0] = v[0] * inv; out[1] = v[1] * inv; out[2] = v[2] * inv;
}

Mat4 mat4_rotate(const Mat4& m, float angle, float ax, float ay, float az) {
    const float a = angle, c = std::cos(a), s = std::sin(a);
    const float v[3] = {ax, ay, az};
    float axis[3];
    normalize3(v, axis);
    const float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    Mat4 out;
    for (int col = 0; col < 3; col++)
        for (int r = 0; r < 4; r++)     // Result[col] = m[0]*R[col][0] + m[1]*R[col][1] + m[2]*R[col][2]
            out.m[4 * col + r] = (m.m[r] * R[col][0] + m.m[4 + r] * R[col][1]) + m.m[8 + r] * R[col][2];
    for (int r = 0; r < 4; r++) out.m[12 + r] = m.m[12 + r];
    return out;
}

Mat4 mat4_mul(const Mat4& a, const Mat4& b) {
    Mat4 out;
    for (int col = 0; col < 4; col++)
        for (int r = 0; r < 4; r++)     // Result[col] = A0*B[col][0] + A1*B[col][1] + A2*B[col][2] + A3*B[col][3], left to right
            out.m[4 * col + r] = ((a.m[r] * b.m[4 * col] + a.m[4 + r] * b.m[4 * col + 1]) + a.m[8 + r] * b.m[4 * col + 2]) + a.m[12 + r] * b.m[4 * col + 3];
    return out;
}

void mat4_mul_vec4(const Mat4& m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; r++)         // (m[0]*v0 + m[1]*v1) + (m[2]*v2 + m[3]*v3)
        out[r] = (m.m[r] * v[0] + m.m[4 + r] * v[1]) + (m.m[8 + r] * v[2] + m.m[12 + r] * v[3]);
}

void vec4_mul_mat4(const float v[4], const Mat4& m, float out[4]) {
    for (int c = 0; c < 4; c++)         // m[c][0]*v0 + m[c][1]*v1 + m[c][2]*v2 + m[c][3]*v3, left to right
        out[c] = ((m.m[4 * c] * v[0] + m.m[4 * c + 1] * v[1]) + m.m[4 * c + 2] * v[2]) + m.m[4 * c + 3] * v[3];
}

void light_update(long long fps, float& light_rotation, const float start[3], float light_pos[3]) {   // render.cpp:388-402
    float increment = 0.3f / fps;
    if (light_rotation >= 360.0f) light_rotation = 0.0f;
    else if (light_rotation >= 180.0f && light_rotation < 270.0f) increment *= 8.0f;
    light_rotation += increment;
    const Mat4 rot = mat4_rotate(mat4_identity(), light_rotation * 0.01745329251994329576923690768489f, 0, 0, 1);
    const float v[4] = {start[0], start[1], start[2], 1.0f};
    float o[4];
    mat4_mul_vec4(rot, v, o);
    light_pos[0] = o[0]; light_pos[1] = o[1]; light_pos[2] = o[2];
}

void mouse_look_matrix(float rx, float ry, Mat4& rotate, float cam_dir[3]) {     // controls.cpp:137-142
    const Mat4 rotX = mat4_rotate(mat4_identity(), rx, 1, 0, 0);
    const Mat4 rotY = mat4_rotate(mat4_identity(), ry, 0, 1, 0);
    rotate = mat4_mul(rotY, rotX);
    const float f[4] = {0, 0, 1, 1};
    float o[4];
    mat4_mul_vec4(rotate, f, o);
    cam_dir[0] = o[0]; cam_dir[1] = o[1]; cam_dir[2] = o[2];
}

int Player::voxel_at(int x, int y, int z) const {
    if (x >= 0 && y >= 0 && z >= 0 && x < w && y < h && z < d) return voxels[x + w * y + w * h * z];
    return 0;                           // getVoxelIndex() == -1: the reference reads voxels[-1], zero in its build => solid
}

int Player::collided() const {                              // controls.cpp:10-19
    int c = 0;
    for (int i = 1; i < PLAYER_HEIGHT; i++)
        if (voxel_at((int)cam_pos[0], (int)cam_pos[1] - PLAYER_HEIGHT + i, (int)cam_pos[2]) > -1) c = i;
    return c;
}

void Player::movementUpdate() {                             // controls.cpp:22-74
    const float speed = 36.0f / fps;
    const Mat4 rot = mat4_rotate(mat4_identity(), 90.0f * 0.01745329251994329576923690768489f, 0, 1, 0);   // glm::radians(90.0f)
    const float flat[3] = {cam_dir[0], 0.0f, cam_dir[2]};
    float front[3], side[3];
    normalize3(flat, front);
    for (float& f : front) f *= speed;
    const float v4[4] = {cam_dir[0], 0.0f, cam_dir[2], 0.0f};
    float rv[4];
    vec4_mul_mat4(v4, rot, rv);
    normalize3(rv, side);
    for (float& f : side) f *= speed;
    float step[3] = {0, 0, 0};
    if (keys[KEY_W]) for (int k = 0; k < 3; k++) { cam_pos[k] += front[k]; step[k] += front[k]; }
    if (keys[KEY_S]) for (int k = 0; k < 3; k++) { cam_pos[k] -= front[k]; step[k] += -front[k]; }
    if (keys[KEY_A]) for (int k = 0; k < 3; k++) { cam_pos[k] += side[k]; step[k] += side[k]; }
    if (keys[KEY_D]) for (int k = 0; k < 3; k++) { cam_pos[k] -= side[k]; step[k] += -side[k]; }
    if (keys[KEY_T]) {                                      // controls.cpp:46-49
        light_requested = true;
        light_pos[0] = cam_pos[0]; light_pos[1] = cam_pos[1] - 1.5f; light_pos[2] = cam_pos[2];
        keys[KEY_T] = false;
    }
    if (keys[SPACE]) {                                      // controls.cpp:50-55
        if (voxel_at((int)cam_pos[0], (int)cam_pos[1] - PLAYER_HEIGHT, (int)cam_pos[2]) > -1 && gravity == 0.0f) gravity += 35.0f;
    }
    if (keys[SHIFT]) { keys[SHIFT] = false; view_depth_field = !view_depth_field; }
    int collision = collided();                             // controls.cpp:61-73
    while (collision) {
        if (step[0] * step[0] + step[1] * step[1] + step[2] * step[2] > 0) {
            collision = collided();
            if (collision > PLAYER_HEIGHT >> 2) for (int k = 0; k < 3; k++) cam_pos[k] -= step[k];
            else cam_pos[1] += collision;
        } else break;       // the reference spins forever here (stuck inside a solid without having moved); we leave the loop
    }
}

static float clampf(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }   // glm::min(glm::max(v, lo), hi)

void Player::doGravity() {                                  // controls.cpp:77-98
    cam_pos[1] += gravity / fps;
    cam_pos[0] = clampf(cam_pos[0], MAP_EDGE_OFFSET, (float)w - MAP_EDGE_OFFSET - 1);
    cam_pos[1] = clampf(cam_pos[1], MAP_EDGE_OFFSET + PLAYER_HEIGHT, (float)h - MAP_EDGE_OFFSET - 1);
    cam_pos[2] = clampf(cam_pos[2], MAP_EDGE_OFFSET, (float)d - MAP_EDGE_OFFSET - 1);
    const int below = voxel_at((int)cam_pos[0], (int)cam_pos[1] - PLAYER_HEIGHT, (int)cam_pos[2]);
    const int collision = collided();
    if (collision) { cam_pos[1] -= gravity / fps; gravity = 0.0f; }
    else if (below < 0) gravity -= 70.0f / fps;
}

void Player::doMouseLook() {                                // controls.cpp:112-144
    const float sensitivity = 7.0f;
    const float lookX = (float)mouse_x * 2.0f / screen_w - 1.0f;
    const float lookY = (float)mouse_y * 2.0f / screen_h - 1.0f;
    if (keys[LMB]) {
        cam_rotation[0] += lookY * sensitivity / (float)fps;
        cam_rotation[1] += lookX * sensitivity / (float)fps;
        if (cam_rotation[0] >= 2 * PI) cam_rotation[0] = 0.0f; else if (cam_rotation[0] <= -2 * PI) cam_rotation[0] = 0.0f;
        if (cam_rotation[1] >= 2 * PI) cam_rotation[1] = 0.0f; else if (cam_rotation[1] <= -2 * PI) cam_rotation[1] = 0.0f;
        cam_rotation[0] = std::fmin(PI / 2.0f, std::fmax(-PI / 2.0f, cam_rotation[0]));
        mouse_look_matrix(cam_rotation[0], cam_rotation[1], rotate_matrix, cam_dir);
    }
}

}  // namespace vxrt_host

// ---- C hooks for the parity tests (tests/test_host_logic.py) ------------------------------------------------
extern "C" {
using namespace vxrt_host;
Player* vxh_player_create(const int32_t* voxels, int w, int h, int d) { Player* p = new Player(); p->voxels = voxels; p->w = w; p->h = h; p->d = d; return p; }
void vxh_player_destroy(Player* p) { delete p; }
void vxh_player_set(Player* p, const float cam[3], const float dir[3], const float camrot[2], long long fps) {
    for (int k = 0; k < 3; k++) { p->cam_pos[k] = cam[k]; p->cam_dir[k] = dir[k]; }
    p->cam_rotation[0] = camrot[0]; p->cam_rotation[1] = camrot[1];
    p->rotate_matrix = mat4_identity();
    p->gravity = 0.0f; p->view_depth_field = 0; p->light_requested = false;
    for (int k = 0; k < KEYS; k++) p->keys[k] = false;
    p->fps = fps;
}
void vxh_player_keys(Player* p, const unsigned char keys[9]) { for (int k = 0; k < KEYS; k++) p->keys[k] = keys[k] != 0; }
void vxh_player_mouse(Player* p, int mx, int my, int sw, int sh) { p->mouse_x = mx; p->mouse_y = my; p->screen_w = sw; p->screen_h = sh; }
// one iteration of the game loop's host part (main.cpp:62-65): movementUpdate, doMouseLook, doGravity
void vxh_player_step(Player* p) { p->movementUpdate(); p->doMouseLook(); p->doGravity(); }
// out: camPos[3] camDir[3] camRotation[2] rotate[16] gravity-free state the reference exposes (24 floats) + view flag via return
int vxh_player_get(Player* p, float out24[24]) {
    for (int k = 0; k < 3; k++) { out24[k] = p->cam_pos[k]; out24[3 + k] = p->cam_dir[k]; }
    out24[6] = p->cam_rotation[0]; out24[7] = p->cam_rotation[1];
    for (int k = 0; k < 16; k++) out24[8 + k] = p->rotate_matrix.m[k];
    return p->view_depth_field;
}
int vxh_player_take_light(Player* p, float out3[3]) {
    if (!p->light_requested) return 0;
    p->light_requested = false;
    for (int k = 0; k < 3; k++) out3[k] = p->light_pos[k];
    return 1;
}
int vxh_player_collided(Player* p, const float cam[3]) { for (int k = 0; k < 3; k++) p->cam_pos[k] = cam[k]; return p->collided(); }
// n calls of lightUpdate from the given angle; out = lightRotation, lightPos[3]
void vxh_light_update(long long fps, float rotation, const float start[3], int n, float out4[4]) {
    float pos[3] = {start[0], start[1], start[2]};
    for (int i = 0; i < n; i++) light_update(fps, rotation, start, pos);
    out4[0] = rotation; out4[1] = pos[0]; out4[2] = pos[1]; out4[3] = pos[2];
}
void vxh_mouse_look_matrix(float rx, float ry, float rot16[16], float dir3[3]) {
    Mat4 m;
    mouse_look_matrix(rx, ry, m, dir3);
    for (int k = 0; k < 16; k++) rot16[k] = m.m[k];
}
void vxh_mat4_rotate(float angle, float ax, float ay, float az, float out16[16]) { Mat4 r = mat4_rotate(mat4_identity(), angle, ax, ay, az); for (int k = 0; k < 16; k++) out16[k] = r.m[k]; }
}
