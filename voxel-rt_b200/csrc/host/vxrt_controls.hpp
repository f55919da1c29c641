// vxrt_controls.hpp -- host gameplay on the host mirror of the device-owned grid: the reference's
// src/controls.cpp (collision, movement, gravity, mouse look) restated without GLM / GLFW so that the game loop
// keeps working when render.cpp is replaced by the B200 path.  Pure host code, no device calls;
// the float operation order follows the reference's vendored GLM (normalize, rotate, mat*mat, vec*mat) so that
// camera trajectories match the reference bit for bit (tests/test_host_logic.py compares against the reference's
// own controls.o).
#pragma once
#include <cstdint>

namespace vxrt_host {

constexpr int PLAYER_HEIGHT = 10;          // controls.hpp:4
enum Key { KEY_W = 0, KEY_S, KEY_A, KEY_D, KEY_T, SPACE, SHIFT, LMB, RMB, KEYS };   // window.hpp:8-18

struct Mat4 { float m[16]; };              // column-major, m[4*c + r]

Mat4 mat4_identity();
Mat4 mat4_rotate(const Mat4& m, float angle, float ax, float ay, float az);   // glm::rotate, ext/matrix_transform.inl:18-46
Mat4 mat4_mul(const Mat4& a, const Mat4& b);                                  // detail/type_mat4x4.inl operator*(mat4, mat4)
void mat4_mul_vec4(const Mat4& m, const float v[4], float out[4]);            // operator*(mat4, vec4)
void vec4_mul_mat4(const float v[4], const Mat4& m, float out[4]);            // operator*(vec4, mat4)

// lightUpdate(), render.cpp:388-402: advances the sun angle (degrees; night runs 8x faster, wraps at 360) and
// rotates the start position about +z
void light_update(long long fps, float& light_rotation, const float start_light_pos[3], float light_pos[3]);
// doMouseLook's matrix for given angles (controls.cpp:137-142): rotate = rotY * rotX, dir = rotate * (0,0,1,1)
void mouse_look_matrix(float rot_x, float rot_y, Mat4& rotate, float cam_dir[3]);

struct Player {
    // state the reference keeps in globals (main.cpp:20-37, controls.cpp:8)
    float cam_pos[3] = {195, 55, 155};
    float cam_dir[3] = {0, 0, 1};
    float cam_rotation[2] = {0, 0};
    Mat4 rotate_matrix = mat4_identity();
    float gravity = 0.0f;
    int view_depth_field = 0;
    bool keys[KEYS] = {};
    long long fps = 60;
    int mouse_x = 0, mouse_y = 0, screen_w = 800, screen_h = 600;
    // the grid mirror
    const int32_t* voxels = nullptr;
    int w = 512, h = 96, d = 512;
    // set by movementUpdate when T was pressed: the caller places the light (placeLocalLight, controls.cpp:46-49)
    bool light_requested = false;
    float light_pos[3] = {0, 0, 0};

    int voxel_at(int x, int y, int z) const;   // voxels[getVoxelIndex(...)]; out of the grid counts as solid (the reference reads voxels[-1] == 0)
    int collided() const;                      // controls.cpp:10-19
    void movementUpdate();                     // controls.cpp:22-74
    void doGravity();                          // controls.cpp:77-98
    void doMouseLook();                        // controls.cpp:112-144
};

}  // namespace vxrt_host
