// ref_host_shim.cpp -- TEST INFRASTRUCTURE.  Lets the reference's OWN host code (level.cpp, render.cpp,
// controls.cpp, window.cpp, Entity.cpp, main.cpp -- compiled unmodified from /root/reference/src by
// oracle/ref_build/Makefile) link and run headless: every GL / GLEW / GLFW symbol those objects
// import is defined here as a recording no-op, and a small extern "C" driver exposes the reference's
// globals and functions to the tests.  Nothing here is reference code; outputs go to oracle/_ref/.
//
// What is recorded: glBufferData / glBufferSubData calls on GL_SHADER_STORAGE_BUFFER (the grid
// upload pattern, render.cpp:199-223,366-369) and every glUniform* value by location (the frame
// parameters, render.cpp:289-296).
#include "window.hpp"
#include "render.hpp"
#include "controls.hpp"
#include "level.hpp"
#include <unistd.h>
#include <cstring>
#include <cstdio>
#include <vector>
#include <string>

// ---- recording state ------------------------------------------------------------------------
struct SubCall { long long offset, size; };
static std::vector<SubCall> g_sub_calls;
static int g_full_uploads = 0;                 // glBufferData on the SSBO target with >= 1 MB
static std::vector<std::string> g_uniform_names;
static float g_uniform_vals[16][64];
static int g_uniform_ints[16];
static int g_draws = 0;

// ---- GLEW function-pointer variables the reference objects import ----------------------------
static void GLAPIENTRY s_AttachShader(GLuint, GLuint) {}
static void GLAPIENTRY s_BindBuffer(GLenum, GLuint) {}
static void GLAPIENTRY s_BindBufferBase(GLenum, GLuint, GLuint) {}
static void GLAPIENTRY s_BindVertexArray(GLuint) {}
static void GLAPIENTRY s_BufferData(GLenum target, GLsizeiptr size, const void*, GLenum) {
    if (target == GL_SHADER_STORAGE_BUFFER && size >= (1 << 20)) g_full_uploads++;
}
static void GLAPIENTRY s_BufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void*) {
    if (target == GL_SHADER_STORAGE_BUFFER) g_sub_calls.push_back({(long long)offset, (long long)size});
}
static void GLAPIENTRY s_CompileShader(GLuint) {}
static GLuint GLAPIENTRY s_CreateProgram() { return 1; }
static GLuint GLAPIENTRY s_CreateShader(GLenum) { return 1; }
static void GLAPIENTRY s_EnableVertexAttribArray(GLuint) {}
static void GLAPIENTRY s_GenBuffers(GLsizei n, GLuint* b) { for (GLsizei i = 0; i < n; i++) b[i] = 1 + i; }
static void GLAPIENTRY s_GenVertexArrays(GLsizei n, GLuint* b) { for (GLsizei i = 0; i < n; i++) b[i] = 1 + i; }
static GLint GLAPIENTRY s_GetAttribLocation(GLuint, const GLchar*) { return 0; }
static void GLAPIENTRY s_GetProgramInfoLog(GLuint, GLsizei, GLsizei*, GLchar*) {}
static void GLAPIENTRY s_GetProgramiv(GLuint, GLenum, GLint* p) { *p = 1; }
static void GLAPIENTRY s_GetShaderInfoLog(GLuint, GLsizei, GLsizei*, GLchar*) {}
static void GLAPIENTRY s_GetShaderiv(GLuint, GLenum, GLint* p) { *p = 1; }
static GLint GLAPIENTRY s_GetUniformLocation(GLuint, const GLchar* name) {
    for (size_t i = 0; i < g_uniform_names.size(); i++) if (g_uniform_names[i] == name) return (GLint)i;
    g_uniform_names.push_back(name);
    return (GLint)g_uniform_names.size() - 1;
}
static void GLAPIENTRY s_LinkProgram(GLuint) {}
static void GLAPIENTRY s_ShaderSource(GLuint, GLsizei, const GLchar* const*, const GLint*) {}
static void GLAPIENTRY s_Uniform1f(GLint l, GLfloat a) { if (l >= 0 && l < 16) g_uniform_vals[l][0] = a; }
static void GLAPIENTRY s_Uniform1i(GLint l, GLint a) { if (l >= 0 && l < 16) g_uniform_ints[l] = a; }
static void GLAPIENTRY s_Uniform2f(GLint l, GLfloat a, GLfloat b) { if (l >= 0 && l < 16) { g_uniform_vals[l][0] = a; g_uniform_vals[l][1] = b; } }
static void GLAPIENTRY s_Uniform3f(GLint l, GLfloat a, GLfloat b, GLfloat c) { if (l >= 0 && l < 16) { g_uniform_vals[l][0] = a; g_uniform_vals[l][1] = b; g_uniform_vals[l][2] = c; } }
static void GLAPIENTRY s_Uniform4fv(GLint l, GLsizei n, const GLfloat* v) { if (l >= 0 && l < 16 && n <= 16) memcpy(g_uniform_vals[l], v, sizeof(float) * 4 * n); }
static void GLAPIENTRY s_UniformMatrix4fv(GLint l, GLsizei, GLboolean, const GLfloat* v) { if (l >= 0 && l < 16) memcpy(g_uniform_vals[l], v, sizeof(float) * 16); }
static void GLAPIENTRY s_UseProgram(GLuint) {}
static void GLAPIENTRY s_VertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}

extern "C" {
PFNGLATTACHSHADERPROC __glewAttachShader = s_AttachShader;
PFNGLBINDBUFFERPROC __glewBindBuffer = s_BindBuffer;
PFNGLBINDBUFFERBASEPROC __glewBindBufferBase = s_BindBufferBase;
PFNGLBINDVERTEXARRAYPROC __glewBindVertexArray = s_BindVertexArray;
PFNGLBUFFERDATAPROC __glewBufferData = s_BufferData;
PFNGLBUFFERSUBDATAPROC __glewBufferSubData = s_BufferSubData;
PFNGLCOMPILESHADERPROC __glewCompileShader = s_CompileShader;
PFNGLCREATEPROGRAMPROC __glewCreateProgram = s_CreateProgram;
PFNGLCREATESHADERPROC __glewCreateShader = s_CreateShader;
PFNGLENABLEVERTEXATTRIBARRAYPROC __glewEnableVertexAttribArray = s_EnableVertexAttribArray;
PFNGLGENBUFFERSPROC __glewGenBuffers = s_GenBuffers;
PFNGLGENVERTEXARRAYSPROC __glewGenVertexArrays = s_GenVertexArrays;
PFNGLGETATTRIBLOCATIONPROC __glewGetAttribLocation = s_GetAttribLocation;
PFNGLGETPROGRAMINFOLOGPROC __glewGetProgramInfoLog = s_GetProgramInfoLog;
PFNGLGETPROGRAMIVPROC __glewGetProgramiv = s_GetProgramiv;
PFNGLGETSHADERINFOLOGPROC __glewGetShaderInfoLog = s_GetShaderInfoLog;
PFNGLGETSHADERIVPROC __glewGetShaderiv = s_GetShaderiv;
PFNGLGETUNIFORMLOCATIONPROC __glewGetUniformLocation = s_GetUniformLocation;
PFNGLLINKPROGRAMPROC __glewLinkProgram = s_LinkProgram;
PFNGLSHADERSOURCEPROC __glewShaderSource = s_ShaderSource;
PFNGLUNIFORM1FPROC __glewUniform1f = s_Uniform1f;
PFNGLUNIFORM1IPROC __glewUniform1i = s_Uniform1i;
PFNGLUNIFORM2FPROC __glewUniform2f = s_Uniform2f;
PFNGLUNIFORM3FPROC __glewUniform3f = s_Uniform3f;
PFNGLUNIFORM4FVPROC __glewUniform4fv = s_Uniform4fv;
PFNGLUNIFORMMATRIX4FVPROC __glewUniformMatrix4fv = s_UniformMatrix4fv;
PFNGLUSEPROGRAMPROC __glewUseProgram = s_UseProgram;
PFNGLVERTEXATTRIBPOINTERPROC __glewVertexAttribPointer = s_VertexAttribPointer;

GLenum GLEWAPIENTRY glewInit(void) { return 0; }
void GLAPIENTRY glDrawArrays(GLenum, GLint, GLsizei) { g_draws++; }
void GLAPIENTRY glShadeModel(GLenum) {}
void GLAPIENTRY glViewport(GLint, GLint, GLsizei, GLsizei) {}

// ---- GLFW ------------------------------------------------------------------------------------
GLFWwindow* glfwCreateWindow(int, int, const char*, GLFWmonitor*, GLFWwindow*) { return (GLFWwindow*)1; }
GLFWmonitor* glfwGetPrimaryMonitor(void) { return nullptr; }
int glfwInit(void) { return 1; }
void glfwMakeContextCurrent(GLFWwindow*) {}
void glfwPollEvents(void) {}
GLFWcursorposfun glfwSetCursorPosCallback(GLFWwindow*, GLFWcursorposfun) { return nullptr; }
GLFWframebuffersizefun glfwSetFramebufferSizeCallback(GLFWwindow*, GLFWframebuffersizefun) { return nullptr; }
GLFWkeyfun glfwSetKeyCallback(GLFWwindow*, GLFWkeyfun) { return nullptr; }
GLFWmousebuttonfun glfwSetMouseButtonCallback(GLFWwindow*, GLFWmousebuttonfun) { return nullptr; }
GLFWscrollfun glfwSetScrollCallback(GLFWwindow*, GLFWscrollfun) { return nullptr; }
void glfwSwapBuffers(GLFWwindow*) {}
void glfwSwapInterval(int) {}
void glfwTerminate(void) {}
int glfwWindowShouldClose(GLFWwindow*) { return 1; }
}

// ---- driver ----------------------------------------------------------------------------------
extern "C" {

int ref_host_voxel_count(void) { return VOXELS_WIDTH * VOXELS_HEIGHT * VOXELS_WIDTH; }
int* ref_host_voxels(void) { return voxels; }

// render.cpp:349-352 without the depth threads: fill -1, then the reference's initVoxels()
void ref_host_level_nodepth(void) {
    for (int i = 0; i < VOXELS_WIDTH * VOXELS_HEIGHT * VOXELS_WIDTH; i++) voxels[i] = -1;
    initVoxels();
}

// initRender() (needs CWD = the reference's src/ because InitShader opens "vshader.glsl" relatively,
// render.cpp:326) then poll updateUniforms() until the "depth field finished" full re-upload
// (render.cpp:298-301) is observed.  Returns the number of full SSBO uploads seen (expected 2).
int ref_host_init(const char* ref_src_dir) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(ref_src_dir) != 0) return -2;
    g_full_uploads = 0;
    initRender();
    if (chdir(cwd) != 0) return -3;
    while (g_full_uploads < 2) { updateUniforms(); usleep(20000); }
    return g_full_uploads;
}

// controls.cpp:100-110 through the reference's own doDestroy(); returns the number of
// glBufferSubData calls it made and copies (offset,size) byte pairs out.
int ref_host_do_destroy(const float cam[3], const float dir[3], long long* offsets, long long* sizes, int max_calls) {
    camPos = glm::vec3(cam[0], cam[1], cam[2]);
    camDir = glm::vec3(dir[0], dir[1], dir[2]);
    keys[RMB] = true;
    g_sub_calls.clear();
    doDestroy();
    int n = (int)g_sub_calls.size();
    for (int i = 0; i < n && i < max_calls; i++) { offsets[i] = g_sub_calls[i].offset; sizes[i] = g_sub_calls[i].size; }
    return n;
}

int ref_host_partial(const float s[3], const float e[3], long long* offsets, long long* sizes, int max_calls) {
    g_sub_calls.clear();
    updatePartialGeometry(glm::vec3(s[0], s[1], s[2]), glm::vec3(e[0], e[1], e[2]));
    int n = (int)g_sub_calls.size();
    for (int i = 0; i < n && i < max_calls; i++) { offsets[i] = g_sub_calls[i].offset; sizes[i] = g_sub_calls[i].size; }
    return n;
}

void ref_host_remove_sphere(int x, int y, int z, int r) { removeSphere(glm::ivec3(x, y, z), r); }
void ref_host_fix_depth_field(int x, int y, int z) { fixDepthField(x, y, z); }
int ref_host_get_voxel_index(int x, int y, int z) { return getVoxelIndex(x, y, z); }
void ref_host_place_voxel(int x, int y, int z, int v) { placeVoxel(x, y, z, v); }
void ref_host_destroy_voxel(int x, int y, int z) { destroyVoxel(x, y, z); }
void ref_host_place_light(float x, float y, float z, float w) { placeLocalLight(x, y, z, w); }
void ref_host_get_lights(float* out64) { memcpy(out64, glm::value_ptr(*localLights), sizeof(float) * 64); }

// Set the reference's globals, call its updateUniforms(), return what reached "GL":
// out = camPos[3] camRotation[2] lightPos[3] aspect rotate[16] lights[64]  (89 floats) ; returns viewDepthField
int ref_host_update_uniforms(const float cam[3], const float camrot[2], const float light[3], float aspect,
                             const float rot16[16], int view, float* out89) {
    camPos = glm::vec3(cam[0], cam[1], cam[2]);
    camRotation = glm::vec2(camrot[0], camrot[1]);
    lightPos = glm::vec3(light[0], light[1], light[2]);
    aspectRatio = aspect;
    memcpy(glm::value_ptr(rotateMatrix), rot16, sizeof(float) * 16);
    viewDepthField = view;
    updateUniforms();
    auto loc = [](const char* n) { return s_GetUniformLocation(0, n); };
    float* o = out89;
    memcpy(o, g_uniform_vals[loc("camPos")], 12); o += 3;
    memcpy(o, g_uniform_vals[loc("camRotation")], 8); o += 2;
    memcpy(o, g_uniform_vals[loc("lightPos")], 12); o += 3;
    *o++ = g_uniform_vals[loc("aspectRatio")][0];
    memcpy(o, g_uniform_vals[loc("rotateMatrix")], 64); o += 16;
    memcpy(o, g_uniform_vals[loc("localLights")], 256);
    return g_uniform_ints[loc("viewDepthField")];
}

// doMouseLook's matrix (controls.cpp:137-142) for given camRotation: rotY * rotX, and camDir
void ref_host_mouse_look_matrix(float rx, float ry, float* rot16, float* dir3) {
    glm::mat4 rotX = glm::rotate(glm::mat4(1.0f), rx, glm::vec3(1, 0, 0));
    glm::mat4 rotY = glm::rotate(glm::mat4(1.0f), ry, glm::vec3(0, 1, 0));
    glm::mat4 m = rotY * rotX;
    glm::vec3 d = m * glm::vec4(0, 0, 1, 1);
    memcpy(rot16, glm::value_ptr(m), 64);
    dir3[0] = d.x; dir3[1] = d.y; dir3[2] = d.z;
}

// ---- gameplay (controls.cpp:10-98, 112-144) through the reference's own functions ----------------------------
}  // extern "C"
int collided();   // controls.cpp:10, external linkage but not in controls.hpp
void initLocalLights();   // render.cpp:304, likewise
extern "C" {
// controls.cpp keeps `gravity` in a file-static; the only way to zero it is the reference's own "hit something"
// branch of doGravity(): stand the camera in a temporarily solid voxel.
void ref_host_reset_gravity(void) {
    const glm::vec3 keep = camPos;
    const int idx = getVoxelIndex(5, 5, 5), old = voxels[idx];
    voxels[idx] = 0;
    camPos = glm::vec3(5.5f, 11.5f, 5.5f);
    doGravity();
    voxels[idx] = old;
    camPos = keep;
}
void ref_host_player_set(const float cam[3], const float dir[3], const float camrot[2], long long fps_) {
    camPos = glm::vec3(cam[0], cam[1], cam[2]);
    camDir = glm::vec3(dir[0], dir[1], dir[2]);
    camRotation = glm::vec2(camrot[0], camrot[1]);
    rotateMatrix = glm::mat4(1.0f);
    viewDepthField = 0;
    fps = fps_;
    for (int k = 0; k < KEYS; k++) keys[k] = false;
}
void ref_host_player_keys(const unsigned char k9[9]) { for (int k = 0; k < KEYS; k++) keys[k] = k9[k] != 0; }
void ref_host_player_mouse(int mx, int my, int sw, int sh) { mouseX = mx; mouseY = my; screenWidth = sw; screenHeight = sh; }
// the host part of one main-loop iteration (main.cpp:62-65)
void ref_host_player_step(void) { movementUpdate(); doMouseLook(); doGravity(); }
int ref_host_player_get(float out24[24]) {
    out24[0] = camPos.x; out24[1] = camPos.y; out24[2] = camPos.z;
    out24[3] = camDir.x; out24[4] = camDir.y; out24[5] = camDir.z;
    out24[6] = camRotation.x; out24[7] = camRotation.y;
    memcpy(out24 + 8, glm::value_ptr(rotateMatrix), 64);
    return viewDepthField;
}
// n calls of the reference's lightUpdate() (render.cpp:388-402) from a given angle; out = lightRotation, lightPos[3]
void ref_host_light_update(long long fps_, float rotation, int n, float out4[4]) {
    fps = fps_;
    lightRotation = rotation;
    for (int i = 0; i < n; i++) lightUpdate();
    out4[0] = lightRotation; out4[1] = lightPos.x; out4[2] = lightPos.y; out4[3] = lightPos.z;
}
void ref_host_init_lights(void) { initLocalLights(); }
int ref_host_collided(const float cam[3]) { camPos = glm::vec3(cam[0], cam[1], cam[2]); return collided(); }

}  // extern "C"
