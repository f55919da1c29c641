// vxrt_glshim.cpp -- the LINK-LEVEL seam: libvxrt_glshim.so defines exactly the
// OpenGL / GLEW / GLFW symbols the reference's objects import (render.o: 28 __glew* function-pointer variables,
// glViewport, glShadeModel; main.o: glewInit, glDrawArrays; window.o: 14 glfw* functions) and forwards them to the
// C ABI of libvxrt.so.  Linked against it instead of -lglew32 -lglfw3 -lopengl32, the reference's six objects run
// UNMODIFIED, headless, with their per-pixel work on the B200:
//
//   reference call (file:line)                                            here
//   glShaderSource(fshader.glsl text)              render.cpp:146     -> grid extents + shader constants are read from
//                                                                        the source; a shader whose constants differ
//                                                                        from what the kernels implement is refused
//   glBufferData(SSBO, 100663296, voxels, ..)      render.cpp:201,368 -> vxrt_create (first time) + vxrt_upload_grid
//   glBufferSubData(SSBO, off, size, &voxels[off]) render.cpp:219     -> copied at call time (GL semantics), batched:
//                                                                        ONE vxrt_upload_rows before the next draw
//   glGetUniformLocation / glUniform*              render.cpp:335-341,289-296,410 -> fields of a vxrt_frame
//   glViewport(0,0,w,h)                            render.cpp:405     -> vxrt_resize before the next draw
//   glDrawArrays(GL_TRIANGLES,0,6)                 main.cpp:59        -> vxrt_set_frame + vxrt_render
//   glfwSwapBuffers                                window.cpp:174     -> vxrt_sync (+ optional PPM of chosen frames)
//   glfwPollEvents                                 window.cpp:175     -> scripted input through the reference's own
//                                                                        callbacks (keys, mouse, resize)
//   glfwWindowShouldClose                          window.cpp:161,177 -> true after VXRT_GLSHIM_FRAMES presented frames
//
// No GL header is needed: the GL / GLFW ABI types are spelled out below (GLenum = unsigned int, ...).
// There is no CPU path: if libvxrt cannot create its context (no sm_100 device) the process exits with the message.
//
// Environment (all optional):
//   VXRT_GLSHIM_FRAMES=N          close the window after N presented frames (default 600; 0 = never)
//   VXRT_GLSHIM_READY_UPLOADS=K   frames are counted (and the script runs) from the K-th whole-grid upload on; K = 2
//                                 waits for the reference's depth-field threads (render.cpp:298-301).  Default 0.
//   VXRT_GLSHIM_EVENTS="f:ev;..." input script, f = counted frame.  ev = key:<SPACE|W|A|S|D|T|SHIFT>:<down|up>,
//                                 lmb:<down|up>, rmb:<down|up>, mouse:<x>,<y>, resize:<w>x<h>
//   VXRT_GLSHIM_DUMP=fmt          printf format with one %d: PPM file of the frames listed in VXRT_GLSHIM_DUMP_FRAMES
//                                 (+ "<file>.frame": the 360 bytes of that frame's vxrt_frame, to replay it elsewhere)
//   VXRT_GLSHIM_DUMP_FRAMES=a,b   counted frames to dump (default: the last one)
//   VXRT_GLSHIM_SAVE_GRID=path    when the window closes, stream the device grid into a VXRTGRD1 file (vxrt_save_grid)
//   VXRT_GLSHIM_FPS=F             deterministic time: clock() advances CLOCKS_PER_SEC/F per frame (window.cpp:164 derives
//                                 `fps` from clock(); movement, gravity and the sun depend on it).  Unset: real clock.
//   VXRT_GLSHIM_DEVICE=D          CUDA device ordinal (default 0)
//   VXRT_GLSHIM_LOG=1             one summary line on stderr when the window closes
#include "../../../include/vxrt.h"

#include <dlfcn.h>
#include <time.h>
#include <cctype>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

// ---- GL / GLFW ABI types (khronos gl.h / glfw3.h, Linux: no calling-convention decoration) ---------------------
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned char GLboolean;
typedef float GLfloat;
typedef char GLchar;
typedef ptrdiff_t GLsizeiptr;
typedef ptrdiff_t GLintptr;
struct GLFWwindow;
struct GLFWmonitor;
typedef void (*GLFWkeyfun)(GLFWwindow*, int, int, int, int);
typedef void (*GLFWcursorposfun)(GLFWwindow*, double, double);
typedef void (*GLFWscrollfun)(GLFWwindow*, double, double);
typedef void (*GLFWmousebuttonfun)(GLFWwindow*, int, int, int);
typedef void (*GLFWframebuffersizefun)(GLFWwindow*, int, int);

namespace {

constexpr GLenum GL_SHADER_STORAGE_BUFFER_ = 0x90D2, GL_FRAGMENT_SHADER_ = 0x8B30;
constexpr int GLFW_RELEASE_ = 0, GLFW_PRESS_ = 1;

[[noreturn]] void die(const std::string& what) {
    fprintf(stderr, "vxrt_glshim: %s\n", what.c_str());
    exit(EXIT_FAILURE);
}
void check(int rc, const char* call) {
    if (rc != VXRT_OK) die(std::string(call) + ": " + vxrt_last_error());
}
long env_long(const char* name, long dflt) {
    const char* v = getenv(name);
    return (v && *v) ? strtol(v, nullptr, 10) : dflt;
}

// the constants the kernels implement (fshader.glsl:3-10, ray.cuh); checked against the shader text the host hands over
struct ShaderConst { const char* name; double want; };
const ShaderConst kShaderConsts[] = {{"RENDER_DIST", 384}, {"MAX_LOCAL_LIGHTS", 16}, {"LOCAL_LIGHT_DIST", 64},
                                     {"AMBIENT", 0.4}, {"DIFFUSE", 0.8}, {"MAX_OVERBRIGHT", 1.25}};

// "const <type> NAME = value" (fshader.glsl:3-10); false when the name is not declared that way
bool shader_constant(const std::string& src, const char* name, double* out) {
    size_t at = 0;
    const size_t len = strlen(name);
    while ((at = src.find(name, at)) != std::string::npos) {
        const size_t end = at + len;
        const bool word = (at == 0 || !(isalnum((unsigned char)src[at - 1]) || src[at - 1] == '_')) &&
                          (end >= src.size() || !(isalnum((unsigned char)src[end]) || src[end] == '_'));
        size_t p = end;
        while (p < src.size() && (src[p] == ' ' || src[p] == '\t')) p++;
        if (word && p < src.size() && src[p] == '=' && (p + 1 >= src.size() || src[p + 1] != '=')) {
            *out = strtod(src.c_str() + p + 1, nullptr);
            return true;
        }
        at = end;
    }
    return false;
}

struct Event { long frame; std::string what; };

// virtual time lives in plain constant-initialised statics: clock() below is interposed for the whole process and may be
// called by other libraries' initialisers before this file's objects are constructed
long s_fixed_fps = -1;                           // -1: VXRT_GLSHIM_FPS not read yet; 0: real clock
clock_t s_vclock = 0;
long fixed_fps() {
    if (s_fixed_fps < 0) s_fixed_fps = env_long("VXRT_GLSHIM_FPS", 0) > 0 ? env_long("VXRT_GLSHIM_FPS", 0) : 0;
    return s_fixed_fps;
}
void advance_virtual_clock() { if (fixed_fps() > 0) s_vclock += (clock_t)(CLOCKS_PER_SEC / fixed_fps()); }

struct Shim {
    // window
    int win_w = 800, win_h = 600;
    bool window_created = false;
    GLFWkeyfun on_key = nullptr;
    GLFWcursorposfun on_cursor = nullptr;
    GLFWscrollfun on_scroll = nullptr;
    GLFWmousebuttonfun on_button = nullptr;
    GLFWframebuffersizefun on_resize = nullptr;
    // GL objects
    GLuint next_name = 1;
    std::vector<GLuint> shader_kind{0};          // shader name -> GL_*_SHADER
    GLuint bound_ssbo = 0;
    int grid_w = 512, grid_h = 96, grid_d = 512; // render.hpp:4-5 unless the shader source says otherwise
    bool shader_seen = false;
    // the draw state
    vxrt_ctx* ctx = nullptr;
    vxrt_frame frame{};                          // the uniforms as the host last set them
    vxrt_frame drawn{};                          // ... as the last draw used them
    int viewport_w = 0, viewport_h = 0;          // pending glViewport (0: none)
    int ctx_w = 0, ctx_h = 0;
    long full_uploads = 0;
    // pending glBufferSubData rows of one length
    size_t row_len = 0;
    std::vector<int64_t> firsts;
    std::vector<int32_t> packed;
    unsigned long long sub_calls = 0, row_batches = 0;
    // frames
    long presented = 0;                          // counted frames (from READY_UPLOADS on)
    long warmup_frames = 0;
    bool drew_since_swap = false, drew_ready = false;
    long max_frames = 600, ready_uploads = 0;
    std::vector<Event> script;
    size_t script_at = 0;
    std::string dump_fmt;
    std::vector<long> dump_frames;
    // time
    std::chrono::steady_clock::time_point t_first_draw;
    bool timing = false;
    bool configured = false;

    void configure() {
        if (configured) return;
        configured = true;
        max_frames = env_long("VXRT_GLSHIM_FRAMES", 600);
        ready_uploads = env_long("VXRT_GLSHIM_READY_UPLOADS", 0);
        if (const char* d = getenv("VXRT_GLSHIM_DUMP")) dump_fmt = d;
        if (const char* f = getenv("VXRT_GLSHIM_DUMP_FRAMES")) {
            for (const char* p = f; *p;) {
                char* e;
                const long v = strtol(p, &e, 10);
                if (e == p) break;
                dump_frames.push_back(v);
                p = (*e == ',') ? e + 1 : e;
            }
        } else if (!dump_fmt.empty() && max_frames > 0) {
            dump_frames.push_back(max_frames - 1);
        }
        if (const char* ev = getenv("VXRT_GLSHIM_EVENTS")) {
            std::string s = ev;
            size_t at = 0;
            while (at < s.size()) {
                size_t semi = s.find(';', at);
                if (semi == std::string::npos) semi = s.size();
                const std::string item = s.substr(at, semi - at);
                const size_t colon = item.find(':');
                if (!item.empty()) {
                    if (colon == std::string::npos) die("VXRT_GLSHIM_EVENTS: '" + item + "' is not frame:event");
                    script.push_back({strtol(item.c_str(), nullptr, 10), item.substr(colon + 1)});
                }
                at = semi + 1;
            }
            for (size_t i = 1; i < script.size(); i++)
                if (script[i].frame < script[i - 1].frame) die("VXRT_GLSHIM_EVENTS: frames must not decrease");
        }
        memset(&frame, 0, sizeof frame);
    }

    bool ready() const { return full_uploads >= ready_uploads; }

    void ensure_context() {
        if (ctx) return;
        vxrt_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.grid_w = grid_w; cfg.grid_h = grid_h; cfg.grid_d = grid_d;
        cfg.width = viewport_w > 0 ? viewport_w : win_w;
        cfg.height = viewport_h > 0 ? viewport_h : win_h;
        cfg.device = (int32_t)env_long("VXRT_GLSHIM_DEVICE", 0);
        cfg.rank = 0; cfg.world = 1; cfg.flags = 0;
        check(vxrt_create(&cfg, &ctx), "vxrt_create");
        ctx_w = cfg.width; ctx_h = cfg.height;
        viewport_w = viewport_h = 0;
    }

    void flush_rows() {
        if (firsts.empty()) return;
        if (!ctx) die("glBufferSubData before the voxel buffer was created (glBufferData)");
        check(vxrt_upload_rows(ctx, firsts.size(), row_len, firsts.data(), packed.data()), "vxrt_upload_rows");
        row_batches++;
        firsts.clear(); packed.clear(); row_len = 0;
    }

    void buffer_data(GLsizeiptr size, const void* data) {
        if ((size_t)size != (size_t)grid_w * grid_h * grid_d * 4)
            die("glBufferData on the voxel buffer: " + std::to_string((long long)size) + " bytes do not match the shader's grid " +
                std::to_string(grid_w) + "x" + std::to_string(grid_h) + "x" + std::to_string(grid_d));
        firsts.clear(); packed.clear(); row_len = 0;      // a whole-buffer respecification supersedes pending rows
        ensure_context();
        if (data) check(vxrt_upload_grid(ctx, (const int32_t*)data, (size_t)size / 4), "vxrt_upload_grid");
        full_uploads++;
    }

    void buffer_sub_data(GLintptr offset, GLsizeiptr size, const void* data) {
        sub_calls++;
        if (!data || size <= 0) return;
        if (offset % 4 || size % 4) die("glBufferSubData on the voxel buffer: offset / size not a multiple of 4 bytes");
        const int64_t first = (int64_t)(offset / 4);
        const size_t count = (size_t)size / 4;
        bool clash = !firsts.empty() && count != row_len;
        for (size_t r = 0; !clash && r < firsts.size(); r++)          // a later write to the same voxels must stay later
            clash = first < firsts[r] + (int64_t)row_len && firsts[r] < first + (int64_t)count;
        if (clash) flush_rows();
        row_len = count;
        firsts.push_back(first);
        const int32_t* src = (const int32_t*)data;
        packed.insert(packed.end(), src, src + count);               // GL copies at call time; so do we
    }

    void draw() {
        if (!ctx) die("glDrawArrays before the voxel buffer exists");
        flush_rows();
        if (viewport_w > 0 && (viewport_w != ctx_w || viewport_h != ctx_h)) {
            check(vxrt_resize(ctx, viewport_w, viewport_h), "vxrt_resize");
            ctx_w = viewport_w; ctx_h = viewport_h;
        }
        viewport_w = viewport_h = 0;
        // vxrt_resize sets aspect = w/h like reshape() does; the host's own uniform value is what the shader sees
        drawn = frame;                                               // the uniforms move on before the frame is presented
        check(vxrt_set_frame(ctx, &frame), "vxrt_set_frame");
        check(vxrt_render(ctx), "vxrt_render");
        drew_since_swap = true;
        drew_ready = ready();
        if (!timing && ready()) { timing = true; t_first_draw = std::chrono::steady_clock::now(); }
    }

    void present() {
        if (!ctx || !drew_since_swap) return;                        // window.cpp swaps once before the first draw
        drew_since_swap = false;
        check(vxrt_sync(ctx), "vxrt_sync");
        if (!drew_ready) { warmup_frames++; return; }                // drawn before the grid was ready: not counted
        for (long f : dump_frames)
            if (f == presented && !dump_fmt.empty()) {
                char path[4096];
                snprintf(path, sizeof path, dump_fmt.c_str(), (int)presented);
                check(vxrt_write_ppm(ctx, path), "vxrt_write_ppm");
                const std::string side = std::string(path) + ".frame";
                if (FILE* fp = fopen(side.c_str(), "wb")) { fwrite(&drawn, sizeof drawn, 1, fp); fclose(fp); }
            }
        presented++;
    }

    static int key_code(const std::string& k) {
        if (k == "SPACE") return 32;
        if (k == "SHIFT") return 340;                                // GLFW_KEY_LEFT_SHIFT
        if (k.size() == 1 && strchr("WASDT", k[0])) return k[0];     // GLFW_KEY_<letter> == ASCII
        die("VXRT_GLSHIM_EVENTS: unknown key '" + k + "'");
    }
    static int action_code(const std::string& a) {
        if (a == "down") return GLFW_PRESS_;
        if (a == "up") return GLFW_RELEASE_;
        die("VXRT_GLSHIM_EVENTS: action must be down or up, not '" + a + "'");
    }

    void run_event(const std::string& ev) {
        GLFWwindow* w = (GLFWwindow*)this;
        const size_t c1 = ev.find(':');
        const std::string kind = ev.substr(0, c1), rest = c1 == std::string::npos ? "" : ev.substr(c1 + 1);
        if (kind == "key") {
            const size_t c2 = rest.find(':');
            if (c2 == std::string::npos) die("VXRT_GLSHIM_EVENTS: key:<name>:<down|up>");
            if (on_key) on_key(w, key_code(rest.substr(0, c2)), 0, action_code(rest.substr(c2 + 1)), 0);
        } else if (kind == "lmb" || kind == "rmb") {
            if (on_button) on_button(w, kind == "lmb" ? 0 : 1, action_code(rest), 0);
        } else if (kind == "mouse") {
            double x = 0, y = 0;
            if (sscanf(rest.c_str(), "%lf,%lf", &x, &y) != 2) die("VXRT_GLSHIM_EVENTS: mouse:<x>,<y>");
            if (on_cursor) on_cursor(w, x, y);
        } else if (kind == "resize") {
            int rw = 0, rh = 0;
            if (sscanf(rest.c_str(), "%dx%d", &rw, &rh) != 2 || rw <= 0 || rh <= 0) die("VXRT_GLSHIM_EVENTS: resize:<w>x<h>");
            win_w = rw; win_h = rh;
            if (on_resize) on_resize(w, rw, rh);
        } else {
            die("VXRT_GLSHIM_EVENTS: unknown event '" + ev + "'");
        }
    }

    void poll() {
        advance_virtual_clock();
        if (!ready()) return;
        while (script_at < script.size() && script[script_at].frame <= presented) run_event(script[script_at++].what);
    }

    void finish();
};

Shim g;

// The window closes: the reference's main() returns right after (main.cpp:58,74).  Everything that needs the device is
// done here, not in an exit handler -- the CUDA runtime may already be shutting down by then.
void Shim::finish() {
    if (!ctx) return;
    if (env_long("VXRT_GLSHIM_LOG", 0)) {
        vxrt_sync(ctx);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_first_draw).count();
        fprintf(stderr, "vxrt_glshim: %ld frames at %dx%d (+%ld before the grid was ready), %.3f ms/frame host loop, "
                        "%ld whole-grid uploads, %llu glBufferSubData calls in %llu batches\n",
                presented, ctx_w, ctx_h, warmup_frames, presented ? 1e3 * s / presented : 0.0, full_uploads, sub_calls, row_batches);
    }
    if (const char* path = getenv("VXRT_GLSHIM_SAVE_GRID")) check(vxrt_save_grid(ctx, path), "vxrt_save_grid");
    vxrt_destroy(ctx);
    ctx = nullptr;
}

// uniform "locations" = which vxrt_frame field (fshader.glsl:20-26)
enum Uniform { U_CAM_POS = 0, U_CAM_ROTATION, U_LIGHT_POS, U_ASPECT, U_ROTATE, U_VIEW, U_LIGHTS, U_COUNT };
const char* const kUniformNames[U_COUNT] = {"camPos", "camRotation", "lightPos", "aspectRatio", "rotateMatrix",
                                             "viewDepthField", "localLights"};

// ---- GL entry points --------------------------------------------------------------------------------------------
void s_AttachShader(GLuint, GLuint) {}
void s_BindBuffer(GLenum target, GLuint name) { if (target == GL_SHADER_STORAGE_BUFFER_) g.bound_ssbo = name; }
void s_BindBufferBase(GLenum, GLuint, GLuint) {}
void s_BindVertexArray(GLuint) {}
void s_BufferData(GLenum target, GLsizeiptr size, const void* data, GLenum) {
    g.configure();
    if (target == GL_SHADER_STORAGE_BUFFER_) g.buffer_data(size, data);       // the quad's GL_ARRAY_BUFFER is not needed
}
void s_BufferSubData(GLenum target, GLintptr offset, GLsizeiptr size, const void* data) {
    if (target == GL_SHADER_STORAGE_BUFFER_) g.buffer_sub_data(offset, size, data);
}
void s_CompileShader(GLuint) {}
GLuint s_CreateProgram() { return g.next_name++; }
GLuint s_CreateShader(GLenum kind) {
    const GLuint name = g.next_name++;
    if (g.shader_kind.size() <= name) g.shader_kind.resize(name + 1, 0);
    g.shader_kind[name] = kind;
    return name;
}
void s_EnableVertexAttribArray(GLuint) {}
void s_GenBuffers(GLsizei n, GLuint* out) { for (GLsizei i = 0; i < n; i++) out[i] = g.next_name++; }
void s_GenVertexArrays(GLsizei n, GLuint* out) { for (GLsizei i = 0; i < n; i++) out[i] = g.next_name++; }
GLint s_GetAttribLocation(GLuint, const GLchar*) { return 0; }
void s_GetProgramInfoLog(GLuint, GLsizei, GLsizei* len, GLchar* log) { if (len) *len = 0; if (log) *log = 0; }
void s_GetProgramiv(GLuint, GLenum, GLint* out) { *out = 1; }                  // GL_LINK_STATUS: linked
void s_GetShaderInfoLog(GLuint, GLsizei, GLsizei* len, GLchar* log) { if (len) *len = 0; if (log) *log = 0; }
void s_GetShaderiv(GLuint, GLenum, GLint* out) { *out = 1; }                   // GL_COMPILE_STATUS: compiled
GLint s_GetUniformLocation(GLuint, const GLchar* name) {
    for (int u = 0; u < U_COUNT; u++) if (!strcmp(name, kUniformNames[u])) return u;
    return -1;                                                                 // GL: not an active uniform
}
void s_LinkProgram(GLuint) {}
// The fragment shader's text carries the grid extents (fshader.glsl:3-4) and the constants the kernels hard-code.
void s_ShaderSource(GLuint shader, GLsizei count, const GLchar* const* strings, const GLint* lengths) {
    if (shader >= g.shader_kind.size() || g.shader_kind[shader] != GL_FRAGMENT_SHADER_) return;
    std::string src;
    for (GLsizei i = 0; i < count; i++) src += (lengths && lengths[i] >= 0) ? std::string(strings[i], lengths[i]) : std::string(strings[i]);
    double v;
    if (shader_constant(src, "VOXELS_WIDTH", &v)) { g.grid_w = (int)v; g.grid_d = (int)v; }
    if (shader_constant(src, "VOXELS_HEIGHT", &v)) g.grid_h = (int)v;
    for (const ShaderConst& c : kShaderConsts)
        if (shader_constant(src, c.name, &v) && (float)v != (float)c.want)
            die(std::string("the fragment shader sets ") + c.name + " = " + std::to_string(v) + ", the CUDA path implements " +
                std::to_string(c.want) + " (fshader.glsl:3-10); refusing to render something else than the shader would");
    g.shader_seen = true;
}
void s_Uniform1f(GLint l, GLfloat a) { if (l == U_ASPECT) g.frame.aspect = a; }
void s_Uniform1i(GLint l, GLint a) { if (l == U_VIEW) g.frame.view_depth_field = a; }
void s_Uniform2f(GLint l, GLfloat a, GLfloat b) { if (l == U_CAM_ROTATION) { g.frame.cam_rotation[0] = a; g.frame.cam_rotation[1] = b; } }
void s_Uniform3f(GLint l, GLfloat a, GLfloat b, GLfloat c) {
    float* dst = l == U_CAM_POS ? g.frame.cam_pos : l == U_LIGHT_POS ? g.frame.light_pos : nullptr;
    if (dst) { dst[0] = a; dst[1] = b; dst[2] = c; }
}
void s_Uniform4fv(GLint l, GLsizei n, const GLfloat* v) {
    if (l == U_LIGHTS && n > 0) memcpy(g.frame.lights, v, sizeof(float) * 4 * (size_t)(n < VXRT_MAX_LOCAL_LIGHTS ? n : VXRT_MAX_LOCAL_LIGHTS));
}
void s_UniformMatrix4fv(GLint l, GLsizei n, GLboolean transpose, const GLfloat* v) {
    if (l != U_ROTATE || n < 1) return;
    if (!transpose) { memcpy(g.frame.rotate, v, sizeof(float) * 16); return; }   // render.cpp:294 passes GL_FALSE
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) g.frame.rotate[c * 4 + r] = v[r * 4 + c];
}
void s_UseProgram(GLuint) {}
void s_VertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}

}  // namespace

extern "C" {

// ---- the GLEW function-pointer variables render.o imports (GL/glew.h declares them as PFNGL...PROC) --------------
void (*__glewAttachShader)(GLuint, GLuint) = s_AttachShader;
void (*__glewBindBuffer)(GLenum, GLuint) = s_BindBuffer;
void (*__glewBindBufferBase)(GLenum, GLuint, GLuint) = s_BindBufferBase;
void (*__glewBindVertexArray)(GLuint) = s_BindVertexArray;
void (*__glewBufferData)(GLenum, GLsizeiptr, const void*, GLenum) = s_BufferData;
void (*__glewBufferSubData)(GLenum, GLintptr, GLsizeiptr, const void*) = s_BufferSubData;
void (*__glewCompileShader)(GLuint) = s_CompileShader;
GLuint (*__glewCreateProgram)(void) = s_CreateProgram;
GLuint (*__glewCreateShader)(GLenum) = s_CreateShader;
void (*__glewEnableVertexAttribArray)(GLuint) = s_EnableVertexAttribArray;
void (*__glewGenBuffers)(GLsizei, GLuint*) = s_GenBuffers;
void (*__glewGenVertexArrays)(GLsizei, GLuint*) = s_GenVertexArrays;
GLint (*__glewGetAttribLocation)(GLuint, const GLchar*) = s_GetAttribLocation;
void (*__glewGetProgramInfoLog)(GLuint, GLsizei, GLsizei*, GLchar*) = s_GetProgramInfoLog;
void (*__glewGetProgramiv)(GLuint, GLenum, GLint*) = s_GetProgramiv;
void (*__glewGetShaderInfoLog)(GLuint, GLsizei, GLsizei*, GLchar*) = s_GetShaderInfoLog;
void (*__glewGetShaderiv)(GLuint, GLenum, GLint*) = s_GetShaderiv;
GLint (*__glewGetUniformLocation)(GLuint, const GLchar*) = s_GetUniformLocation;
void (*__glewLinkProgram)(GLuint) = s_LinkProgram;
void (*__glewShaderSource)(GLuint, GLsizei, const GLchar* const*, const GLint*) = s_ShaderSource;
void (*__glewUniform1f)(GLint, GLfloat) = s_Uniform1f;
void (*__glewUniform1i)(GLint, GLint) = s_Uniform1i;
void (*__glewUniform2f)(GLint, GLfloat, GLfloat) = s_Uniform2f;
void (*__glewUniform3f)(GLint, GLfloat, GLfloat, GLfloat) = s_Uniform3f;
void (*__glewUniform4fv)(GLint, GLsizei, const GLfloat*) = s_Uniform4fv;
void (*__glewUniformMatrix4fv)(GLint, GLsizei, GLboolean, const GLfloat*) = s_UniformMatrix4fv;
void (*__glewUseProgram)(GLuint) = s_UseProgram;
void (*__glewVertexAttribPointer)(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) = s_VertexAttribPointer;

GLenum glewInit(void) { g.configure(); return 0; }                             // GLEW_OK
void glDrawArrays(GLenum, GLint, GLsizei) { g.configure(); g.draw(); }         // main.cpp:59: the full-screen quad
void glShadeModel(GLenum) {}
void glViewport(GLint, GLint, GLsizei w, GLsizei h) { if (w > 0 && h > 0) { g.viewport_w = w; g.viewport_h = h; } }

// ---- GLFW (window.cpp:130-178) -------------------------------------------------------------------------------------
int glfwInit(void) {
    g.configure();
    advance_virtual_clock();                     // window.cpp:132 read clock() just before: the first frame sees one tick
    return 1;
}
void glfwTerminate(void) {}
GLFWwindow* glfwCreateWindow(int w, int h, const char*, GLFWmonitor*, GLFWwindow*) {
    g.configure();
    if (w > 0 && h > 0) { g.win_w = w; g.win_h = h; }
    g.window_created = true;
    return (GLFWwindow*)&g;
}
GLFWmonitor* glfwGetPrimaryMonitor(void) { return nullptr; }
void glfwMakeContextCurrent(GLFWwindow*) {}
void glfwSwapInterval(int) {}
GLFWkeyfun glfwSetKeyCallback(GLFWwindow*, GLFWkeyfun f) { GLFWkeyfun o = g.on_key; g.on_key = f; return o; }
GLFWcursorposfun glfwSetCursorPosCallback(GLFWwindow*, GLFWcursorposfun f) { GLFWcursorposfun o = g.on_cursor; g.on_cursor = f; return o; }
GLFWscrollfun glfwSetScrollCallback(GLFWwindow*, GLFWscrollfun f) { GLFWscrollfun o = g.on_scroll; g.on_scroll = f; return o; }
GLFWmousebuttonfun glfwSetMouseButtonCallback(GLFWwindow*, GLFWmousebuttonfun f) { GLFWmousebuttonfun o = g.on_button; g.on_button = f; return o; }
GLFWframebuffersizefun glfwSetFramebufferSizeCallback(GLFWwindow*, GLFWframebuffersizefun f) { GLFWframebuffersizefun o = g.on_resize; g.on_resize = f; return o; }
void glfwSwapBuffers(GLFWwindow*) { g.present(); }
void glfwPollEvents(void) { g.poll(); }
int glfwWindowShouldClose(GLFWwindow*) {
    const bool close = g.max_frames > 0 && g.presented >= g.max_frames;
    if (close) g.finish();
    return close;
}

// ---- deterministic time (only with VXRT_GLSHIM_FPS): window.cpp:164 computes fps = CLOCKS_PER_SEC / (clock() - start) --
clock_t clock(void) noexcept {
    if (fixed_fps() > 0) return s_vclock;
    static clock_t (*real_clock)(void) = (clock_t(*)(void))dlsym(RTLD_NEXT, "clock");
    return real_clock ? real_clock() : (clock_t)-1;
}

}  // extern "C"
