// ref_shader_compat.cpp -- TEST INFRASTRUCTURE.  Compiles the reference's fragment shader,
// /root/reference/src/fshader.glsl, UNMODIFIED as C++: the Makefile pipes the file (minus its
// "#version 430" line, which is not a C++ preprocessing directive) into the compiler as
// VXRT_SHADER_SOURCE, and this prelude maps the GLSL vocabulary onto the reference's own vendored
// GLM (vec3/ivec3/mat4, normalize, length, dot, sign, abs, intBitsToFloat, swizzles).  The result is
// "the reference shader run on the CPU": strict IEEE binary32 (-ffp-contract=off), GLM's operation
// order for normalize / length / mat4*vec4.  Outputs go to oracle/_ref/ only.
//
// GLSL features GLM lacks, added here (semantics = GLSL's implicit int->float conversion):
//   max(int, float)            fshader.glsl:155,177   "max(0, dot(..))"
//   ivec3 - vec3               fshader.glsl:79,120    "(currCheck + forwardSteps - startPosition)"
#define GLM_FORCE_SWIZZLE
#define _MSC_EXTENSIONS 1      // only effect: GLM enables anonymous-struct swizzle members (.xyz) on gcc
#include <glm/glm.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

using namespace glm;
static inline float max(int a, float b) { float fa = (float)a; return (fa < b) ? b : fa; }
static inline vec3 operator-(ivec3 const& a, vec3 const& b) { return vec3(a) - b; }

#define layout(...)
#define buffer static union
#define voxelBuffer
#define uniform
#define in
#define out
#define main shader_main
#include VXRT_SHADER_SOURCE
#undef main
#undef in
#undef out
#undef uniform
#undef buffer
#undef voxelBuffer
#undef layout

static void reset_invocation() {           // GLSL globals are initialised per invocation (fshader.glsl:28-31)
    hitPos = vec3(0, 0, 0);
    hitNormal = vec3(0, 0, 0);
    stepCount = 0.0f;
}

static void set_vpos(int px, int py, int width, int height) {
    // vshader.glsl:6-9 + quad render.cpp:36-44: vPos = NDC of the pixel centre (same formula as the oracle)
    float vx = ((float)px + 0.5f) / (float)width * 2.0f - 1.0f;
    float vy = ((float)py + 0.5f) / (float)height * 2.0f - 1.0f;
    vPos = vec4(vx, vy, 0.0f, 1.0f);
}

extern "C" {

int ref_shader_voxel_count(void) { return (int)(sizeof(voxels) / sizeof(int)); }
void ref_shader_upload(const int32_t* src) { memcpy(voxels, src, sizeof(voxels)); }       // == glBufferData
void ref_shader_upload_range(long long first, long long count, const int32_t* src) {       // == glBufferSubData
    memcpy(voxels + first, src, (size_t)count * 4);
}

// frame89 = camPos[3] camRotation[2] lightPos[3] aspect rotate[16] lights[64]
void ref_shader_set_uniforms(const float* frame89, int view) {
    const float* p = frame89;
    camPos = vec3(p[0], p[1], p[2]); p += 3;
    camRotation = vec2(p[0], p[1]); p += 2;
    lightPos = vec3(p[0], p[1], p[2]); p += 3;
    aspectRatio = *p++;
    memcpy(value_ptr(rotateMatrix), p, 64); p += 16;
    for (int i = 0; i < MAX_LOCAL_LIGHTS; i++) localLights[i] = vec4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
    viewDepthField = view;
}

// the shader's castRay called directly: returns its return value; out7 = hitPos[3] hitNormal[3] stepCount
int ref_shader_cast_ray(const float start[3], const float dir[3], int dist, float* out7) {
    reset_invocation();
    int r = castRay(vec3(start[0], start[1], start[2]), vec3(dir[0], dir[1], dir[2]), dist);
    out7[0] = hitPos.x; out7[1] = hitPos.y; out7[2] = hitPos.z;
    out7[3] = hitNormal.x; out7[4] = hitNormal.y; out7[5] = hitNormal.z;
    out7[6] = stepCount;
    return r;
}

// batch version for known-answer tests: n rays, starts/dirs are n*3 floats
void ref_shader_cast_rays(int n, const float* starts, const float* dirs, const int* dists, int* ret, float* out7) {
    for (int i = 0; i < n; i++) ret[i] = ref_shader_cast_ray(starts + 3 * i, dirs + 3 * i, dists[i], out7 + 7 * i);
}

// main() for one pixel: rgba = fColor, total_steps = stepCount after main (all castRay calls of the invocation)
void ref_shader_pixel(int px, int py, int width, int height, float rgba[4], float* total_steps) {
    reset_invocation();
    set_vpos(px, py, width, height);
    shader_main();
    rgba[0] = fColor.r; rgba[1] = fColor.g; rgba[2] = fColor.b; rgba[3] = fColor.a;
    if (total_steps) *total_steps = stepCount;
}

static int g_row_stride = 1;     // bounded samples: only 8-row blocks b with b % stride == 0 are rendered
void ref_shader_set_row_stride(int s) { g_row_stride = s < 1 ? 1 : s; }

// rows [y0,y1) that pass the stride filter, dealt out 2 rows at a time to worker w of nw (nw = 1: all of them)
static void render_rows(int width, int height, int y0, int y1, float* rgba, float* total_steps, int w = 0, int nw = 1) {
    int selected = 0;
    for (int py = y0; py < y1; py++) {
        if (((py >> 3) % g_row_stride) != 0) continue;
        const int mine = ((selected++ >> 1) % nw) == w;
        if (!mine) continue;
        for (int px = 0; px < width; px++) {
            size_t p = (size_t)py * width + px;
            ref_shader_pixel(px, py, width, height, rgba + 4 * p, total_steps ? total_steps + p : nullptr);
        }
    }
}

// Rows [y0,y1) into full-frame arrays.  The shader's globals make it single-threaded per process, so
// nproc > 1 forks workers that write interleaved row blocks into shared anonymous mappings.
// Returns 0 on success.
int ref_shader_render(int width, int height, int y0, int y1, float* rgba, float* total_steps, int nproc) {
    if (nproc <= 1) { render_rows(width, height, y0, y1, rgba, total_steps); return 0; }
    size_t npix = (size_t)width * height;
    float* sh_rgba = (float*)mmap(nullptr, npix * 16, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    float* sh_steps = (float*)mmap(nullptr, npix * 4, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh_rgba == MAP_FAILED || sh_steps == MAP_FAILED) return -1;
    int failed = 0;
    for (int w = 0; w < nproc; w++) {
        pid_t pid = fork();
        if (pid == 0) {
            render_rows(width, height, y0, y1, sh_rgba, sh_steps, w, nproc);
            _exit(0);
        } else if (pid < 0) failed = 1;
    }
    int status;
    while (wait(&status) > 0) if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) failed = 1;
    for (int py = y0; py < y1; py++) {
        memcpy(rgba + 4 * (size_t)py * width, sh_rgba + 4 * (size_t)py * width, (size_t)width * 16);
        if (total_steps) memcpy(total_steps + (size_t)py * width, sh_steps + (size_t)py * width, (size_t)width * 4);
    }
    munmap(sh_rgba, npix * 16); munmap(sh_steps, npix * 4);
    return failed ? -2 : 0;
}

}  // extern "C"

#ifdef VXRT_REF_SHADER_MAIN
// Stand-alone timing / dump tool (used by bench.py --impl reference and cpu_baseline, so that the fork()s do not happen inside a
// Python process):
//   ref_shader_cli <grid.i32> <frame89+view.bin> <width> <height> <nproc> <reps> <row_stride> [out_rgba.f32]
// prints one line per repetition: "seconds".  The nproc workers are forked ONCE and stay for all repetitions (the shader's
// globals make it single-threaded per process); a repetition is the time between two process-shared barriers, i.e. rendering
// only -- no fork, no mapping, no copy of the 132 MB float frame inside the timed region.
#include <chrono>
#include <vector>
#include <algorithm>
#include <pthread.h>
struct cli_shared { pthread_barrier_t start, done; int quit; };
int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: %s grid.i32 frame.bin width height nproc reps row_stride [out.f32]\n", argv[0]); return 2; }
    FILE* fg = fopen(argv[1], "rb"); if (!fg) { perror("grid"); return 1; }
    if (fread(voxels, 1, sizeof(voxels), fg) != sizeof(voxels)) { fprintf(stderr, "short grid file\n"); return 1; }
    fclose(fg);
    float fr[90]; FILE* ff = fopen(argv[2], "rb"); if (!ff) { perror("frame"); return 1; }
    if (fread(fr, 4, 90, ff) != 90) { fprintf(stderr, "short frame file\n"); return 1; }
    fclose(ff);
    int view; memcpy(&view, &fr[89], 4);
    ref_shader_set_uniforms(fr, view);
    const int width = atoi(argv[3]), height = atoi(argv[4]), reps = atoi(argv[6]);
    int nproc = atoi(argv[5]); if (nproc < 1) nproc = 1;
    ref_shader_set_row_stride(atoi(argv[7]));
    const size_t npix = (size_t)width * height;
    float* rgba = (float*)mmap(nullptr, npix * 16, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    cli_shared* sh = (cli_shared*)mmap(nullptr, sizeof(cli_shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (rgba == MAP_FAILED || sh == MAP_FAILED) { perror("mmap"); return 1; }
    pthread_barrierattr_t ba; pthread_barrierattr_init(&ba); pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->start, &ba, (unsigned)nproc + 1); pthread_barrier_init(&sh->done, &ba, (unsigned)nproc + 1);
    sh->quit = 0;
    for (int w = 0; w < nproc; w++) {
        pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 1; }
        if (pid == 0) {
            for (;;) {
                pthread_barrier_wait(&sh->start);
                if (sh->quit) _exit(0);
                render_rows(width, height, 0, height, rgba, nullptr, w, nproc);
                pthread_barrier_wait(&sh->done);
            }
        }
    }
    std::vector<double> t;
    for (int r = 0; r < reps; r++) {
        auto a = std::chrono::steady_clock::now();
        pthread_barrier_wait(&sh->start);
        pthread_barrier_wait(&sh->done);
        auto b = std::chrono::steady_clock::now();
        t.push_back(std::chrono::duration<double>(b - a).count());
    }
    sh->quit = 1;
    pthread_barrier_wait(&sh->start);
    int status; while (wait(&status) > 0) {}
    for (double x : t) printf("%.6f\n", x);
    if (argc > 8) { FILE* fo = fopen(argv[8], "wb"); fwrite(rgba, 4, npix * 4, fo); fclose(fo); }
    return 0;
}
#endif
