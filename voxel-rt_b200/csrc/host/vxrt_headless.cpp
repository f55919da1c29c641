// vxrt_headless.cpp -- the reference's game loop (src/main.cpp:47-75) run headless on the B200 path: no window,
// frames go to PPM files / a raw RGBA8 dump instead of glfwSwapBuffers.  Used by tests/test_gpu_host.py.
//
//   vxrt_headless [--size W H] [--frames N] [--lights] [--pitched] [--view] [--destroy] [--walk N] [--load in.vxg] [--save out.vxg] [--ppm out.ppm] [--raw out.rgba]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vxrt_render.hpp"

using namespace vxrt_host;

int main(int argc, char** argv) {
    int W = 1280, H = 720, frames = 1, walk = 0;
    bool lights = false, pitched = false, view = false, destroy = false;
    std::string ppm, raw, load, save;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--size" && i + 2 < argc) { W = atoi(argv[++i]); H = atoi(argv[++i]); }
        else if (a == "--frames" && i + 1 < argc) frames = atoi(argv[++i]);
        else if (a == "--lights") lights = true;
        else if (a == "--pitched") pitched = true;
        else if (a == "--view") view = true;
        else if (a == "--destroy") destroy = true;
        else if (a == "--walk" && i + 1 < argc) walk = atoi(argv[++i]);
        else if (a == "--load" && i + 1 < argc) load = argv[++i];
        else if (a == "--save" && i + 1 < argc) save = argv[++i];
        else if (a == "--ppm" && i + 1 < argc) ppm = argv[++i];
        else if (a == "--raw" && i + 1 < argc) raw = argv[++i];
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    Render r;
    r.screenWidth = W; r.screenHeight = H;
    const auto t0 = std::chrono::steady_clock::now();
    r.initRender();                                           // main.cpp:55
    const double init_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!load.empty() && !r.loadGrid(load)) { fprintf(stderr, "%s\n", r.lastError().c_str()); return 1; }
    if (pitched) { r.camPos = vec3{195, 60, 155}; r.setMouseLook(0.5f, 0.6f); }
    if (lights)
        for (int i = 0; i < 16; i++)                          // T key, controls.cpp:46-49, on the 4x4 pattern of SURVEY.md 8d
            r.placeLocalLight(r.camPos.x - 36 + 24 * (i % 4), 40.0f, r.camPos.z + 10 + 24 * (i / 4), 0.5f);
    r.viewDepthField = view ? 1 : 0;
    r.updateUniforms();
    for (int f = 0; f < walk; f++) {                          // main.cpp:58-66 with scripted input at a fixed 60 fps
        r.draw();
        r.keys[KEY_W] = true; r.keys[LMB] = true;             // walk forward while turning (mouse right of / below the centre)
        r.mouseX = W / 2 + W / 20; r.mouseY = H / 2 + H / 60;
        r.keys[SPACE] = f % 45 == 0;
        if (f == walk / 2) r.keys[KEY_T] = true;              // drop a light half way
        if (f == walk - 1) r.keys[RMB] = true;                // and dig on the last frame
        r.movementUpdate(); r.doMouseLook(); r.doGravity();
        if (r.keys[RMB]) { r.keys[RMB] = false; r.doDestroy(); }
        r.updateUniforms();
    }
    if (walk) printf("walk %d: cam %a %a %a dir %a %a %a rot %a %a\n", walk, r.camPos.x, r.camPos.y, r.camPos.z, r.camDir.x, r.camDir.y, r.camDir.z,
                     r.camRotation.x, r.camRotation.y);
    double ms_sum = 0;
    for (int f = 0; f < frames; f++) {                        // main.cpp:58-72 without the window
        r.draw();
        if (frames > 1) r.lightUpdate();                      // single-frame runs keep the initial sun
        if (destroy && f == 0) { r.camDir = vec3{0, -1, 0}; r.doDestroy(); r.camDir = vec3{0, 0, 1}; }
        r.updateUniforms();
        ms_sum += r.stats().ms_total;
    }
    if (destroy || walk || frames > 1) r.draw();                      // show the state after the last update
    const vxrt_stats s = r.stats();
    printf("init %.3f s; %d frame(s) %dx%d: %.3f ms/frame (device), last frame: %llu rays, %llu voxel fetches\n", init_s, frames, W, H,
           ms_sum / frames, (unsigned long long)(s.rays_primary + s.rays_global + s.rays_local), (unsigned long long)s.fetches);
    if (!save.empty() && !r.saveGrid(save)) { fprintf(stderr, "%s\n", r.lastError().c_str()); return 1; }
    if (!ppm.empty() && !r.writePPM(ppm)) return 1;
    if (!raw.empty()) {
        std::vector<uint8_t> px;
        if (!r.readPixels(px)) return 1;
        FILE* fp = fopen(raw.c_str(), "wb");
        if (!fp) return 1;
        fwrite(px.data(), 1, px.size(), fp);
        fclose(fp);
    }
    return 0;
}
