// vxrt_render.hpp -- C++ host-side mirror of the reference's render interface (src/render.hpp:30-41 and the
// globals of src/render.hpp:15-28 / src/main.cpp:7-44) on top of the C ABI (include/vxrt.h).
//
// Same function names, argument meaning and (absent) error behaviour as the reference: all void / int, state
// lives in public members that play the role of the reference's extern globals.  A maintainer swaps
// src/render.cpp for this class; main.cpp / controls.cpp / level.cpp call sites stay as
// they are.  No GL, no GLM dependency (a 3-float struct replaces glm::vec3 at this boundary).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/vxrt.h"
#include "vxrt_controls.hpp"

namespace vxrt_host {

struct vec3 { float x, y, z; };
struct ivec3 { int x, y, z; };
struct vec2 { float x, y; };

constexpr int DEPTH_FIELD_RADIUS = 7;          // render.hpp:7
constexpr int MAX_LOCAL_LIGHTS = 16;           // render.hpp:11

class Render {
public:
    // ---- the reference's globals (render.hpp:15-28, main.cpp:7-44) ----
    int VOXELS_WIDTH = 512, VOXELS_HEIGHT = 96, VOXELS_DEPTH = 512;    // render.hpp:4-5 (runtime here)
    int screenWidth = 800, screenHeight = 600;                         // main.cpp:11-12
    long long fps = 60;                                                // main.cpp:7 (window.cpp:164 updates it)
    vec3 camPos{195, 55, 155};                                         // main.cpp:26
    vec3 camDir{0, 0, 1};
    vec2 camRotation{0, 0};
    float rotateMatrix[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // column-major, main.cpp:29
    std::vector<int32_t> voxels;                                       // host mirror; CPU collision code reads it
    vec3 startLightPos{256, 1536, 256};                                // main.cpp:33
    vec3 lightPos{256, 1536, 256};
    float aspectRatio = 800.0f / 600.0f;
    float lightRotation = -45.0f;                                      // main.cpp:36
    int viewDepthField = 0;
    float localLights[MAX_LOCAL_LIGHTS][4];
    bool keys[KEYS] = {};                                              // main.cpp:15
    int mouseX = 0, mouseY = 0;                                        // main.cpp:9-10

    Render() = default;
    ~Render();
    Render(const Render&) = delete;
    Render& operator=(const Render&) = delete;

    // ---- render.hpp:30-41 ----
    void updateGeometry();                                   // render.cpp:199-202
    void updatePartialGeometry(vec3 start, vec3 end);        // render.cpp:204-223
    void initRender();                                       // render.cpp:313-372 (level + depth field built on the device)
    int getVoxelIndex(int x, int y, int z) const;            // render.cpp:189-196
    void placeVoxel(int x, int y, int z, int voxel);         // render.cpp:256-262 (host mirror AND device)
    void destroyVoxel(int x, int y, int z);                  // render.cpp:265-271
    void fixDepthField(int x, int y, int z);                 // render.cpp:226-253 (host mirror; OOB = solid)
    void lightUpdate();                                      // render.cpp:388-402
    void updateUniforms();                                   // render.cpp:289-302
    void removeSphere(ivec3 pos, int radius);                // level.cpp:30-56, executed on the device + mirror sync
    void reshape(int width, int height);                     // render.cpp:404-411
    void placeLocalLight(float x, float y, float z, float diffuse);   // render.cpp:375-385

    // ---- what main.cpp / controls.cpp do around it ----
    void draw();                                             // glDrawArrays(GL_TRIANGLES,0,6) main.cpp:59
    void doDestroy();                                        // controls.cpp:100-110 body (the RMB branch)
    void movementUpdate();                                   // controls.cpp:22-74   } host code on the grid mirror
    void doGravity();                                        // controls.cpp:77-98   } (vxrt_controls.cpp), state
    void doMouseLook();                                      // controls.cpp:112-144 } synced with the members above
    void setMouseLook(float rotX, float rotY);               // controls.cpp:137-142: rotateMatrix = rotY*rotX, camDir
    bool saveGrid(const std::string& path);                  // device grid -> VXRTGRD1 file (include/vxrt.h)
    bool loadGrid(const std::string& path);                  // file -> device grid + host mirror
    bool writePPM(const std::string& path);                  // headless "swap buffers"
    bool readPixels(std::vector<uint8_t>& rgba);             // bottom-up RGBA8
    vxrt_stats stats();
    vxrt_ctx* context() { return ctx_; }
    const std::string& lastError() const { return err_; }

    int device = 0;

private:
    vxrt_ctx* ctx_ = nullptr;
    std::string err_;
    Player player_;
    void check(int rc);
    void toPlayer();
    void fromPlayer();
};

}  // namespace vxrt_host
