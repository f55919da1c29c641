// trav.cuh -- the TRAVERSAL GRID: the device copy of the voxel grid that the rays read (north_star item 1: occupancy information
// for empty-space skipping that preserves the reference's first-hit voxel).
//
// Same index as the reference grid (x + w*y + w*h*z, render.cpp:189-196), 4 bytes per cell:
//   reference value v >= 0 (solid, 24-bit colour) or a depth-field jump (negative, != -1)   ->  v itself
//   v == -1 (empty, no jump: the band the depth field leaves next to surfaces, render.cpp:240-251)
//          ->  BAND WORD 0x80000000 | for each travel quadrant q = (stepx > 0) | (stepz > 0) << 1, 7 bits at bit 7q:
//                 K (4 bits, 0..15): every cell of the same y layer at quadrant offsets (a, b), a, b >= 0, 1 <= a + b <= K, is
//                                    inside the grid and holds -1
//                 U (3 bits, 0..7) : additionally every cell of the layer ABOVE at offsets a + b <= U - 1 is inside the grid and
//                                    holds -1; U <= K
// i.e. per -1 cell and quadrant the L1 run length to the nearest cell that is NOT a -1 cell -- an occupancy summary of the
// cell's surroundings in the unused bits of the word the ray loads anyway.  castRay (ray.cuh, TRAV) takes the steps a band word
// promises without index arithmetic, range test or load; ADJACENCY makes that exact: a step moves one cell along one axis
// whatever the float state says, ties included.  The reference-layout grid stays the master copy (uploads, downloads, edits,
// fingerprints, the depth field); every change to it is followed by trav_build_kernel over the cells whose words can change.
// A grid that holds a negative value with bit 30 clear other than via -1 (not producible by the reference: jumps are <= -2.0f)
// cannot be encoded: the kernel counts such cells and the host then renders from the reference-layout grid (TRAV = false).
// Host statement for the tests: the traversal-grid unit under oracle/ (test infrastructure; the product never links it).
#pragma once
#include <cstdint>

namespace vxrt {

constexpr int TRAV_K_CAP = 15, TRAV_U_CAP = 7;
constexpr int TRAV_REACH = TRAV_K_CAP + 1;      // a change at a cell can alter words up to this far away (in x and z; one layer below)

struct TravBox { int x0, y0, z0, nx, ny, nz; };

__device__ __forceinline__ bool trav_free(const int32_t* __restrict__ vox, int w, int h, int d, int x, int y, int z) {
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h || (unsigned)z >= (unsigned)d) return false;
    return vox[(size_t)x + (size_t)w * y + (size_t)w * h * z] == -1;
}

// min(cap, quadrant-L1 distance from (x,y,z) to the nearest cell of layer y that is not a -1 cell inside the grid), the cell
// itself included
__device__ __forceinline__ int trav_quad_dist(const int32_t* __restrict__ vox, int w, int h, int d, int x, int y, int z, int sx, int sz, int cap) {
    int best = cap;
    for (int b = 0; b < best; b++)
        for (int a = 0; a + b < best; a++)
            if (!trav_free(vox, w, h, d, x + sx * a, y, z + sz * b)) { best = a + b; break; }
    return best;
}

// the traversal word of one cell of the box (clipped to the grid); *bad += cells whose value cannot be encoded
__device__ __forceinline__ void trav_cell(const int32_t* __restrict__ vox, int32_t* __restrict__ trav, int w, int h, int d,
                                          const TravBox& b, long long t, unsigned long long* __restrict__ bad) {
    const int x = b.x0 + (int)(t % b.nx), y = b.y0 + (int)((t / b.nx) % b.ny), z = b.z0 + (int)(t / ((long long)b.nx * b.ny));
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h || (unsigned)z >= (unsigned)d) return;
    const size_t i = (size_t)x + (size_t)w * y + (size_t)w * h * z;
    const int32_t v = vox[i];
    if (v != -1) {
        if (v < 0 && !(v & 0x40000000)) atomicAdd(bad, 1ull);
        trav[i] = v;
        return;
    }
    uint32_t word = 0x80000000u;
    #pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const int sx = (q & 1) ? 1 : -1, sz = (q & 2) ? 1 : -1;
        const int K = trav_quad_dist(vox, w, h, d, x, y, z, sx, sz, TRAV_K_CAP + 1) - 1;
        int U = trav_quad_dist(vox, w, h, d, x, y + 1, z, sx, sz, TRAV_U_CAP);
        U = U < K ? U : K;
        word |= (uint32_t)(K | (U << 4)) << (7 * q);
    }
    trav[i] = (int32_t)word;
}

// one thread per cell of the box
__global__ void __launch_bounds__(256) trav_build_kernel(const int32_t* __restrict__ vox, int32_t* __restrict__ trav, int w, int h, int d,
                                                         TravBox b, unsigned long long* __restrict__ bad) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)b.nx * b.ny * b.nz) trav_cell(vox, trav, w, h, d, b, t, bad);
}

// the rebuild that follows a removeSphere whose command {cx, cy, cz, radius} lives in device memory (kernels.cuh *_cmd_kernel):
// the repaired sphere's box grown by TRAV_REACH in x and z and by one layer downward; the grid is sized for max_r by the host
__global__ void __launch_bounds__(256) trav_cmd_kernel(const int32_t* __restrict__ vox, int32_t* __restrict__ trav, int w, int h, int d,
                                                       const int* __restrict__ cmd, int max_r, unsigned long long* __restrict__ bad) {
    const int r = cmd[3];
    if (r < 0 || r > max_r) return;
    const int r2 = r + (7 >> 1);
    const TravBox b{cmd[0] - r2 - TRAV_REACH, cmd[1] - r2 - 1, cmd[2] - r2 - TRAV_REACH, 2 * (r2 + TRAV_REACH), 2 * r2 + 1, 2 * (r2 + TRAV_REACH)};
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)b.nx * b.ny * b.nz) trav_cell(vox, trav, w, h, d, b, t, bad);
}

}  // namespace vxrt
