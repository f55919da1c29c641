// kernels.cuh -- the sm_100a kernels of the per-pixel path and of the grid plumbing around it.
//
//   primary_kernel   fshader.glsl:131-145 + :59-129   one thread per pixel, 8x4-pixel warps inside 32x8 tiles;
//                    misses / step-count view are finished here, hits are ballot-compacted into a hit queue
//   shade_kernel     fshader.glsl:147-187             one thread per queued hit pixel: global shadow ray, the
//                    sequential local-light loop (light list compacted into shared memory), shading, store
//   depth_kernel     render.cpp:226-253               fixDepthField for every cell of a box / lopsided sphere
//   carve_kernel     level.cpp:31-42                  removeSphere's destroy pass
//   scatter_rows     render.cpp:204-223               updatePartialGeometry's rows, one staged copy + scatter
//   assemble_kernel  multi-GPU: un-tile the all-gathered per-rank tile buffers into a raster frame
#pragma once
#include "ray.cuh"
#include "trav.cuh"

namespace vxrt {

struct FrameParams {                    // == vxrt_frame (include/vxrt.h), fshader.glsl:20-26
    float cam_pos[3];
    float cam_rotation[2];
    float light_pos[3];
    float aspect;
    float rotate[16];
    int32_t view_depth_field;
    float lights[16][4];
};

constexpr int TILE_W = 32, TILE_H = 8, TILE_PIX = TILE_W * TILE_H;

struct TileMap {
    int width, height;
    int tx, ty, ntiles;                 // tiles per row / column / total
    int rank, world, nlocal;            // this context renders tiles t = j*world + rank, j in [0,nlocal)
    int tile_base;                      // first local tile of this launch (frames can be rendered in bands of tile rows)
    const int2* tile_xy;                // optional, per local tile: pixel origin (x0, y0) of its global tile, y0 < 0 for padding tiles -- the host's
                                        // statement of tile_of(), so that the render kernels need no integer division to find their pixels
    int rows;                           // 1: whole TILE ROWS are the interleaved unit (vxrt_set_partition): tile row r belongs to rank r % world,
                                        // local tile j = (local row j / tx, column j % tx); a rank's pixels are then contiguous 8-row strips
                                        // of the raster frame, which one strided DMA moves to a host frame
};
// Tile partition: the tiles are dealt in groups of `world` consecutive tiles (row-major), group j = tiles j*world .. j*world + world - 1,
// one to each rank -- rotated by the tile row the group starts in, so that a rank's tiles do not line up in columns when the
// number of tiles per row is a multiple of world (3840 / 32 = 120 tiles, 8 ranks: a rank would own every 8th 32-pixel column, and
// vertical features of the scene -- tree trunks -- would land on one rank: measured 10 % above the mean for the slowest of 8).
__host__ __device__ __forceinline__ int tile_rotation(const TileMap& m, int j) { return ((j * m.world) / m.tx) % m.world; }
// global tile of local tile j (>= ntiles: padding)
__host__ __device__ __forceinline__ int tile_of(const TileMap& m, int j) {
    return m.rows ? ((j / m.tx) * m.world + m.rank) * m.tx + (j % m.tx) : j * m.world + (m.rank + tile_rotation(m, j)) % m.world;
}
// rank that owns global tile t (tile partition)
__host__ __device__ __forceinline__ int tile_owner(const TileMap& m, int t) {
    const int j = t / m.world;
    return (t % m.world - tile_rotation(m, j) + m.world) % m.world;
}

struct Counters {
    unsigned int hit_count;
    unsigned int global_traced;         // counted variants: global-light rays actually traced (== hit_count unless unlit rays are skipped)
    unsigned long long rays_local, fetches_primary, fetches_shadow;
    unsigned long long rays_dark;       // shadow / light rays whose surface faces away from the light (cannot change the pixel)
};

struct Outputs {
    uint32_t* rgba8;                    // raster [height][width] (world==1 or peer-memory target); else tile-compact [nlocal][8][32]
    int raster;                         // 1: rgba8 is a raster frame
    int skip_dark;                      // do not trace rays whose outcome cannot change the pixel (N.L <= 0): production frames, and the
                                        // counted variants under vxrt_set_stats(2) (count what the production kernels execute)
    float4* hitq;                       // per local tile, 256 slots: hitPos.xyz, w = colour(24) | normal(4)<<24 of the tile's hit pixels, compacted
    uint32_t* hitpix;                   // raster pixel id of each entry
    uint32_t* tile_hits;                // per local tile: number of entries
    const uint32_t* shade_order;        // optional: launch order of the shade units (a unit = one shade block's share of a tile)
    uint32_t* shade_cost;               // optional: per shade unit, SM cycles its block took this frame
    int shade_unit_base;                // first shade unit of this launch
    Counters* counters;
    const uint32_t* tile_order;         // optional: launch order of the local tiles (longest first, from the previous frame)
    uint32_t* tile_cost;                // optional: per local tile, SM cycles its block took this frame
    // overlap of the two passes (programmatic dependent launch): the shade pass starts while the primary pass's last blocks
    // still run; a shade block waits for ITS tile's flag instead of the kernel boundary
    uint32_t* tile_ready;               // per local tile: frame_seq of the last primary pass that finished it
    uint32_t frame_seq;
    int overlap;
    int* overlap_err;                   // set when a shade block gave up waiting (bounded spin)
    // peer-memory frames (programmatic dependent launch): the first render kernel of a frame is launched while p2p_begin_kernel still
    // waits for the owner (pdl_wait: griddepcontrol.wait before this kernel's first global write); the last one lets the kernel that
    // publishes the rank's completion flag become resident early (pdl_trigger)
    int pdl_wait, pdl_trigger;
    int wide_blocks;                    // frame_kernel: its first wide_blocks blocks render the wide_blocks / 2 heaviest tiles, two blocks each, with
                                        // two threads per hit pixel (see shade_wide_*); 0: every block renders one tile
    int32_t* dbg_hit;                   // optional (VXRT_FLAG_DEBUG_OUTPUTS), raster layout
    uint16_t* dbg_steps;
    uint32_t* dbg_occl;
    uint32_t* dbg_cast;
};

// the pixel store.  VXRT_EXP_STREAMING_STORES (variant library, voxel-rt_b200/build.py): st.global.cs -- the 33 MB of a 4K frame are
// written once and read by the copy engine, they need not displace the grid in L2
__device__ __forceinline__ void store_pixel(uint32_t* p, uint32_t v) {
#ifdef VXRT_EXP_STREAMING_STORES
    __stcs(p, v);
#else
    *p = v;
#endif
}

__device__ __forceinline__ uint32_t out_index_of(const TileMap& m, int raster, int px, int py) {
    if (raster) return (uint32_t)(py * m.width + px);
    if (m.rows) return (uint32_t)((((py / TILE_H) / m.world) * TILE_H + (py % TILE_H)) * m.width + px);   // this rank's strips, raster inside
    const int t = (py / TILE_H) * m.tx + (px / TILE_W);
    const int local = t / m.world;
    return (uint32_t)(local * TILE_PIX + (py % TILE_H) * TILE_W + (px % TILE_W));
}

// ------------------------------------------------------------------------------------------------
// COUNT: maintain the per-ray iteration counter (needed by the step-count view, the debug planes and the fetch
// statistics); the production frame path runs without it.
// TRAV: g.vox is the traversal grid (trav.cuh)

// fshader.glsl:131-145 for the pixel of this thread: ray set-up, primary castRay, misses / the step-count view finished here.
// Returns true when the pixel hit a voxel (its lighting follows, fshader.glsl:147-187).
template <bool COUNT, class Grid, bool TRAV, bool EARLY = false>
__device__ __forceinline__ bool primary_pixel(const Grid& g, const FrameParams& f, const TileMap& m, const Outputs& o, int local_tile,
                                              int warp, int lane, RayHit& r, uint32_t& pid) {
    const int lx = (warp & 3) * 8 + (lane & 7), ly = (warp >> 2) * 4 + (lane >> 3);
    int x0, y0;
    bool tile_ok;
    if (m.tile_xy) {                                                 // the partition, tabulated by the host (vxrt.cu alloc_frame_buffers)
        const int2 o2 = __ldg(m.tile_xy + local_tile);
        x0 = o2.x; y0 = o2.y; tile_ok = y0 >= 0;
    } else {
        const int t = tile_of(m, local_tile);                        // global tile
        x0 = (t % m.tx) * TILE_W; y0 = (t / m.tx) * TILE_H; tile_ok = t < m.ntiles;
    }
    const int px = x0 + lx, py = y0 + ly;
    const bool valid = tile_ok && (px < m.width) && (py < m.height);
    bool hit = false;
    r.steps = 0; r.idx = -1;
    pid = (uint32_t)(py * m.width + px);
    if (valid) {
        // vshader.glsl:6-9 + quad render.cpp:36-44: vPos = NDC of the pixel centre
        const float vx = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)px, 0.5f), (float)m.width), 2.0f), 1.0f);
        const float vy = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)py, 0.5f), (float)m.height), 2.0f), 1.0f);
        float dxn = __fmul_rn(vx, f.aspect), dyn = vy, dzn = 1.0f;   // :136
        normalize3(dxn, dyn, dzn);
        // :137  mat4 * vec4(dir,0) = (m[0]*v0 + m[1]*v1) + (m[2]*v2 + m[3]*0)   (GLM type_mat4x4.inl:561-572)
        const float* M = f.rotate;
        const float rx = __fadd_rn(__fadd_rn(__fmul_rn(M[0], dxn), __fmul_rn(M[4], dyn)), __fadd_rn(__fmul_rn(M[8], dzn), __fmul_rn(M[12], 0.0f)));
        const float ry = __fadd_rn(__fadd_rn(__fmul_rn(M[1], dxn), __fmul_rn(M[5], dyn)), __fadd_rn(__fmul_rn(M[9], dzn), __fmul_rn(M[13], 0.0f)));
        const float rz = __fadd_rn(__fadd_rn(__fmul_rn(M[2], dxn), __fmul_rn(M[6], dyn)), __fadd_rn(__fmul_rn(M[10], dzn), __fmul_rn(M[14], 0.0f)));
        r = cast_ray<COUNT, false, true, Grid, TRAV, EARLY>(g, f.cam_pos[0], f.cam_pos[1], f.cam_pos[2], rx, ry, rz, VXRT_RENDER_DIST);   // :139
        if (o.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");   // nothing is written before the kernel ahead has completed
        if (f.view_depth_field == 1) {                               // :143-145
            const float grey = __fdiv_rn((float)r.steps, 100.0f);
            store_pixel(o.rgba8 + out_index_of(m, o.raster, px, py), pack_rgba8(grey, grey, grey, 1.0f));
        } else if (r.idx >= 0) {
            hit = true;                                              // lighting follows
        } else {
            store_pixel(o.rgba8 + out_index_of(m, o.raster, px, py), pack_rgba8(0.6f, 0.7f, 0.8f, 1.0f));   // :133
        }
        if (o.dbg_hit) {
            o.dbg_hit[pid] = r.idx;
            o.dbg_steps[pid] = (uint16_t)(r.steps > 65535 ? 65535 : r.steps);
            if (!hit) { o.dbg_occl[pid] = 0u; o.dbg_cast[pid] = 0u; }
        }
    }
    return hit;
}
__device__ __forceinline__ float4 hit_record(const RayHit& r) {     // hitPos.xyz, w = colour(24) | normal(4) << 24
    return make_float4(r.hx, r.hy, r.hz, __uint_as_float(((uint32_t)r.voxel & 0x00FFFFFFu) | ((uint32_t)r.normal << 24)));
}

template <bool COUNT, class Grid, bool TRAV>
__global__ void __launch_bounds__(256, 6) primary_kernel(Grid g, const __grid_constant__ FrameParams f,
                                                      TileMap m, Outputs o) {
    __shared__ unsigned int s_warp_hits[8];
    __shared__ unsigned long long s_fetches;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_fetches = 0ull;
    // overlap: the shade pass (launched with programmatic stream serialization) may be scheduled as soon as every block of this
    // grid has started, i.e. into the SM capacity this pass's tail leaves idle
    if (o.overlap || o.pdl_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long t_start = clock64();
    // blocks are handed out in launch order: with a tile order from the previous frame the slowest tiles start first,
    // which shortens the kernel's tail (it matters once a GPU renders only 1/4 or 1/8 of the frame)
    const int local_tile = o.tile_order ? (int)o.tile_order[blockIdx.x] : (int)blockIdx.x + m.tile_base;
    RayHit r;
    uint32_t pid;
    const bool hit = primary_pixel<COUNT, Grid, TRAV, true>(g, f, m, o, local_tile, warp, lane, r, pid);   // (EARLY: ray.cuh cast_ray)
    if (o.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");       // (threads without a pixel have not waited yet)
    // ---- hit compaction into the tile's 256 slots: warp ballot -> block prefix (no global ordering: the shade pass
    //      schedules tiles by its own cost feedback) ---------------------------------------------------------------
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    const unsigned wfetch = __reduce_add_sync(0xffffffffu, (unsigned)r.steps);
    if (lane == 0) s_warp_hits[warp] = __popc(ballot);
    __syncthreads();
    if (lane == 0 && wfetch) atomicAdd(&s_fetches, (unsigned long long)wfetch);
    if (tid == 0) {
        unsigned total = 0;
        #pragma unroll
        for (int w = 0; w < 8; w++) { const unsigned c = s_warp_hits[w]; s_warp_hits[w] = total; total += c; }
        o.tile_hits[local_tile] = total;
        if (total) atomicAdd(&o.counters->hit_count, total);          // statistics only
    }
    __syncthreads();
    if (hit) {
        const unsigned pos = (unsigned)local_tile * TILE_PIX + s_warp_hits[warp] + __popc(ballot & ((1u << lane) - 1u));
        o.hitq[pos] = hit_record(r);
        o.hitpix[pos] = pid;
    }
    if (tid == 0 && s_fetches) atomicAdd(&o.counters->fetches_primary, s_fetches);
    if (tid == 0 && o.tile_cost) o.tile_cost[local_tile] = (uint32_t)min((long long)0xffffffffll, clock64() - t_start);
    if (o.overlap) {
        // publish the tile: every thread's hit slots (and the tile's hit count) are written; the barrier orders them before
        // thread 0's release store, which the tile's shade blocks acquire
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(o.tile_ready + local_tile), "r"(o.frame_seq) : "memory");
    }
}

// cost[n] -> order[n], most expensive first: one-block counting sort on a 128-bucket logarithmic key.  Per-warp
// private histograms and match_any ranking keep it free of contended atomics (costs cluster in a few buckets).
__global__ void __launch_bounds__(1024) tile_order_kernel(const uint32_t* __restrict__ cost, uint32_t* __restrict__ order, int n) {
    __shared__ unsigned int s_hist[32][128];                        // [warp][bucket] counts, then start positions
    __shared__ unsigned int s_total[128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32 * 128; i += 1024) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    auto key = [](uint32_t c) -> int {                               // 127 = most expensive
        if (c < 4u) return 0;
        const int e = 31 - __clz(c);
        const int k = e * 4 + (int)((c >> (e - 2)) & 3u);
        return k > 127 ? 127 : k;
    };
    const int per_warp = (n + 31) / 32, w0 = min(n, warp * per_warp), w1 = min(n, w0 + per_warp);
    for (int base = w0; base < w1; base += 32) {
        const int i = base + lane;
        const bool in = i < w1;
        const int k = in ? key(cost[i]) : 128 + lane;                // out-of-range lanes get unique keys
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (in && lane == __ffs(peers) - 1) s_hist[warp][k] += __popc(peers);
        __syncwarp();                                                // the next batch's leader of the same key may be another lane
    }
    __syncthreads();
    if (tid < 128) {                                                 // per bucket: per-warp offsets inside the bucket, bucket total
        unsigned acc = 0;
        for (int w = 0; w < 32; w++) { const unsigned c = s_hist[w][tid]; s_hist[w][tid] = acc; acc += c; }
        s_total[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {                                                  // bucket start positions, most expensive bucket first
        unsigned acc = 0;
        for (int b = 127; b >= 0; b--) { const unsigned c = s_total[b]; s_total[b] = acc; acc += c; }
    }
    __syncthreads();
    for (int base = w0; base < w1; base += 32) {
        const int i = base + lane;
        const bool in = i < w1;
        const int k = in ? key(cost[i]) : 128 + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        if (in) {
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            order[s_total[k] + s_hist[warp][k] + rank] = (uint32_t)i;
        }
        __syncwarp();
        if (in && lane == __ffs(peers) - 1) s_hist[warp][k] += __popc(peers);
        __syncwarp();
    }
}

// TRAV: g.vox is the traversal grid (trav.cuh): shadow / light rays take the runs its band words promise

struct LightList {                      // active lights of the frame, compacted (slot order preserved), in shared memory
    float4 light[16];
    int slot[16];
    int nactive;
};
// first warp of the block: fshader.glsl:167, a slot is in the scene iff x,y,z >= 0 (a barrier must follow)
__device__ __forceinline__ void compact_lights(const FrameParams& f, LightList& L, int tid, int lane) {
    if (tid < 32) {
        const bool act = (tid < 16) && f.lights[tid & 15][0] >= 0.0f && f.lights[tid & 15][1] >= 0.0f && f.lights[tid & 15][2] >= 0.0f;
        const unsigned b = __ballot_sync(0xffffffffu, act);
        if (act) {
            const int k = __popc(b & ((1u << lane) - 1u));
            L.light[k] = make_float4(f.lights[tid][0], f.lights[tid][1], f.lights[tid][2], f.lights[tid][3]);
            L.slot[k] = tid;
        }
        if (tid == 0) L.nactive = __popc(b);
    }
}
struct ShadeCounts { unsigned fetches, nlocal, ndark, nglobal; };

// fshader.glsl:147-187 for one hit pixel: global-light shadow ray, the sequential local-light loop, colour, store
template <bool COUNT, class Grid, bool TRAV>
__device__ __forceinline__ void shade_pixel(const Grid& g, const FrameParams& f, const TileMap& m, const Outputs& o, const LightList& LL,
                                            const float4 rec, const uint32_t pid, ShadeCounts& n) {
    const uint32_t packed = __float_as_uint(rec.w);
    float nx, ny, nz;
    unpack_normal((int)(packed >> 24), nx, ny, nz);
    const float hx = rec.x, hy = rec.y, hz = rec.z;
    // Rays toward a light the surface faces away from (N.L <= 0) add exactly +-0 to the multiplier whether they
    // are occluded or not (fshader.glsl:155,177: "* max(0, dot(N, L))"), so production frames do not trace them; the
    // sign test uses the unnormalised direction (normalising multiplies by a positive number).  Counted variants
    // (statistics / debug planes) trace every ray the reference casts and report how many were of this kind, unless
    // the host asks them to count what the production kernels execute (skip_dark set: rays_dark = rays skipped).
    const bool skip_dark = o.skip_dark != 0;
    // :147
    float lx = __fsub_rn(f.light_pos[0], hx), ly = __fsub_rn(f.light_pos[1], hy), lz = __fsub_rn(f.light_pos[2], hz);
    const bool g_lit = dot3(nx, ny, nz, lx, ly, lz) > 0.0f;
    float multiplier = VXRT_AMBIENT;                             // :149
    uint32_t occl = 0u, cast = 1u;
    if (COUNT && !g_lit) n.ndark++;
    if (g_lit || !skip_dark) {   // :154 global-light shadow ray
        if (COUNT) n.nglobal++;
        normalize3(lx, ly, lz);
        const RayHit s = cast_ray<COUNT, true, true, Grid, TRAV>(g, __fadd_rn(hx, __fmul_rn(lx, 0.001f)), __fadd_rn(hy, __fmul_rn(ly, 0.001f)),
                                  __fadd_rn(hz, __fmul_rn(lz, 0.001f)), lx, ly, lz, VXRT_RENDER_DIST);
        n.fetches += (unsigned)s.steps;
        if (s.idx == -1) multiplier = __fadd_rn(multiplier, __fmul_rn(VXRT_DIFFUSE, max0(dot3(nx, ny, nz, lx, ly, lz))));   // :155
        else occl |= 1u;
    }
    // :159-181.  The reference walks slots 0..15 and tests the overbright clamp at the TOP of every
    // iteration (active slot or not); walking only the active slots is equivalent when the clamp is also
    // applied for the inactive iterations that follow the last active slot.
    const int nact = LL.nactive;
    int last_slot = -1;
    bool broke = false;
    for (int k = 0; k < nact; k++) {
        if (multiplier >= VXRT_MAX_OVERBRIGHT) { multiplier = VXRT_MAX_OVERBRIGHT; broke = true; break; }   // :161-164
        const float4 L = LL.light[k];
        const int slot = LL.slot[k];
        last_slot = slot;
        float tx = __fsub_rn(L.x, hx), ty = __fsub_rn(L.y, hy), tz = __fsub_rn(L.z, hz);
        // (a non-finite weight would turn the +-0 into NaN: such a light is always traced)
        const bool lit = dot3(nx, ny, nz, tx, ty, tz) > 0.0f || !(fabsf(L.w) <= 3.0e38f);
        if (!COUNT && skip_dark && !lit) continue;
        const float lld = __fsqrt_rn(dot3(tx, ty, tz, tx, ty, tz));                      // :168
        if (lld <= (float)VXRT_LOCAL_LIGHT_DIST) {                                       // :171
            if (COUNT && !lit) { n.ndark++; if (skip_dark) continue; }
            normalize3_with_length(tx, ty, tz, lld);                                    // :173 (same dot, same sqrt as :168)
            cast |= 2u << slot; n.nlocal++;
            const RayHit s = cast_ray<COUNT, true, false, Grid, TRAV>(g, __fadd_rn(hx, __fmul_rn(tx, 0.001f)), __fadd_rn(hy, __fmul_rn(ty, 0.001f)),
                                      __fadd_rn(hz, __fmul_rn(tz, 0.001f)), tx, ty, tz, f2i(__fadd_rn(lld, 1.0f)));   // :175
            n.fetches += (unsigned)s.steps;
            if (s.idx == -1) {                                                          // :177
                // (64 - d) / 64: dividing by a power of two is an exact scaling, identical to the IEEE quotient
                const float fall = __fmul_rn(__fsub_rn((float)VXRT_LOCAL_LIGHT_DIST, lld), 1.0f / (float)VXRT_LOCAL_LIGHT_DIST);
                multiplier = __fadd_rn(multiplier, __fmul_rn(__fmul_rn(L.w, max0(dot3(nx, ny, nz, tx, ty, tz))), fall));
            } else occl |= 2u << slot;
        }
    }
    if (!broke && last_slot < 15 && multiplier >= VXRT_MAX_OVERBRIGHT) multiplier = VXRT_MAX_OVERBRIGHT;
    // :184-187
    const float cr = __fmul_rn(__fdiv_rn((float)((packed >> 16) & 255u), 255.0f), multiplier);
    const float cg = __fmul_rn(__fdiv_rn((float)((packed >> 8) & 255u), 255.0f), multiplier);
    const float cb = __fmul_rn(__fdiv_rn((float)(packed & 255u), 255.0f), multiplier);
    const int px = (int)(pid % (uint32_t)m.width), py = (int)(pid / (uint32_t)m.width);
    store_pixel(o.rgba8 + out_index_of(m, o.raster, px, py), pack_rgba8(cr, cg, cb, 1.0f));
    if (o.dbg_occl) { o.dbg_occl[pid] = occl; o.dbg_cast[pid] = cast; }
}
// ---- wide blocks (frame_kernel; production variants only) --------------------------------------------------------------------
// A small share of a frame lasts as long as its slowest block, and the slowest block is one tile whose pixels each walk all
// 16 lights with long unobstructed rays: 1 + 16 rays in sequence per thread.  What a light ADDS to the multiplier
// (fshader.glsl:166-179) does not depend on the multiplier, only whether the loop still runs does (:161-164), so the lights of a
// pixel can be evaluated by two threads -- the second one speculatively, without the early-out -- and combined in slot order
// afterwards: same additions, same order, same clamp tests.  The K heaviest tiles (previous frame's block times) are rendered
// that way by two blocks each: a block takes half a tile (128 pixels); threads 0-127 trace the primary rays, the global shadow
// ray and the first half of the active lights, threads 128-255 the second half of the lights of the same hit pixels.
#ifdef VXRT_EXP_WIDE_NOINLINE           /* variant library: the wide path as functions of their own (register allocation apart from the kernel's) */
#define VXRT_WIDE_FN __device__ __noinline__
#else
#define VXRT_WIDE_FN __device__ __forceinline__
#endif
struct WideShared {
    float c[8][128];                    // [light of the second half][hit]: what it adds
    unsigned char has[128];             // bit j: light j of the second half adds c[j]
};
// fshader.glsl:166-179 for one light and one hit pixel: true when the light adds c to the multiplier
template <class Grid, bool TRAV>
__device__ __forceinline__ bool light_term(const Grid& g, const float4 L, float hx, float hy, float hz, float nx, float ny, float nz,
                                           bool skip_dark, float& c) {
    float tx = __fsub_rn(L.x, hx), ty = __fsub_rn(L.y, hy), tz = __fsub_rn(L.z, hz);
    const bool lit = dot3(nx, ny, nz, tx, ty, tz) > 0.0f || !(fabsf(L.w) <= 3.0e38f);
    if (skip_dark && !lit) return false;
    const float lld = __fsqrt_rn(dot3(tx, ty, tz, tx, ty, tz));                          // :168
    if (!(lld <= (float)VXRT_LOCAL_LIGHT_DIST)) return false;                            // :171
    normalize3_with_length(tx, ty, tz, lld);                                            // :173
    const RayHit s = cast_ray<false, true, false, Grid, TRAV>(g, __fadd_rn(hx, __fmul_rn(tx, 0.001f)), __fadd_rn(hy, __fmul_rn(ty, 0.001f)),
                              __fadd_rn(hz, __fmul_rn(tz, 0.001f)), tx, ty, tz, f2i(__fadd_rn(lld, 1.0f)));   // :175
    if (s.idx != -1) return false;                                                      // :177
    const float fall = __fmul_rn(__fsub_rn((float)VXRT_LOCAL_LIGHT_DIST, lld), 1.0f / (float)VXRT_LOCAL_LIGHT_DIST);
    c = __fmul_rn(__fmul_rn(L.w, max0(dot3(nx, ny, nz, tx, ty, tz))), fall);
    return true;
}
struct WideState { float multiplier; int last_slot; bool broke; };
// before the barrier.  role 0: global shadow ray + lights [0, n0) in sequence (returns the loop's state); role 1: lights [n0, nactive)
// without the early-out, into shared memory
template <class Grid, bool TRAV>
VXRT_WIDE_FN WideState shade_wide_first(const Grid& g, const FrameParams& f, const Outputs& o, const LightList& LL, const float4 rec,
                                                      int role, int idx, WideShared& W) {
    const uint32_t packed = __float_as_uint(rec.w);
    float nx, ny, nz;
    unpack_normal((int)(packed >> 24), nx, ny, nz);
    const float hx = rec.x, hy = rec.y, hz = rec.z;
    const bool skip_dark = o.skip_dark != 0;
    const int nact = LL.nactive, n0 = nact >> 1;
    WideState st = {VXRT_AMBIENT, -1, false};
    if (role == 0) {
        float lx = __fsub_rn(f.light_pos[0], hx), ly = __fsub_rn(f.light_pos[1], hy), lz = __fsub_rn(f.light_pos[2], hz);   // :147
        if (dot3(nx, ny, nz, lx, ly, lz) > 0.0f || !skip_dark) {                        // :154
            normalize3(lx, ly, lz);
            const RayHit s = cast_ray<false, true, true, Grid, TRAV>(g, __fadd_rn(hx, __fmul_rn(lx, 0.001f)), __fadd_rn(hy, __fmul_rn(ly, 0.001f)),
                                      __fadd_rn(hz, __fmul_rn(lz, 0.001f)), lx, ly, lz, VXRT_RENDER_DIST);
            if (s.idx == -1) st.multiplier = __fadd_rn(st.multiplier, __fmul_rn(VXRT_DIFFUSE, max0(dot3(nx, ny, nz, lx, ly, lz))));   // :155
        }
        for (int k = 0; k < n0; k++) {
            if (st.multiplier >= VXRT_MAX_OVERBRIGHT) { st.multiplier = VXRT_MAX_OVERBRIGHT; st.broke = true; break; }   // :161-164
            st.last_slot = LL.slot[k];
            float c;
            if (light_term<Grid, TRAV>(g, LL.light[k], hx, hy, hz, nx, ny, nz, skip_dark, c)) st.multiplier = __fadd_rn(st.multiplier, c);
        }
    } else {
        unsigned has = 0u;
        for (int k = n0; k < nact; k++) {
            float c = 0.0f;
            if (light_term<Grid, TRAV>(g, LL.light[k], hx, hy, hz, nx, ny, nz, skip_dark, c)) has |= 1u << (k - n0);
            W.c[k - n0][idx] = c;
        }
        W.has[idx] = (unsigned char)has;
    }
    return st;
}
// after the barrier, role 0: the second half's terms in slot order, with the clamp test in front of each; colour; store
VXRT_WIDE_FN void shade_wide_second(const TileMap& m, const Outputs& o, const LightList& LL, const float4 rec, const uint32_t pid,
                                                  WideState st, int idx, const WideShared& W) {
    const uint32_t packed = __float_as_uint(rec.w);
    const int nact = LL.nactive, n0 = nact >> 1;
    if (!st.broke) {
        const unsigned has = W.has[idx];
        for (int k = n0; k < nact; k++) {
            if (st.multiplier >= VXRT_MAX_OVERBRIGHT) { st.multiplier = VXRT_MAX_OVERBRIGHT; st.broke = true; break; }   // :161-164
            st.last_slot = LL.slot[k];
            if ((has >> (k - n0)) & 1u) st.multiplier = __fadd_rn(st.multiplier, W.c[k - n0][idx]);
        }
    }
    if (!st.broke && st.last_slot < 15 && st.multiplier >= VXRT_MAX_OVERBRIGHT) st.multiplier = VXRT_MAX_OVERBRIGHT;
    const float cr = __fmul_rn(__fdiv_rn((float)((packed >> 16) & 255u), 255.0f), st.multiplier);      // :184-187
    const float cg = __fmul_rn(__fdiv_rn((float)((packed >> 8) & 255u), 255.0f), st.multiplier);
    const float cb = __fmul_rn(__fdiv_rn((float)(packed & 255u), 255.0f), st.multiplier);
    const int px = (int)(pid % (uint32_t)m.width), py = (int)(pid / (uint32_t)m.width);
    store_pixel(o.rgba8 + out_index_of(m, o.raster, px, py), pack_rgba8(cr, cg, cb, 1.0f));
}

// block-level sums of the shade counters into the frame's counters (s_fetches / s_local: zeroed shared scratch; ends with a barrier)
template <bool COUNT>
__device__ __forceinline__ void shade_counts_to_global(const Outputs& o, const ShadeCounts& n, unsigned long long& s_fetches,
                                                       unsigned long long& s_local, int tid, int lane) {
    const unsigned wf = __reduce_add_sync(0xffffffffu, n.fetches), wl = __reduce_add_sync(0xffffffffu, n.nlocal);
    if (COUNT) {
        const unsigned wd = __reduce_add_sync(0xffffffffu, n.ndark), wg = __reduce_add_sync(0xffffffffu, n.nglobal);
        if (lane == 0 && wd) atomicAdd(&o.counters->rays_dark, (unsigned long long)wd);
        if (lane == 0 && wg) atomicAdd(&o.counters->global_traced, wg);
    }
    if (lane == 0 && wf) { atomicAdd(&s_fetches, (unsigned long long)wf); atomicAdd(&s_local, (unsigned long long)wl); }
    __syncthreads();
    if (tid == 0 && s_fetches) { atomicAdd(&o.counters->fetches_shadow, s_fetches); atomicAdd(&o.counters->rays_local, s_local); }
}

template <bool COUNT, class Grid, bool TRAV>
__global__ void __launch_bounds__(256, 5) shade_kernel(Grid g, const __grid_constant__ FrameParams f,
                                                    TileMap m, Outputs o) {
    __shared__ LightList s_lights;
    __shared__ unsigned long long s_fetches, s_local;
    const int tid = threadIdx.x, lane = tid & 31;
    if (o.pdl_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // a shade unit = blockDim.x consecutive slots of one tile; units are launched slowest-first (previous frame's times)
    const long long t_start = clock64();
    const int unit = o.shade_order ? (int)o.shade_order[blockIdx.x] : (int)blockIdx.x + o.shade_unit_base;
    const int units_per_tile = TILE_PIX / (int)blockDim.x;
    const int tile = unit / units_per_tile, slot0 = (unit % units_per_tile) * (int)blockDim.x, slot = slot0 + tid;
    if (o.overlap) {
        // the primary pass may still be running: wait until it has published this tile (bounded: ~2 s, then give up loudly)
        if (tid == 0) {
            unsigned v;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(o.tile_ready + tile) : "memory");
                if (v == o.frame_seq) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 2000000000ull) { *o.overlap_err = 1; break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    const unsigned count = o.tile_hits[tile];
    if ((unsigned)slot0 >= count) {                                  // sky tile / empty part of a tile: nothing to shade
        if (tid == 0 && o.shade_cost) o.shade_cost[unit] = 0u;
        return;
    }
    compact_lights(f, s_lights, tid, lane);
    if (tid == 0) { s_fetches = 0ull; s_local = 0ull; }
    __syncthreads();
    const unsigned i = (unsigned)tile * TILE_PIX + (unsigned)slot;
    ShadeCounts n = {0u, 0u, 0u, 0u};
    if ((unsigned)slot < count) shade_pixel<COUNT, Grid, TRAV>(g, f, m, o, s_lights, o.hitq[i], o.hitpix[i], n);
    shade_counts_to_global<COUNT>(o, n, s_fetches, s_local, tid, lane);
    if (tid == 0 && o.shade_cost) o.shade_cost[unit] = (uint32_t)min((long long)0xffffffffll, clock64() - t_start);
}

// The whole of fshader.glsl's main() for one tile in ONE block: primary rays, hit compaction in shared memory, then the block
// shades its own hits.  Used when a context renders a small share of the frame (vxrt_set_fusion, auto): there the frame time is set
// by the serial chains -- the longest primary ray (~50 us), then the longest pixel's sequential light loop (~100 us) -- and a
// kernel boundary between the passes makes the two add; fused, a tile's lighting starts the moment its own primary rays are done,
// and blocks launch slowest-first by their combined time.  No hit slots in global memory, one launch.
template <bool COUNT, class Grid, bool TRAV>
__global__ void __launch_bounds__(256, 5) frame_kernel(Grid g, const __grid_constant__ FrameParams f,
                                                    TileMap m, Outputs o) {
    __shared__ LightList s_lights;
    __shared__ float4 s_hit[TILE_PIX];
    __shared__ uint32_t s_pix[TILE_PIX];
    __shared__ unsigned int s_warp_hits[8];
    __shared__ unsigned int s_total;
    __shared__ unsigned long long s_fetches, s_local, s_fetches_primary;
    __shared__ WideShared s_wide;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (o.pdl_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long t_start = clock64();
    // wide blocks (production variants): the first o.wide_blocks blocks take the heaviest tiles, two blocks per tile (shade_wide_*)
    const bool wide = !COUNT && (int)blockIdx.x < o.wide_blocks;
    const int pos = wide ? (int)(blockIdx.x >> 1) : (int)blockIdx.x - (COUNT ? 0 : (o.wide_blocks >> 1));
    const int local_tile = o.tile_order ? (int)o.tile_order[pos] : pos + m.tile_base;
    compact_lights(f, s_lights, tid, lane);
    if (tid == 0) { s_fetches = 0ull; s_local = 0ull; s_fetches_primary = 0ull; }
    RayHit r;
    uint32_t pid = 0u;
    bool hit = false;
    if (!wide) hit = primary_pixel<COUNT, Grid, TRAV>(g, f, m, o, local_tile, warp, lane, r, pid);
    else if (warp < 4) hit = primary_pixel<COUNT, Grid, TRAV>(g, f, m, o, local_tile, (int)(blockIdx.x & 1u) * 4 + warp, lane, r, pid);
    else { r.steps = 0; r.idx = -1; }                                // (the second half's threads join for the lights)
    if (o.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");       // (threads without a pixel have not waited yet)
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    const unsigned wfetch = __reduce_add_sync(0xffffffffu, (unsigned)r.steps);
    if (lane == 0) s_warp_hits[warp] = __popc(ballot);
    __syncthreads();
    if (lane == 0 && wfetch) atomicAdd(&s_fetches_primary, (unsigned long long)wfetch);
    if (tid == 0) {
        unsigned total = 0;
        #pragma unroll
        for (int w = 0; w < 8; w++) { const unsigned c = s_warp_hits[w]; s_warp_hits[w] = total; total += c; }
        s_total = total;
        o.tile_hits[local_tile] = total;
        if (total) atomicAdd(&o.counters->hit_count, total);          // statistics only
    }
    __syncthreads();
    if (hit) {
        const unsigned pos = s_warp_hits[warp] + __popc(ballot & ((1u << lane) - 1u));
        s_hit[pos] = hit_record(r);
        s_pix[pos] = pid;
    }
    __syncthreads();
    if (tid == 0 && s_fetches_primary) atomicAdd(&o.counters->fetches_primary, s_fetches_primary);
    ShadeCounts n = {0u, 0u, 0u, 0u};
    if (wide) {
        if (!COUNT) {                                                // (never taken by the counted variants: wide is false there)
            const int idx = tid & 127, role = tid >> 7;
            const bool work = f.view_depth_field != 1 && (unsigned)idx < s_total;
            WideState st = {VXRT_AMBIENT, -1, false};
            if (work) st = shade_wide_first<Grid, TRAV>(g, f, o, s_lights, s_hit[idx], role, idx, s_wide);
            __syncthreads();
            if (work && role == 0) shade_wide_second(m, o, s_lights, s_hit[idx], s_pix[idx], st, idx, s_wide);
        }
    } else if (f.view_depth_field != 1 && (unsigned)tid < s_total) shade_pixel<COUNT, Grid, TRAV>(g, f, m, o, s_lights, s_hit[tid], s_pix[tid], n);
    shade_counts_to_global<COUNT>(o, n, s_fetches, s_local, tid, lane);
    // (a wide block records twice its time, and only the first half does: what the tile would cost one block keeps it among the heaviest)
    if (tid == 0 && o.tile_cost && (!wide || (blockIdx.x & 1u) == 0u))
        o.tile_cost[local_tile] = (uint32_t)min((long long)0xffffffffll, (clock64() - t_start) * (wide ? 2 : 1));
}

// ------------------------------------------------------------------------------------------------
// depth field.  Offsets table = computeDepthIndices() render.cpp:66-100, built on the host in the same
// z,y,x order and uploaded to constant memory.
constexpr int MAX_DEPTH_OFFSETS = 1536;
__constant__ float c_off_dist[MAX_DEPTH_OFFSETS];
__constant__ int c_off_xyz[MAX_DEPTH_OFFSETS];      // (dx+8) | (dy+8)<<8 | (dz+8)<<16
__constant__ int c_off_count;

struct EditBox {
    int x0, y0, z0, nx, ny, nz;         // box origin / extents (threads); cells outside the grid are skipped
    int cx, cy, cz, r2;                 // sphere test x^2+y^2+z^2 < r2 relative to (cx,cy,cz); r2 < 0: whole box
};

// render.cpp:226-253 for every cell of the box (optionally restricted to the sphere).  In-place like the
// reference: writers only turn negative values into other negative values, readers only test the sign.
// Out-of-grid neighbours count as solid (oracle/vxo.c fix_depth_field_n).
__device__ __forceinline__ void depth_cell(int32_t* __restrict__ vox, int w, int h, int d, const EditBox& b, long long t) {
    const int bx = (int)(t % b.nx), by = (int)((t / b.nx) % b.ny), bz = (int)(t / ((long long)b.nx * b.ny));
    const int x = b.x0 + bx, y = b.y0 + by, z = b.z0 + bz;
    if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return;
    if (b.r2 >= 0) {
        const int rx = x - b.cx, ry = y - b.cy, rz = z - b.cz;
        if (!(rx * rx + ry * ry + rz * rz < b.r2)) return;
    }
    const int index = x + w * y + w * h * z;                       // w*h*d < 2^31 (the shader indexes with int)
    if (vox[index] >= 0) return;                                   // :231 only empty cells
    float nearest = -6.0f;                                          // :227-228
    const int n = c_off_count;
    for (int i = 0; i < n; i++) {
        const int o = c_off_xyz[i];
        const int xc = x + (o & 255) - 8, yc = y + ((o >> 8) & 255) - 8, zc = z + ((o >> 16) & 255) - 8;
        const bool oob = (unsigned)xc >= (unsigned)w || (unsigned)yc >= (unsigned)h || (unsigned)zc >= (unsigned)d;
        const int v = oob ? 0 : vox[xc + w * yc + w * h * zc];
        const float dist = c_off_dist[i];
        if (v >= 0 && dist > nearest) nearest = (dist <= -2.0f) ? dist : 0.0f;     // :240-247
    }
    if (nearest < 0.0f) vox[index] = __float_as_int(nearest);       // :249-251
}
__global__ void __launch_bounds__(256) depth_kernel(int32_t* __restrict__ vox, int w, int h, int d, EditBox b) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)b.nx * b.ny * b.nz) depth_cell(vox, w, h, d, b, t);
}

// level.cpp:31-42: cells of the lopsided sphere become -1 (destroyVoxel render.cpp:265-271)
__global__ void carve_kernel(int32_t* __restrict__ vox, int w, int h, int d, EditBox b) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.nx * b.ny * b.nz) return;
    const int x = b.x0 + t % b.nx, y = b.y0 + (t / b.nx) % b.ny, z = b.z0 + t / (b.nx * b.ny);
    if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return;
    const int rx = x - b.cx, ry = y - b.cy, rz = z - b.cz;
    if (rx * rx + ry * ry + rz * rz < b.r2) vox[x + w * y + w * h * z] = -1;
}

// ---- removeSphere from a 16-byte command in DEVICE memory (multi-GPU: the command arrives by an NCCL broadcast on the render
// stream, the replicas replay it without the host ever reading it).  cmd = {cx, cy, cz, radius}; the grids are sized for
// max_r by the host, threads beyond the command's own box retire; a radius outside [0, max_r] is refused (*err = 1).
__device__ __forceinline__ bool edit_cmd(const int* __restrict__ cmd, int max_r, int& cx, int& cy, int& cz, int& r, int* err) {
    cx = cmd[0]; cy = cmd[1]; cz = cmd[2]; r = cmd[3];
    if (r < 0 || r > max_r) { if (blockIdx.x == 0 && threadIdx.x == 0) *err = 1; return false; }
    return true;
}
__global__ void __launch_bounds__(256) carve_cmd_kernel(int32_t* __restrict__ vox, int w, int h, int d, const int* __restrict__ cmd, int max_r, int* err) {
    int cx, cy, cz, r;
    if (!edit_cmd(cmd, max_r, cx, cy, cz, r, err) || r == 0) return;
    const int side = 2 * r, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= side * side * side) return;
    const int x = cx - r + t % side, y = cy - r + (t / side) % side, z = cz - r + t / (side * side);
    if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return;
    const int rx = x - cx, ry = y - cy, rz = z - cz;
    if (rx * rx + ry * ry + rz * rz < r * r) vox[x + w * y + w * h * z] = -1;          // level.cpp:31-42
}
__global__ void __launch_bounds__(256) depth_cmd_kernel(int32_t* __restrict__ vox, int w, int h, int d, const int* __restrict__ cmd, int max_r, int* err) {
    int cx, cy, cz, r;
    if (!edit_cmd(cmd, max_r, cx, cy, cz, r, err)) return;
    const int r2 = r + (7 >> 1), side = 2 * r2;                                        // level.cpp:43
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)side * side * side) return;
    const EditBox f{cx - r2, cy - r2, cz - r2, side, side, side, cx, cy, cz, r2 * r2};
    depth_cell(vox, w, h, d, f, t);
}

// Pull the part of the grid rays can read (rows 0 .. ytop-1 of every z slab) into L2 ahead of the traversal: a frame
// that starts with a cold L2 otherwise pays an HBM round trip (~1 us) on the critical path of every long ray for each
// new line it touches -- which is what bounds the frame once a GPU renders only 1/4 or 1/8 of it.  One prefetch per
// 128-byte line; ~10 us for the reference level.
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const int32_t* __restrict__ vox, int lines_per_slab, int slab_stride_ints, long long nlines,
                                                          int* __restrict__ sink) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    int acc = 0;
    for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += stride) {
        const long long z = l / lines_per_slab, j = l - z * lines_per_slab;
        const int32_t* p = vox + z * slab_stride_ints + j * 32;
        int v;
        asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));     // one sector request per 128-byte line allocates it in L2
        acc ^= v;
    }
    if (acc == 0x7fffffff && sink) *sink = acc;                       // never true for grid data that has a -1 or a colour; keeps the loads
}

// rows y that hold a solid voxel, over the linear range [first, first+count): block min/max -> atomics on out[0..1]
__global__ void __launch_bounds__(256) yrange_kernel(const int32_t* __restrict__ vox, long long first, long long count, int w, int h,
                                                     int* __restrict__ out) {
    int lo = INT_MAX, hi = INT_MIN;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        if (vox[first + i] >= 0) {
            const int y = (int)(((first + i) / w) % h);
            lo = min(lo, y); hi = max(hi, y);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        if (lo != INT_MAX) atomicMin(&out[0], lo);
        if (hi != INT_MIN) atomicMax(&out[1], hi);
    }
}

__global__ void set_voxel_kernel(int32_t* vox, long long index, int32_t v) { vox[index] = v; }
// a batch of placeVoxel calls (distinct cells)
__global__ void set_voxels_kernel(int32_t* __restrict__ vox, const long long* __restrict__ index, const int32_t* __restrict__ v, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vox[index[i]] = v[i];
}

// one block per uploaded row: staging holds the rows back to back
__global__ void scatter_rows_kernel(int32_t* __restrict__ vox, const int32_t* __restrict__ staging,
                                    const long long* __restrict__ first, int row_len) {
    const long long dst = first[blockIdx.x];
    const int32_t* src = staging + (long long)blockIdx.x * row_len;
    for (int i = threadIdx.x; i < row_len; i += blockDim.x) vox[dst + i] = src[i];
}

// known-answer hook: n independent castRay calls
template <bool TRAV>
__global__ void cast_rays_kernel(GridView g, int n, const float* __restrict__ starts, const float* __restrict__ dirs,
                                 const int32_t* __restrict__ dists, int32_t* __restrict__ ret, float* __restrict__ out7) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RayHit r = (i & 1) ? cast_ray<true, true, false, GridView, TRAV>(g, starts[3 * i], starts[3 * i + 1], starts[3 * i + 2], dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], dists[i])
                             : cast_ray<true, false, false, GridView, TRAV>(g, starts[3 * i], starts[3 * i + 1], starts[3 * i + 2], dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], dists[i]);
    ret[i] = r.idx;
    float nx, ny, nz;
    unpack_normal(r.normal, nx, ny, nz);
    if (r.steps == 0) { nx = 0.0f; ny = 0.0f; nz = 0.0f; }          // hitNormal keeps its initial value when no step ran
    float* o = out7 + 7 * i;
    o[0] = r.hx; o[1] = r.hy; o[2] = r.hz; o[3] = nx; o[4] = ny; o[5] = nz; o[6] = (float)r.steps;
}

__host__ __device__ inline uint64_t splitmix64(uint64_t x);
// self-test of ray.cuh's exact division: random (a, b) inside the fast domain (|b| in [2^-40, 2], |a| in
// [2^-40, 2^31], all exponents / mantissas / signs), div_by(a, b, refined_rcp(b)) against IEEE __fdiv_rn(a, b)
__global__ void division_selftest_kernel(unsigned long long n, unsigned long long seed, unsigned long long* mismatches) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long h1 = splitmix64(seed + 2 * i), h2 = splitmix64(seed + 2 * i + 1);
        // exponent of b in [-40, 0], of a in [-40, 30]; full 23-bit mantissas; half of the a's are "near-integer minus
        // position" style values (difference of an integer and a float of magnitude < 2^10), like the kernel's dividends
        const int eb = 127 - 40 + (int)((h1 >> 40) % 41), ea = 127 - 40 + (int)((h2 >> 40) % 71);
        const float b = __uint_as_float((unsigned)((h1 >> 63) << 31) | ((unsigned)eb << 23) | (unsigned)(h1 & 0x7FFFFF));
        float a = __uint_as_float((unsigned)((h2 >> 63) << 31) | ((unsigned)ea << 23) | (unsigned)(h2 & 0x7FFFFF));
        if (i & 1) {
            const float pos = __uint_as_float((unsigned)((127 + (h2 >> 50) % 10) << 23) | (unsigned)(h2 & 0x7FFFFF));
            a = __fsub_rn((float)(__float2int_rz(pos) + (int)((h1 >> 33) & 1)), pos);
            if (!(fabsf(a) >= VXRT_DIV_LO)) continue;
        }
        const float q = div_by(a, b, refined_rcp(b)), want = __fdiv_rn(a, b);
        if (__float_as_uint(q) != __float_as_uint(want)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// self-test of the step reciprocals (ray.cuh cast_ray, :74-76): refined_rcp(b) against IEEE rcp.rn(b) for EVERY float b in
// [2^-40, 4) -- 42 binades x 2^23 mantissas (the kernels use it for b = |dir + 0.000001| in [2^-40, 2.000001])
__global__ void reciprocal_selftest_kernel(unsigned long long* mismatches) {
    const unsigned long long n = 42ull << 23, stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float b = __uint_as_float(((unsigned)(127 - 40) << 23) + (unsigned)i);       // consecutive bit patterns from 2^-40 up
        if (__float_as_uint(refined_rcp(b)) != __float_as_uint(__frcp_rn(b))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---- peer-memory frame target (multi-GPU without a gather) -----------------------------------------
// The display rank owns {2 raster frames, done[world] flags, consumed counter}; every rank's kernels store their
// pixels straight into the owner's frame over NVLink, then publish completion with a system-scope release; the
// owner acquires all flags before it consumes the frame.  Spins are bounded (~4 s) so that a lost peer cannot
// hang the GPU: on time-out *err is set and the kernel returns.
struct P2PShared {
    unsigned long long consumed;        // frames the owner has finished reading (written by the owner)
    unsigned long long width, height, world;   // the owner's frame extents and world size (importers must match them)
    unsigned long long owner_device;    // 1 + (PCI domain << 16 | bus << 8 | device) of the owner's GPU (0: unknown)
    unsigned long long pad0[11];
    unsigned long long done[16 * 16];   // done[16*r] = frames rank r has completely written (stride 128 B)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long target) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(p) < target) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 4000000000ull) return false;
        __nanosleep(200);
    }
    return true;
}
// all ranks, before rendering frame `seq` into buffer seq%2: the owner must have consumed frame seq-2
__global__ void p2p_wait_consumed_kernel(const P2PShared* sh, unsigned long long seq, int* err) {
    if (seq >= 2 && !spin_until(&sh->consumed, seq - 1)) *err = 1;
}
// The same wait as the head of a frame whose first render kernel is launched with programmatic stream serialization: that kernel's
// blocks start at once (launch_dependents below) and trace their primary rays while this one still waits; they execute
// griddepcontrol.wait -- this grid completed, its writes visible -- before their first global write.  Also zeroes the frame's
// counters (instead of a memset node in the stream).  One block of 64 threads.
__global__ void __launch_bounds__(64) p2p_begin_kernel(const P2PShared* sh, unsigned long long seq, int* err, unsigned int* counters, int nwords,
                                                       unsigned long long* chain_time) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0 && chain_time) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(chain_time[0]));   // the frame's chain starts (vxrt_get_stats)
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) counters[i] = 0u;
    if (threadIdx.x == 0 && seq >= 2 && !spin_until(&sh->consumed, seq - 1)) *err = 1;
}
// all ranks, after rendering frame `seq` (griddepcontrol.wait: no-op unless launched with programmatic stream serialization, where
// it returns once the render kernel ahead has completed and its stores -- the pixels in the owner's frame -- are performed)
__global__ void p2p_signal_done_kernel(P2PShared* sh, int rank, unsigned long long seq, unsigned long long* chain_time) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // (the owner's wait / release may become resident behind this one)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the frame's chain ends here: the stream carries no event records between the chain's kernels (they would break the programmatic
    // dependencies), the times come from the device's clock instead
    if (chain_time) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(chain_time[1]));
    __threadfence_system();
    st_release_sys(&sh->done[16 * rank], seq + 1);
}
// owner: all ranks have written frame `seq`
__global__ void p2p_wait_done_kernel(const P2PShared* sh, int world, unsigned long long seq, int* err) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // (a release queued right behind may become resident now)
    const int r = threadIdx.x;
    if (r < world && !spin_until(&sh->done[16 * r], seq + 1)) *err = 2;
}
// owner: frame `seq` consumed
__global__ void p2p_release_kernel(P2PShared* sh, unsigned long long seq) {
    asm volatile("griddepcontrol.wait;" ::: "memory");                 // no-op unless launched with programmatic stream serialization
    __threadfence_system();
    st_release_sys(&sh->consumed, seq + 1);
}

// gathered: [world][nlocal][8][32] RGBA8 -> raster [height][width]
// completion flag of a host frame: system-scope store after the stream's earlier kernels (posted writes of one device
// stay ordered, so the host sees the pixels before the flag)
__global__ void host_flag_kernel(unsigned long long* flag, unsigned long long seq) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(seq) : "memory");
}

__global__ void assemble_kernel(const uint32_t* __restrict__ gathered, uint32_t* __restrict__ dst, TileMap m) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m.width * m.height) return;
    const int px = p % m.width, py = p / m.width;
    if (m.rows) {                                                   // [world][local strip][8][width]
        const int row = py / TILE_H, rank = row % m.world;
        dst[p] = gathered[(size_t)rank * m.nlocal * TILE_PIX + (size_t)((row / m.world) * TILE_H + (py % TILE_H)) * m.width + px];
        return;
    }
    const int t = (py / TILE_H) * m.tx + (px / TILE_W);
    const int rank = tile_owner(m, t), local = t / m.world;
    dst[p] = gathered[((size_t)rank * m.nlocal + local) * TILE_PIX + (py % TILE_H) * TILE_W + (px % TILE_W)];
}

}  // namespace vxrt

// ================================================================================================
// Procedural levels generated on the device: the reference's default level
// (level.cpp:82-138, incl. its origin-carving quirk) and the synthetic terrain of config C4.
namespace vxrt {

__device__ __forceinline__ int slab_voxel(int x, int y, int z, int s_stone, int s_dirt, int s_grass) {
    const int cv = 5 * ((x + y + z) % 3);                            // level.cpp:89
    if (y <= s_stone) return ((90 + cv) << 16) | ((90 + cv) << 8) | (90 + cv);        // level.cpp:94-103
    if (y <= s_dirt) return ((120 + cv) << 16) | ((100 + cv) << 8);                   // level.cpp:105-114
    if (y <= s_grass) return (10 << 16) | ((130 + cv) << 8) | 10;                     // level.cpp:116-125
    return -1;                                                       // destroyVoxel level.cpp:91
}

// level.cpp:85-128: stone y<=25, dirt y<=33, grass y<=36
__global__ void default_slabs_kernel(int32_t* __restrict__ vox, int w, int h, int d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)w * h * d) return;
    const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((long long)w * h));
    vox[i] = slab_voxel(x, y, z, 25, 33, 36);
}

__device__ __forceinline__ void dev_place(int32_t* vox, int w, int h, int d, int x, int y, int z, int v) {   // render.cpp:256-262
    if (x >= 0 && y >= 0 && z >= 0 && x < w && y < h && z < d) vox[x + w * y + w * h * z] = v;
}
__device__ __forceinline__ void dev_destroy(int32_t* vox, int w, int h, int d, int x, int y, int z) {        // render.cpp:265-271
    if (x >= 0 && y >= 0 && z >= 0 && x < w && y < h && z < d) vox[x + w * y + w * h * z] = -1;
}

// One block per tree (level.cpp:130-137): placeTrunk (level.cpp:59-79) then placeBush (level.cpp:4-27).
// `relative_destroy` reproduces the reference calling destroyVoxel with sphere-RELATIVE coordinates
// (level.cpp:11,64), which carves a small cavity at the world origin; trees never overlap each other or
// that cavity, so blocks are independent.  base_y < 0: take the trunk base from `surface` (terrain mode).
__global__ void trees_kernel(int32_t* __restrict__ vox, int w, int h, int d, int ntx, int base_y,
                             const int* __restrict__ surface, int relative_destroy) {
    const int tx = blockIdx.x % ntx, tz = blockIdx.x / ntx;
    const int x = 30 * (tx + 1), z = 25 * (tz + 1);                  // x % 30 == 0, z % 25 == 0, both >= 10
    if (x >= w - 10 || z >= d - 10) return;
    const int by = base_y >= 0 ? base_y : surface[x + (long long)w * z];
    {   // placeTrunk(ivec3(x + 1 + z % 7, 36, z), (128,100,15), 6): rel x,z in [-2,1), y in [0,6)
        const int px = x + 1 + z % 7, py = by, pz = z;
        for (int t = threadIdx.x; t < 3 * 6 * 3; t += blockDim.x) {
            const int rx = t % 3 - 2, ry = (t / 3) % 6, rz = t / 18 - 2;
            if (rx + px < w && ry + py < h && rz + pz < d) {
                if (relative_destroy) dev_destroy(vox, w, h, d, rx, ry, rz);
                const int m = ((rx + rz) % 2) * 10;                  // C remainder keeps the sign
                dev_place(vox, w, h, d, rx + px, ry + py, rz + pz, ((128 - m) << 16) | ((100 - m) << 8) | 15);
            }
        }
    }
    __syncthreads();                                                 // the bush overwrites trunk cells it overlaps
    {   // placeBush(ivec3(x + z % 7, 36 + 10, z), (15,128,15), 6)
        const int px = x + z % 7, py = by + 10, pz = z;
        for (int t = threadIdx.x; t < 12 * 12 * 12; t += blockDim.x) {
            const int rx = t % 12 - 6, ry = (t / 12) % 12 - 6, rz = t / 144 - 6;
            if (rx + px < w && ry + py < h && rz + pz < d && rx + px >= 0 && ry + py >= 0 && rz + pz >= 0 &&
                rx * rx + ry * ry + rz * rz < 36) {
                if (relative_destroy) dev_destroy(vox, w, h, d, rx, ry, rz);
                dev_place(vox, w, h, d, rx + px, ry + py, rz + pz, (15 << 16) | ((128 - ((rx + ry + rz) % 3) * 20) << 8) | 15);
            }
        }
    }
}

// ---- synthetic terrain (config C4), integer-only so that any implementation reproduces it bit for bit ----
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint32_t lattice16(uint64_t seed, int ix, int iz, int oct) {
    const uint64_t k = seed ^ ((uint64_t)(uint32_t)ix * 73856093ull) ^ ((uint64_t)(uint32_t)iz * 19349663ull) ^ ((uint64_t)oct * 83492791ull);
    return (uint32_t)(splitmix64(k) >> 48);                          // 16-bit lattice value
}
// 5 octaves of bilinear value noise, cell sizes 256,128,64,32,16, weights 1/2,1/4,...; result in [0, 65535]
__host__ __device__ inline uint32_t fbm16(uint64_t seed, int x, int z) {
    uint64_t acc = 0;
    for (int o = 0; o < 5; o++) {
        const int S = 256 >> o;
        const int ix = x / S, iz = z / S, fx = x % S, fz = z % S;
        const uint64_t v00 = lattice16(seed, ix, iz, o), v10 = lattice16(seed, ix + 1, iz, o);
        const uint64_t v01 = lattice16(seed, ix, iz + 1, o), v11 = lattice16(seed, ix + 1, iz + 1, o);
        const uint64_t top = v00 * (uint64_t)(S - fx) + v10 * (uint64_t)fx, bot = v01 * (uint64_t)(S - fx) + v11 * (uint64_t)fx;
        const uint64_t v = (top * (uint64_t)(S - fz) + bot * (uint64_t)fz) / ((uint64_t)S * S);
        acc += v >> (o + 1);
    }
    return (uint32_t)(acc > 65535 ? 65535 : acc);
}
// surface height: h/4 + fbm * (h*3/8) / 65536  -> for h = 1024: [256, 640)
__host__ __device__ inline int terrain_height(uint64_t seed, int x, int z, int h) {
    return h / 4 + (int)(((uint64_t)fbm16(seed, x, z) * (uint64_t)(h * 3 / 8)) >> 16);
}

__global__ void terrain_surface_kernel(int* __restrict__ surface, int w, int h, int d, uint64_t seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * d) return;
    surface[i] = terrain_height(seed, i % w, i / w, h);
}
// material bands relative to the surface s: stone y <= s-11, dirt y <= s-3 (8 layers), grass y <= s (3 layers)
__global__ void terrain_slabs_kernel(int32_t* __restrict__ vox, const int* __restrict__ surface, int w, int h, int d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)w * h * d) return;
    const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((long long)w * h));
    const int s = surface[x + (long long)w * z];
    vox[i] = slab_voxel(x, y, z, s - 11, s - 3, s);
}

}  // namespace vxrt
