// vxrt.cu -- C ABI of libvxrt.so (include/vxrt.h): host side of the B200 per-pixel path.
// Replaces the GL upload / uniform / draw calls of the reference's src/render.cpp; each entry point cites
// the reference line it stands in for.  There is NO CPU fallback: without an sm_100 device vxrt_create fails.
#include "../../include/vxrt.h"
#include "kernels.cuh"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <thread>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace vxrt;

static_assert(sizeof(FrameParams) == sizeof(vxrt_frame), "FrameParams must mirror vxrt_frame");
static_assert(VXRT_TILE_W == TILE_W && VXRT_TILE_H == TILE_H, "tile constants");

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(VXRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

constexpr int MAX_BANDS = 16;

struct vxrt_ctx {
    vxrt_config cfg{};
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;         // device->host read-back of finished bands, overlapped with rendering
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_band[MAX_BANDS] = {};        // end of band b's kernels (the band's read-back waits for it)
    cudaEvent_t ev_band_start[MAX_BANDS] = {};
    float band_ms[MAX_BANDS] = {};              // kernel time of each band in the last banded frame ...
    int band_ms_n = 0;                          // ... rendered with this many bands (0: none yet)
    cudaEvent_t ev_copy = nullptr;
    int readback_bands = 2;
    // pipelined read-back (vxrt_submit_frame_host): second device frame + per-slot "copy finished" events
    uint32_t* d_rgba8_alt = nullptr;
    cudaEvent_t ev_slot[2] = {nullptr, nullptr};
    bool slot_busy[2] = {false, false};
    unsigned long long submit_seq = 0;
    // grid: d_vox = the reference-layout master copy (uploads, downloads, edits, depth field); d_trav = the traversal grid the
    // rays read (trav.cuh), rebuilt over the affected cells after every change to d_vox
    int32_t* d_vox = nullptr;
    int32_t* d_trav = nullptr;
    unsigned long long* d_trav_bad = nullptr;
    unsigned long long trav_bad = 0;    // cells whose value cannot be encoded (then the rays read d_vox with the plain kernels)
    int trav_mode = 2;                  // vxrt_set_traversal: 0 off, 1 both passes, 2 auto (shade pass always; primary pass when this context
                                        // renders <= 12,000 tiles: there its runs shorten the critical path, on a whole 4K frame they cost throughput)
    size_t nvox = 0;
    bool grid_loaded = false;
    int yrange[2] = {INT_MAX, INT_MIN};  // rows holding solid voxels (never shrinks on destruction: conservative)
    int* d_yrange = nullptr;            // [0..1] the range, [2] sink of the L2 sweep, [3] "a device-side edit command was refused"
    bool edit_err_zeroed = false;
    // frame state
    vxrt_frame frame{};
    TileMap map{};
    // per-frame buffers
    uint32_t* d_rgba8 = nullptr;
    size_t out_pixels = 0;              // entries of d_rgba8
    float4* d_hitq = nullptr;
    uint32_t* d_hitpix = nullptr;
    Counters* d_counters = nullptr;
    uint32_t* d_tile_cost = nullptr;    // per local tile: block cycles of the last primary pass
    uint32_t* d_tile_order = nullptr;   // launch order derived from it (longest first)
    uint32_t* d_tile_hits = nullptr;    // per local tile: hit pixels queued for the shade pass
    uint32_t* d_shade_cost = nullptr;   // per shade unit: block cycles of the last shade pass
    uint32_t* d_shade_order = nullptr;
    bool have_tile_order = false, have_shade_order = false;
    uint32_t* d_tile_order_back = nullptr;     // the sorts write here (side stream); adopted by a later frame (adopt_launch_orders)
    uint32_t* d_shade_order_back = nullptr;
    uint32_t* d_cost_snap = nullptr;           // snapshot of the block times the sorts read: [nlocal] tiles, then [nlocal * 4] shade units
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_order_src = nullptr, ev_order_done = nullptr;
    bool order_pending = false, pend_tiles = false, pend_shade = false;
    int pend_shade_threads = 0;
    int order_shade_threads = 0;        // block size the shade order was recorded with
    unsigned long long order_frame = 0; // whole-frame launches since the ordering was (re)started
    bool use_tile_order = true;
    int fusion = 2;                     // vxrt_set_fusion: 0 two passes, 1 one fused kernel per frame, 2 auto (fused when this context renders
                                        // <= 12,000 tiles: a tile's lighting then starts the moment its own primary rays are done)
    int overlap = 0;                    // vxrt_set_overlap: 0 off (default), 1 on, 2 auto (on when this context renders <= 12,000 tiles) -- the shade
                                        // pass starts inside the primary pass's tail (programmatic dependent launch + per-tile flags)
    uint32_t* d_tile_ready = nullptr;   // per local tile: frame_seq of the last primary pass that finished it
    int2* d_tile_xy = nullptr;          // per local tile: pixel origin of its global tile (TileMap::tile_xy)
    int* d_overlap_err = nullptr;
    uint32_t frame_seq = 0;
    bool last_fused = false;            // the last frame ran as one fused kernel
    bool row_partition = false;         // vxrt_set_partition: whole tile rows are the interleaved unit
    int wide_tiles = 8;                 // vxrt_set_wide_tiles: how many of the heaviest tiles of a fused frame get two blocks / two threads per hit pixel
    bool use_culling = true;
    int l2_prefetch = 2;                // vxrt_set_l2_prefetch: 0 off, 1 on, 2 auto (on when this context renders <= 12,000 tiles)
    int shade_threads = 128;            // threads per shade block (VXRT_SHADE_THREADS: 64 / 128 / 256; 128 measured best)
    int32_t* d_dbg_hit = nullptr;
    uint16_t* d_dbg_steps = nullptr;
    uint32_t* d_dbg_occl = nullptr;
    uint32_t* d_dbg_cast = nullptr;
    // staging for delta uploads
    int32_t* h_stage = nullptr;         // pinned
    int32_t* d_stage = nullptr;
    long long* h_first = nullptr;       // pinned
    long long* d_first = nullptr;
    size_t stage_cap = 0, first_cap = 0;
    uint8_t* h_frame = nullptr;         // pinned read-back buffer
    size_t h_frame_cap = 0;
    bool rendered = false;
    const uint32_t* d_last_frame = nullptr;     // the device buffer the last frame was rendered into (vxrt_read_rgba8 / vxrt_write_ppm)
    int bands_used = 1;
    // peer-memory frame target (vxrt_p2p_*)
    bool p2p = false, p2p_owner = false, p2p_attached = false;
    uint8_t* p2p_base = nullptr;        // owner: cudaMalloc'ed; importer: cudaIpcOpenMemHandle'd
    size_t p2p_frame_bytes = 0;
    unsigned long long p2p_seq = 0;     // frames rendered into the target so far
    int* d_p2p_err = nullptr;
    unsigned long long* d_chain_time = nullptr;   // [2]: globaltimer at the start / end of a programmatically chained frame
    bool chain_timed = false, chain_has_ev1 = false;
    bool p2p_begin_overlap = false;     // the frame's first render kernel may start while p2p_begin_kernel waits (importers on another device only)
    bool p2p_pdl = true;                // flag protocol chained with programmatic dependent launches (VXRT_P2P_PDL=0: separate plain launches)
    int stats_mode = 0;                 // vxrt_set_stats: 0 = production kernels (no per-iteration counters); 1 = counted variants, every ray
                                        // the reference casts is marched to its end (ray / fetch counts by the reference's casting rule);
                                        // 2 = counted variants that skip what the production kernels skip (counts of what is executed)
    uint32_t launches = 0;
};

// ---- helpers -----------------------------------------------------------------------------------
static TileMap make_map(int width, int height, int rank, int world, bool rows) {
    TileMap m;
    m.width = width; m.height = height;
    m.tx = (width + TILE_W - 1) / TILE_W; m.ty = (height + TILE_H - 1) / TILE_H;
    m.ntiles = m.tx * m.ty;
    m.rank = rank; m.world = world;
    m.rows = (rows && world > 1) ? 1 : 0;
    m.nlocal = m.rows ? ((m.ty + world - 1) / world) * m.tx : (m.ntiles + world - 1) / world;
    m.tile_base = 0;
    m.tile_xy = nullptr;
    return m;
}

static bool use_trav(const vxrt_ctx* c) { return c->trav_mode != 0 && c->trav_bad == 0 && c->d_trav != nullptr; }
static bool use_trav_primary(const vxrt_ctx* c) { return use_trav(c) && (c->trav_mode == 1 || c->map.nlocal <= 12000); }

static GridView grid_view(const vxrt_ctx* c) {
    GridView g;
    g.vox = use_trav(c) ? c->d_trav : c->d_vox; g.w = c->cfg.grid_w; g.h = c->cfg.grid_h; g.d = c->cfg.grid_d;
    g.wh = g.w * g.h; g.n = g.w * g.h * g.d;
    g.ymin = c->yrange[0]; g.ymax = c->yrange[1];
    // "every row may hold a solid": nothing is ever culled -- culling switched off, or counted variants that follow the
    // reference's casting rule (stats mode 1, debug planes)
    const bool counted_full = c->stats_mode == 1 || (c->stats_mode == 0 && c->d_dbg_hit != nullptr);
    // (the step-count view shows the number of iterations: its rays always march to their end)
    if (!c->use_culling || counted_full || c->frame.view_depth_field == 1) { g.ymin = INT_MIN / 2; g.ymax = INT_MAX / 2; }
    return g;
}

static void drop_pending_orders(vxrt_ctx* c);
static void free_frame_buffers(vxrt_ctx* c) {
    drop_pending_orders(c);
    cudaFree(c->d_tile_order_back); cudaFree(c->d_shade_order_back); c->d_tile_order_back = nullptr; c->d_shade_order_back = nullptr;
    cudaFree(c->d_cost_snap); c->d_cost_snap = nullptr;
    cudaFree(c->d_rgba8); cudaFree(c->d_rgba8_alt); cudaFree(c->d_hitq); cudaFree(c->d_hitpix);
    cudaFree(c->d_tile_cost); cudaFree(c->d_tile_order); cudaFree(c->d_tile_hits); cudaFree(c->d_shade_cost); cudaFree(c->d_shade_order);
    cudaFree(c->d_tile_ready); c->d_tile_ready = nullptr;
    cudaFree(c->d_tile_xy); c->d_tile_xy = nullptr; c->map.tile_xy = nullptr;
    c->d_last_frame = nullptr;
    c->d_tile_cost = nullptr; c->d_tile_order = nullptr; c->d_tile_hits = nullptr; c->d_shade_cost = nullptr; c->d_shade_order = nullptr;
    c->have_tile_order = false; c->have_shade_order = false; c->order_frame = 0;
    c->d_rgba8_alt = nullptr; c->slot_busy[0] = c->slot_busy[1] = false;
    cudaFree(c->d_dbg_hit); cudaFree(c->d_dbg_steps); cudaFree(c->d_dbg_occl); cudaFree(c->d_dbg_cast);
    c->d_rgba8 = nullptr; c->d_hitq = nullptr; c->d_hitpix = nullptr;
    c->d_dbg_hit = nullptr; c->d_dbg_steps = nullptr; c->d_dbg_occl = nullptr; c->d_dbg_cast = nullptr;
    if (c->h_frame) cudaFreeHost(c->h_frame);
    c->h_frame = nullptr; c->h_frame_cap = 0;
}

static int alloc_frame_buffers(vxrt_ctx* c) {
    free_frame_buffers(c);
    c->map = make_map(c->cfg.width, c->cfg.height, c->cfg.rank, c->cfg.world, c->row_partition);
    const size_t npix = (size_t)c->cfg.width * c->cfg.height;
    const size_t local_pix = (size_t)c->map.nlocal * TILE_PIX;
    {   // the partition as a table: pixel origin of every local tile's global tile (y < 0: padding)
        std::vector<int2> xy((size_t)c->map.nlocal);
        for (int j = 0; j < c->map.nlocal; j++) {
            const int t = tile_of(c->map, j);
            xy[(size_t)j] = (t < c->map.ntiles) ? make_int2((t % c->map.tx) * TILE_W, (t / c->map.tx) * TILE_H) : make_int2(0, -1);
        }
        CUDA_TRY(cudaMalloc(&c->d_tile_xy, xy.size() * sizeof(int2)));
        CUDA_TRY(cudaMemcpyAsync(c->d_tile_xy, xy.data(), xy.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->map.tile_xy = c->d_tile_xy;
    }
    c->out_pixels = (c->cfg.world == 1) ? npix : local_pix;
    CUDA_TRY(cudaMalloc(&c->d_rgba8, c->out_pixels * 4));
    CUDA_TRY(cudaMemsetAsync(c->d_rgba8, 0, c->out_pixels * 4, c->stream));
    CUDA_TRY(cudaMalloc(&c->d_hitq, local_pix * sizeof(float4)));
    CUDA_TRY(cudaMalloc(&c->d_hitpix, local_pix * 4));
    CUDA_TRY(cudaMalloc(&c->d_tile_cost, (size_t)c->map.nlocal * 4));
    CUDA_TRY(cudaMalloc(&c->d_tile_order, (size_t)c->map.nlocal * 4));
    CUDA_TRY(cudaMalloc(&c->d_tile_order_back, (size_t)c->map.nlocal * 4));
    CUDA_TRY(cudaMalloc(&c->d_shade_order_back, (size_t)c->map.nlocal * 4 * 4));
    CUDA_TRY(cudaMalloc(&c->d_cost_snap, (size_t)c->map.nlocal * 5 * 4));
    CUDA_TRY(cudaMalloc(&c->d_tile_hits, (size_t)c->map.nlocal * 4));
    CUDA_TRY(cudaMalloc(&c->d_tile_ready, (size_t)c->map.nlocal * 4));
    CUDA_TRY(cudaMemsetAsync(c->d_tile_ready, 0, (size_t)c->map.nlocal * 4, c->stream));
    if (!c->d_overlap_err) { CUDA_TRY(cudaMalloc(&c->d_overlap_err, sizeof(int))); CUDA_TRY(cudaMemsetAsync(c->d_overlap_err, 0, sizeof(int), c->stream)); }
    c->frame_seq = 0;
    CUDA_TRY(cudaMalloc(&c->d_shade_cost, (size_t)c->map.nlocal * 4 * 4));      // up to 4 units per tile (64-thread blocks)
    CUDA_TRY(cudaMalloc(&c->d_shade_order, (size_t)c->map.nlocal * 4 * 4));
    CUDA_TRY(cudaMemsetAsync(c->d_tile_hits, 0, (size_t)c->map.nlocal * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_shade_cost, 0, (size_t)c->map.nlocal * 4 * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_tile_cost, 0, (size_t)c->map.nlocal * 4, c->stream));
    if (c->cfg.flags & VXRT_FLAG_DEBUG_OUTPUTS) {
        CUDA_TRY(cudaMalloc(&c->d_dbg_hit, npix * 4));
        CUDA_TRY(cudaMalloc(&c->d_dbg_steps, npix * 2));
        CUDA_TRY(cudaMalloc(&c->d_dbg_occl, npix * 4));
        CUDA_TRY(cudaMalloc(&c->d_dbg_cast, npix * 4));
        CUDA_TRY(cudaMemsetAsync(c->d_dbg_hit, 0xFF, npix * 4, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->d_dbg_steps, 0, npix * 2, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->d_dbg_occl, 0, npix * 4, c->stream));
        CUDA_TRY(cudaMemsetAsync(c->d_dbg_cast, 0, npix * 4, c->stream));
    }
    c->h_frame_cap = c->out_pixels * 4;
    CUDA_TRY(cudaMallocHost(&c->h_frame, c->h_frame_cap));
    c->rendered = false;
    return VXRT_OK;
}

// computeDepthIndices() render.cpp:66-100: same z,y,x order, dist = (float)-sqrt((double)k)
static int upload_depth_offsets() {
    std::vector<float> dist; std::vector<int> xyz;
    const int R = 7;
    for (int zc = -R; zc <= R; zc++)
        for (int yc = -R; yc <= R; yc++)
            for (int xc = -R; xc <= R; xc++)
                if (xc * xc + yc * yc + zc * zc <= R * R) {
                    const int xd = xc - (xc > 0) + (xc < 0), yd = yc - (yc > 0) + (yc < 0), zd = zc - (zc > 0) + (zc < 0);
                    const float d = (float)(-std::sqrt((double)(xd * xd + yd * yd + zd * zd)));
                    if (d <= (float)R) { dist.push_back(d); xyz.push_back((xc + 8) | ((yc + 8) << 8) | ((zc + 8) << 16)); }
                }
    const int n = (int)dist.size();
    if (n > MAX_DEPTH_OFFSETS) return fail(VXRT_ERR_INVALID, "depth offset table overflow");
    CUDA_TRY(cudaMemcpyToSymbol(c_off_dist, dist.data(), n * sizeof(float)));
    CUDA_TRY(cudaMemcpyToSymbol(c_off_xyz, xyz.data(), n * sizeof(int)));
    CUDA_TRY(cudaMemcpyToSymbol(c_off_count, &n, sizeof(int)));
    return VXRT_OK;
}

static int ensure_stage(vxrt_ctx* c, size_t elems, size_t rows) {
    if (elems > c->stage_cap) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        cudaFree(c->d_stage);
        c->h_stage = nullptr; c->d_stage = nullptr; c->stage_cap = 0;
        size_t cap = elems < 65536 ? 65536 : elems;
        CUDA_TRY(cudaMallocHost(&c->h_stage, cap * 4));
        CUDA_TRY(cudaMalloc(&c->d_stage, cap * 4));
        c->stage_cap = cap;
    }
    if (rows > c->first_cap) {
        if (c->h_first) cudaFreeHost(c->h_first);
        cudaFree(c->d_first);
        c->h_first = nullptr; c->d_first = nullptr; c->first_cap = 0;
        size_t cap = rows < 4096 ? 4096 : rows;
        CUDA_TRY(cudaMallocHost(&c->h_first, cap * sizeof(long long)));
        CUDA_TRY(cudaMalloc(&c->d_first, cap * sizeof(long long)));
        c->first_cap = cap;
    }
    return VXRT_OK;
}

// occupancy summary used by cast_ray's CULL: expand (or, reset = true, recompute) the range of rows with solid voxels
// from the device grid's linear range [first, first+count)
static int update_yrange(vxrt_ctx* c, size_t first, size_t count, bool reset) {
    if (!c->d_yrange) CUDA_TRY(cudaMalloc(&c->d_yrange, 4 * sizeof(int)));      // [0..1] the range, [2] sink of the L2 sweep
    if (reset) { c->yrange[0] = INT_MAX; c->yrange[1] = INT_MIN; }
    CUDA_TRY(cudaMemcpyAsync(c->d_yrange, c->yrange, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (count) {
        const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 148 * 16);
        yrange_kernel<<<blocks, 256, 0, c->stream>>>(c->d_vox, (long long)first, (long long)count, c->cfg.grid_w, c->cfg.grid_h, c->d_yrange);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(c->yrange, c->d_yrange, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

// ---- traversal grid maintenance ------------------------------------------------------------------------------
// Rebuild the traversal words of every cell whose word can change when the cells of the box [x0,x1) x [y0,y1) x [z0,z1) of
// d_vox changed: the box grown by TRAV_REACH in x and z and by one layer downward (a band word looks at most that far).
// full: the whole grid (the count of values that cannot be encoded starts over).  readback: fetch that count (synchronises).
static int trav_sync(vxrt_ctx* c, int x0, int y0, int z0, int x1, int y1, int z1, bool full, bool readback) {
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    if (full) {
        x0 = 0; y0 = 0; z0 = 0; x1 = W; y1 = H; z1 = D;
        CUDA_TRY(cudaMemsetAsync(c->d_trav_bad, 0, sizeof(unsigned long long), c->stream));
    } else {
        x0 -= TRAV_REACH; x1 += TRAV_REACH; z0 -= TRAV_REACH; z1 += TRAV_REACH; y0 -= 1;
    }
    x0 = std::max(x0, 0); y0 = std::max(y0, 0); z0 = std::max(z0, 0);
    x1 = std::min(x1, W); y1 = std::min(y1, H); z1 = std::min(z1, D);
    if (x1 > x0 && y1 > y0 && z1 > z0) {
        TravBox b{x0, y0, z0, x1 - x0, y1 - y0, z1 - z0};
        const long long n = (long long)b.nx * b.ny * b.nz;
        trav_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_vox, c->d_trav, W, H, D, b, c->d_trav_bad);
        CUDA_TRY(cudaGetLastError());
    }
    if (readback) {
        unsigned long long bad = 0;
        CUDA_TRY(cudaMemcpyAsync(&bad, c->d_trav_bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->trav_bad = bad;
    }
    return VXRT_OK;
}
// the cells of the linear index range [first, first + count)
static int trav_sync_range(vxrt_ctx* c, size_t first, size_t count, bool readback) {
    if (count == 0) return VXRT_OK;
    const size_t W = (size_t)c->cfg.grid_w, WH = W * (size_t)c->cfg.grid_h;
    const size_t last = first + count - 1;
    const int z0 = (int)(first / WH), z1 = (int)(last / WH) + 1;
    int x0 = 0, x1 = (int)W, y0 = 0, y1 = c->cfg.grid_h;
    if (z1 - z0 == 1) {
        y0 = (int)((first / W) % (size_t)c->cfg.grid_h); y1 = (int)((last / W) % (size_t)c->cfg.grid_h) + 1;
        if (y1 - y0 == 1) { x0 = (int)(first % W); x1 = (int)(last % W) + 1; }
    }
    return trav_sync(c, x0, y0, z0, x1, y1, z1, false, readback);
}

static int host_index(const vxrt_config& g, int x, int y, int z) {     // getVoxelIndex render.cpp:189-196
    if (x >= 0 && y >= 0 && z >= 0 && x < g.grid_w && y < g.grid_h && z < g.grid_d)
        return x + g.grid_w * y + g.grid_w * g.grid_h * z;
    return -1;
}

#define CHECK_CTX(c)                                                            \
    do {                                                                        \
        if (!(c)) return fail(VXRT_ERR_INVALID, "null context");                \
        CUDA_TRY(cudaSetDevice((c)->cfg.device));                               \
    } while (0)

// ---- lifetime ----------------------------------------------------------------------------------
extern "C" int vxrt_device_available(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return 0; }
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) return 1;
    }
    return 0;
}

extern "C" const char* vxrt_last_error(void) { return g_last_error.c_str(); }

extern "C" void* vxrt_host_alloc(size_t nbytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, nbytes) != cudaSuccess) { cudaGetLastError(); g_last_error = "cudaMallocHost failed"; return nullptr; }
    return p;
}
extern "C" void vxrt_host_free(void* p) { if (p) cudaFreeHost(p); }

// FNV-1a-64 as SURVEY.md 8c fingerprints grids: prime 1099511628211, offset basis 1469598103934665603 (the survey's
// basis, one digit short of the textbook 14695981039346656037 -- kept, since every recorded fingerprint uses it)
static const uint64_t FNV_BASIS = 1469598103934665603ull;
static uint64_t fnv1a64_update(uint64_t h, const void* data, size_t nbytes) {
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < nbytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
extern "C" uint64_t vxrt_fnv1a64(const void* data, size_t nbytes) { return fnv1a64_update(FNV_BASIS, data, nbytes); }

extern "C" int vxrt_create(const vxrt_config* cfg, vxrt_ctx** out) {
    if (!cfg || !out) return fail(VXRT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->grid_w <= 0 || cfg->grid_h <= 0 || cfg->grid_d <= 0 || cfg->width <= 0 || cfg->height <= 0)
        return fail(VXRT_ERR_INVALID, "grid and frame extents must be positive");
    if ((long long)cfg->grid_w * cfg->grid_h * cfg->grid_d > 0x7fffffffLL)
        return fail(VXRT_ERR_INVALID, "grid too large: the shader indexes voxels with a 32-bit int (fshader.glsl:33-52)");
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world)
        return fail(VXRT_ERR_INVALID, "bad rank/world");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(VXRT_ERR_NO_DEVICE, "no CUDA device: libvxrt has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(VXRT_ERR_INVALID, "bad device ordinal");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(VXRT_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100: libvxrt is built for sm_100a only");
    CUDA_TRY(cudaSetDevice(cfg->device));
    vxrt_ctx* c = new vxrt_ctx();
    c->cfg = *cfg;
    c->nvox = (size_t)cfg->grid_w * cfg->grid_h * cfg->grid_d;
    auto bail = [&](int code) { vxrt_destroy(c); return code; };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaStreamCreate failed"));
    for (auto& e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaMalloc(&c->d_vox, c->nvox * 4) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaMalloc(grid) failed"));
    if (cudaMalloc(&c->d_trav, c->nvox * 4) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaMalloc(traversal grid) failed"));
    if (cudaMalloc(&c->d_trav_bad, sizeof(unsigned long long)) != cudaSuccess || cudaMemset(c->d_trav_bad, 0, sizeof(unsigned long long)) != cudaSuccess)
        return bail(fail(VXRT_ERR_CUDA, "cudaMalloc(traversal counter) failed"));
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaStreamCreate failed"));
    for (auto& e : c->ev_band) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaEventCreate failed"));
    for (auto& e : c->ev_band_start) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaEventCreate failed"));
    for (auto& e : c->ev_slot) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaStreamCreate failed"));
    if (cudaEventCreateWithFlags(&c->ev_order_src, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_order_done, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(VXRT_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaMalloc(&c->d_counters, sizeof(Counters) * MAX_BANDS) != cudaSuccess) return bail(fail(VXRT_ERR_CUDA, "cudaMalloc(counters) failed"));
    int rc = upload_depth_offsets();
    if (rc != VXRT_OK) return bail(rc);
    rc = alloc_frame_buffers(c);
    if (rc != VXRT_OK) return bail(rc);
    // frame defaults == the reference's globals, main.cpp:26-37
    memset(&c->frame, 0, sizeof c->frame);
    c->frame.cam_pos[0] = 195.0f; c->frame.cam_pos[1] = 55.0f; c->frame.cam_pos[2] = 155.0f;
    c->frame.light_pos[0] = cfg->grid_w / 2.0f; c->frame.light_pos[1] = cfg->grid_w * 3.0f; c->frame.light_pos[2] = cfg->grid_w / 2.0f;
    c->frame.aspect = (float)cfg->width / cfg->height;
    for (int i = 0; i < 4; i++) c->frame.rotate[5 * i] = 1.0f;
    vxrt_init_local_lights(c);
    if (const char* e = getenv("VXRT_L2_PREFETCH")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->l2_prefetch = v; }
    if (const char* e = getenv("VXRT_TRAVERSAL")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->trav_mode = v; }
    if (const char* e = getenv("VXRT_OVERLAP")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->overlap = v; }
    if (const char* e = getenv("VXRT_FUSION")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->fusion = v; }
    if (const char* e = getenv("VXRT_P2P_PDL")) c->p2p_pdl = atoi(e) != 0;
    if (const char* e = getenv("VXRT_WIDE_TILES")) { const int v = atoi(e); if (v >= 0 && v <= 64) c->wide_tiles = v; }
    if (const char* e = getenv("VXRT_SHADE_THREADS")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) c->shade_threads = v;
    }
    *out = c;
    return VXRT_OK;
}

extern "C" void vxrt_destroy(vxrt_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);     // before the buffers it copies from go away
    free_frame_buffers(c);
    if (c->p2p_base) { if (c->p2p_owner) cudaFree(c->p2p_base); else if (!c->p2p_attached) cudaIpcCloseMemHandle(c->p2p_base); }
    cudaFree(c->d_p2p_err);
    cudaFree(c->d_chain_time);
    cudaFree(c->d_overlap_err);
    cudaFree(c->d_yrange);
    cudaFree(c->d_vox); cudaFree(c->d_trav); cudaFree(c->d_trav_bad); cudaFree(c->d_counters); cudaFree(c->d_stage); cudaFree(c->d_first);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_first) cudaFreeHost(c->h_first);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_band) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_band_start) if (e) cudaEventDestroy(e);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    for (auto& e : c->ev_slot) if (e) cudaEventDestroy(e);
    if (c->ev_order_src) cudaEventDestroy(c->ev_order_src);
    if (c->ev_order_done) cudaEventDestroy(c->ev_order_done);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---- grid --------------------------------------------------------------------------------------
extern "C" int vxrt_upload_grid(vxrt_ctx* c, const int32_t* voxels, size_t count) {
    CHECK_CTX(c);
    if (!voxels || count != c->nvox) return fail(VXRT_ERR_INVALID, "upload_grid: count must equal grid_w*grid_h*grid_d");
    CUDA_TRY(cudaMemcpyAsync(c->d_vox, voxels, count * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));      // GL semantics: the caller may modify its array on return
    c->grid_loaded = true;
    int rc = trav_sync(c, 0, 0, 0, 0, 0, 0, true, true);
    if (rc != VXRT_OK) return rc;
    return update_yrange(c, 0, c->nvox, true);
}

extern "C" int vxrt_upload_range(vxrt_ctx* c, size_t first, size_t count, const int32_t* src) {
    CHECK_CTX(c);
    if (!src) return fail(VXRT_ERR_INVALID, "upload_range: null source");
    if (first > c->nvox || count > c->nvox - first) return fail(VXRT_ERR_INVALID, "upload_range: range outside the buffer (GL_INVALID_VALUE)");
    CUDA_TRY(cudaMemcpyAsync(c->d_vox + first, src, count * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int rc = trav_sync_range(c, first, count, true);
    if (rc != VXRT_OK) return rc;
    return update_yrange(c, first, count, false);
}

// rows of `row_len` voxels staged back to back in h_stage, their first grid indices in h_first: one copy of each to the
// device, one scatter kernel (one block per row), then the occupancy summary takes in any new solids
static int scatter_staged_rows(vxrt_ctx* c, size_t rows, int row_len) {
    const size_t elems = rows * (size_t)row_len;
    CUDA_TRY(cudaMemcpyAsync(c->d_stage, c->h_stage, elems * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_first, c->h_first, rows * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    scatter_rows_kernel<<<(unsigned)rows, 64, 0, c->stream>>>(c->d_vox, c->d_stage, c->d_first, row_len);
    CUDA_TRY(cudaGetLastError());
    {   // traversal words around the rows: one rebuild over their bounding box
        const size_t W = (size_t)c->cfg.grid_w, WH = W * (size_t)c->cfg.grid_h;
        int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
        for (size_t r = 0; r < rows; r++) {
            const size_t f = (size_t)c->h_first[r], l = f + (size_t)row_len - 1;
            const int z0 = (int)(f / WH), z1 = (int)(l / WH);
            int y0 = 0, y1 = c->cfg.grid_h - 1, x0 = 0, x1 = (int)W - 1;
            if (z0 == z1) {
                y0 = (int)((f / W) % (size_t)c->cfg.grid_h); y1 = (int)((l / W) % (size_t)c->cfg.grid_h);
                if (y0 == y1) { x0 = (int)(f % W); x1 = (int)(l % W); }
            }
            lo[0] = std::min(lo[0], x0); lo[1] = std::min(lo[1], y0); lo[2] = std::min(lo[2], z0);
            hi[0] = std::max(hi[0], x1); hi[1] = std::max(hi[1], y1); hi[2] = std::max(hi[2], z1);
        }
        int rc = trav_sync(c, lo[0], lo[1], lo[2], hi[0] + 1, hi[1] + 1, hi[2] + 1, false, false);
        if (rc != VXRT_OK) return rc;
        unsigned long long bad = 0;
        CUDA_TRY(cudaMemcpyAsync(&bad, c->d_trav_bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // staging buffers are reused by the next call
        c->trav_bad = bad;
    }
    for (size_t r = 0; r < rows; r++)                 // rows uploaded may hold new solids
        for (int i = 0; i < row_len; i++)
            if (c->h_stage[r * (size_t)row_len + i] >= 0) {
                const int y = (int)(((c->h_first[r] + i) / c->cfg.grid_w) % c->cfg.grid_h);
                if (y < c->yrange[0]) c->yrange[0] = y;
                if (y > c->yrange[1]) c->yrange[1] = y;
            }
    return VXRT_OK;
}

// updatePartialGeometry(start,end) render.cpp:204-223 -- same (int) casts, same float loop counters, same
// start/end swap on linear indices, same "skip the row when its first index is out of bounds".
extern "C" int vxrt_update_partial(vxrt_ctx* c, const float start_in[3], const float end_in[3],
                                   const int32_t* host_voxels, int32_t* rows_out) {
    CHECK_CTX(c);
    if (!start_in || !end_in || !host_voxels) return fail(VXRT_ERR_INVALID, "update_partial: null argument");
    float start[3] = {start_in[0], start_in[1], start_in[2]}, end[3] = {end_in[0], end_in[1], end_in[2]};
    const int startInd = host_index(c->cfg, (int)start[0], (int)start[1], (int)start[2]);
    const int endInd = host_index(c->cfg, (int)end[0], (int)end[1], (int)end[2]);
    if (startInd > endInd) for (int k = 0; k < 3; k++) { const float t = start[k]; start[k] = end[k]; end[k] = t; }
    const int xLength = (int)(end[0] - start[0]) + 1;
    std::vector<long long> firsts;
    for (float i = start[2]; i < end[2]; i++)
        for (float j = start[1]; j < end[1]; j++) {
            const int offset = host_index(c->cfg, (int)start[0], (int)j, (int)i);
            if (offset != -1 && xLength > 0 && (size_t)offset + (size_t)xLength <= c->nvox) firsts.push_back(offset);
        }
    if (rows_out) *rows_out = (int32_t)firsts.size();
    if (firsts.empty()) return VXRT_OK;
    const size_t rows = firsts.size();
    int rc = ensure_stage(c, rows * (size_t)xLength, rows);
    if (rc != VXRT_OK) return rc;
    for (size_t r = 0; r < rows; r++) {
        memcpy(c->h_stage + r * xLength, host_voxels + firsts[r], (size_t)xLength * 4);
        c->h_first[r] = firsts[r];
    }
    return scatter_staged_rows(c, rows, xLength);
}

// A batch of glBufferSubData calls of one length (what one updatePartialGeometry issues, render.cpp:214-221) as ONE
// staged copy + scatter kernel; packed holds the rows back to back, copied at call time like GL does.
extern "C" int vxrt_upload_rows(vxrt_ctx* c, size_t rows, size_t row_len, const int64_t* firsts, const int32_t* packed) {
    CHECK_CTX(c);
    if (rows == 0 || row_len == 0) return VXRT_OK;
    if (!firsts || !packed) return fail(VXRT_ERR_INVALID, "upload_rows: null argument");
    if (row_len > (size_t)INT_MAX || rows > (size_t)INT_MAX) return fail(VXRT_ERR_INVALID, "upload_rows: too many / too long rows");
    for (size_t r = 0; r < rows; r++)
        if (firsts[r] < 0 || (size_t)firsts[r] > c->nvox || row_len > c->nvox - (size_t)firsts[r])
            return fail(VXRT_ERR_INVALID, "upload_rows: row outside the buffer (GL_INVALID_VALUE)");
    int rc = ensure_stage(c, rows * row_len, rows);
    if (rc != VXRT_OK) return rc;
    memcpy(c->h_stage, packed, rows * row_len * 4);
    for (size_t r = 0; r < rows; r++) c->h_first[r] = firsts[r];
    return scatter_staged_rows(c, rows, (int)row_len);
}

extern "C" int vxrt_download_grid(vxrt_ctx* c, int32_t* out, size_t count) {
    CHECK_CTX(c);
    if (!out || count != c->nvox) return fail(VXRT_ERR_INVALID, "download_grid: count must equal grid_w*grid_h*grid_d");
    CUDA_TRY(cudaMemcpyAsync(out, c->d_vox, count * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

// ---- grid files -------------------------------------------------------------------------------
struct GridFileHeader {                 // 64 bytes, little-endian (include/vxrt.h)
    char magic[8];
    uint32_t w, h, d, flags;
    uint64_t count, fnv;
    uint8_t reserved[24];
};
static_assert(sizeof(GridFileHeader) == 64, "grid file header is 64 bytes");
static const size_t GRID_FILE_CHUNK = (size_t)8 << 20;      // voxels per staged copy (32 MiB)

extern "C" int vxrt_save_grid(vxrt_ctx* c, const char* path) {
    CHECK_CTX(c);
    if (!path) return fail(VXRT_ERR_INVALID, "save_grid: null path");
    if (!c->grid_loaded) return fail(VXRT_ERR_STATE, "save_grid before any grid upload");
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(VXRT_ERR_IO, std::string("save_grid: cannot open ") + path);
    GridFileHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "VXRTGRD1", 8);
    hd.w = (uint32_t)c->cfg.grid_w; hd.h = (uint32_t)c->cfg.grid_h; hd.d = (uint32_t)c->cfg.grid_d;
    hd.count = c->nvox;
    int32_t* stage = nullptr;
    const size_t chunk = std::min(GRID_FILE_CHUNK, (size_t)c->nvox);
    if (cudaMallocHost(&stage, chunk * 4) != cudaSuccess) { fclose(fp); cudaGetLastError(); return fail(VXRT_ERR_CUDA, "save_grid: pinned staging allocation failed"); }
    bool ok = fwrite(&hd, sizeof hd, 1, fp) == 1;
    uint64_t h = FNV_BASIS;
    for (size_t first = 0; ok && first < c->nvox; first += chunk) {
        const size_t n = std::min(chunk, (size_t)c->nvox - first);
        if (cudaMemcpyAsync(stage, c->d_vox + first, n * 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { cudaFreeHost(stage); fclose(fp); return fail(VXRT_ERR_CUDA, "save_grid: device read failed"); }
        h = fnv1a64_update(h, stage, n * 4);
        ok = fwrite(stage, 4, n, fp) == n;
    }
    cudaFreeHost(stage);
    hd.fnv = h;
    ok = ok && fseek(fp, 0, SEEK_SET) == 0 && fwrite(&hd, sizeof hd, 1, fp) == 1;
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return fail(VXRT_ERR_IO, std::string("save_grid: write failed on ") + path);
    return VXRT_OK;
}

extern "C" int vxrt_load_grid(vxrt_ctx* c, const char* path) {
    CHECK_CTX(c);
    if (!path) return fail(VXRT_ERR_INVALID, "load_grid: null path");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(VXRT_ERR_IO, std::string("load_grid: cannot open ") + path);
    GridFileHeader hd;
    if (fread(&hd, sizeof hd, 1, fp) != 1 || memcmp(hd.magic, "VXRTGRD1", 8) != 0) { fclose(fp); return fail(VXRT_ERR_IO, "load_grid: not a VXRTGRD1 file"); }
    if (hd.w != (uint32_t)c->cfg.grid_w || hd.h != (uint32_t)c->cfg.grid_h || hd.d != (uint32_t)c->cfg.grid_d || hd.count != (uint64_t)c->nvox) {
        fclose(fp);
        return fail(VXRT_ERR_INVALID, "load_grid: the file's extents differ from the context's grid");
    }
    int32_t* stage = nullptr;
    const size_t chunk = std::min(GRID_FILE_CHUNK, (size_t)c->nvox);
    if (cudaMallocHost(&stage, chunk * 4) != cudaSuccess) { fclose(fp); cudaGetLastError(); return fail(VXRT_ERR_CUDA, "load_grid: pinned staging allocation failed"); }
    c->grid_loaded = false;                                  // a failed load leaves no half-written grid in use
    uint64_t h = FNV_BASIS;
    bool ok = true;
    for (size_t first = 0; ok && first < c->nvox; first += chunk) {
        const size_t n = std::min(chunk, (size_t)c->nvox - first);
        ok = fread(stage, 4, n, fp) == n;
        if (!ok) break;
        h = fnv1a64_update(h, stage, n * 4);
        if (cudaMemcpyAsync(c->d_vox + first, stage, n * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { cudaFreeHost(stage); fclose(fp); return fail(VXRT_ERR_CUDA, "load_grid: device write failed"); }
    }
    cudaFreeHost(stage);
    fclose(fp);
    if (!ok) return fail(VXRT_ERR_IO, "load_grid: file shorter than its header says");
    if (h != hd.fnv) return fail(VXRT_ERR_IO, "load_grid: payload fingerprint does not match the header");
    c->grid_loaded = true;
    int rc = trav_sync(c, 0, 0, 0, 0, 0, 0, true, true);
    if (rc != VXRT_OK) return rc;
    return update_yrange(c, 0, c->nvox, true);
}

extern "C" int vxrt_download_box(vxrt_ctx* c, const int32_t lo[3], const int32_t hi[3], int32_t* host_voxels) {
    CHECK_CTX(c);
    if (!lo || !hi || !host_voxels) return fail(VXRT_ERR_INVALID, "download_box: null argument");
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    const int x0 = lo[0] < 0 ? 0 : lo[0], y0 = lo[1] < 0 ? 0 : lo[1], z0 = lo[2] < 0 ? 0 : lo[2];
    const int x1 = hi[0] > W ? W : hi[0], y1 = hi[1] > H ? H : hi[1], z1 = hi[2] > D ? D : hi[2];
    if (x1 <= x0 || y1 <= y0 || z1 <= z0) return VXRT_OK;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(c->d_vox, (size_t)W * 4, (size_t)W, (size_t)H);
    p.dstPtr = make_cudaPitchedPtr(host_voxels, (size_t)W * 4, (size_t)W, (size_t)H);
    p.srcPos = make_cudaPos((size_t)x0 * 4, (size_t)y0, (size_t)z0);
    p.dstPos = p.srcPos;
    p.extent = make_cudaExtent((size_t)(x1 - x0) * 4, (size_t)(y1 - y0), (size_t)(z1 - z0));
    p.kind = cudaMemcpyDeviceToHost;
    CUDA_TRY(cudaMemcpy3DAsync(&p, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

static bool unencodable(int32_t v) { return v < 0 && v != -1 && !(v & 0x40000000); }      // see trav.cuh

extern "C" int vxrt_place_voxel(vxrt_ctx* c, int x, int y, int z, int32_t voxel) {      // render.cpp:256-262
    CHECK_CTX(c);
    const int index = host_index(c->cfg, x, y, z);
    if (index >= 0) {
        set_voxel_kernel<<<1, 1, 0, c->stream>>>(c->d_vox, index, voxel);
        CUDA_TRY(cudaGetLastError());
        if (voxel >= 0) { if (y < c->yrange[0]) c->yrange[0] = y; if (y > c->yrange[1]) c->yrange[1] = y; }
        if (unencodable(voxel)) c->trav_bad++;
        return trav_sync(c, x, y, z, x + 1, y + 1, z + 1, false, false);
    }
    return VXRT_OK;
}

// n placeVoxel calls (render.cpp:256-262; the reference's level code places voxels in bulk: level.cpp:21,74,102,113,124) as one
// staged copy + one kernel; later entries win where cells repeat, like the sequence of calls would
extern "C" int vxrt_place_voxels(vxrt_ctx* c, size_t n, const int32_t* xyz, const int32_t* voxels) {
    CHECK_CTX(c);
    if (n == 0) return VXRT_OK;
    if (!xyz || !voxels) return fail(VXRT_ERR_INVALID, "place_voxels: null argument");
    if (n > (size_t)INT_MAX / 2) return fail(VXRT_ERR_INVALID, "place_voxels: too many voxels");
    std::vector<long long> idx; std::vector<int32_t> val;
    idx.reserve(n); val.reserve(n);
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (size_t i = 0; i < n; i++) {
        const int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const int index = host_index(c->cfg, x, y, z);
        if (index < 0) continue;                                                         // outside the grid: ignored, like placeVoxel
        idx.push_back(index); val.push_back(voxels[i]);
        lo[0] = std::min(lo[0], x); lo[1] = std::min(lo[1], y); lo[2] = std::min(lo[2], z);
        hi[0] = std::max(hi[0], x); hi[1] = std::max(hi[1], y); hi[2] = std::max(hi[2], z);
        if (voxels[i] >= 0) { if (y < c->yrange[0]) c->yrange[0] = y; if (y > c->yrange[1]) c->yrange[1] = y; }
        if (unencodable(voxels[i])) c->trav_bad++;
    }
    if (idx.empty()) return VXRT_OK;
    // cells that repeat: keep the last value (a kernel's threads are not ordered)
    {
        std::vector<size_t> order(idx.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return idx[a] < idx[b]; });
        std::vector<long long> i2; std::vector<int32_t> v2;
        for (size_t k = 0; k < order.size(); k++)
            if (k + 1 == order.size() || idx[order[k + 1]] != idx[order[k]]) { i2.push_back(idx[order[k]]); v2.push_back(val[order[k]]); }
        idx.swap(i2); val.swap(v2);
    }
    const size_t m = idx.size();
    int rc = ensure_stage(c, m, m);
    if (rc != VXRT_OK) return rc;
    memcpy(c->h_stage, val.data(), m * 4);
    memcpy(c->h_first, idx.data(), m * sizeof(long long));
    CUDA_TRY(cudaMemcpyAsync(c->d_stage, c->h_stage, m * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->d_first, c->h_first, m * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    set_voxels_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->d_vox, c->d_first, c->d_stage, (long long)m);
    CUDA_TRY(cudaGetLastError());
    rc = trav_sync(c, lo[0], lo[1], lo[2], hi[0] + 1, hi[1] + 1, hi[2] + 1, false, false);
    if (rc != VXRT_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));          // staging buffers are reused by the next call
    return VXRT_OK;
}

extern "C" int vxrt_destroy_voxel(vxrt_ctx* c, int x, int y, int z) {                   // render.cpp:265-271
    CHECK_CTX(c);
    const int index = host_index(c->cfg, x, y, z);
    if (x >= 0 && y >= 0 && z >= 0 && index >= 0) {
        set_voxel_kernel<<<1, 1, 0, c->stream>>>(c->d_vox, index, -1);
        CUDA_TRY(cudaGetLastError());
        return trav_sync(c, x, y, z, x + 1, y + 1, z + 1, false, false);
    }
    return VXRT_OK;
}

// removeSphere level.cpp:30-56: loops are [-r, r) per axis, test x^2+y^2+z^2 < r^2; then r += 7>>1 and
// fixDepthField over that larger lopsided sphere.
extern "C" int vxrt_edit_remove_sphere(vxrt_ctx* c, int cx, int cy, int cz, int radius) {
    CHECK_CTX(c);
    if (radius < 0 || radius > 512) return fail(VXRT_ERR_INVALID, "remove_sphere: radius out of range");
    if (!c->grid_loaded) return fail(VXRT_ERR_STATE, "remove_sphere before any grid upload");
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    if (radius > 0) {
        EditBox b{cx - radius, cy - radius, cz - radius, 2 * radius, 2 * radius, 2 * radius, cx, cy, cz, radius * radius};
        const int n = b.nx * b.ny * b.nz;
        carve_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_vox, W, H, D, b);
        CUDA_TRY(cudaGetLastError());
    }
    const int r2 = radius + (7 >> 1);                                                    // level.cpp:43
    EditBox f{cx - r2, cy - r2, cz - r2, 2 * r2, 2 * r2, 2 * r2, cx, cy, cz, r2 * r2};
    const long long n = (long long)f.nx * f.ny * f.nz;
    depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_vox, W, H, D, f);
    CUDA_TRY(cudaGetLastError());
    // the carved cells lie inside the repaired box: traversal words around it, same stream, no synchronisation
    return trav_sync(c, f.x0, f.y0, f.z0, f.x0 + f.nx, f.y0 + f.ny, f.z0 + f.nz, false, false);
}

// removeSphere whose 16-byte command {cx, cy, cz, radius} lives in DEVICE memory (e.g. the target of an NCCL broadcast queued on
// vxrt_stream()): queued behind it on the context's stream, the host never reads the command.  radius must be in [0, max_radius].
extern "C" int vxrt_edit_remove_sphere_cmd(vxrt_ctx* c, const int32_t* device_cmd, int max_radius) {
    CHECK_CTX(c);
    if (!device_cmd) return fail(VXRT_ERR_INVALID, "remove_sphere_cmd: null command");
    if (max_radius < 0 || max_radius > 64) return fail(VXRT_ERR_INVALID, "remove_sphere_cmd: max_radius must be in [0, 64]");
    if (!c->grid_loaded) return fail(VXRT_ERR_STATE, "remove_sphere_cmd before any grid upload");
    if (!c->d_yrange) CUDA_TRY(cudaMalloc(&c->d_yrange, 4 * sizeof(int)));
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    int* err = c->d_yrange + 3;
    if (!c->edit_err_zeroed) { CUDA_TRY(cudaMemsetAsync(err, 0, sizeof(int), c->stream)); c->edit_err_zeroed = true; }
    if (max_radius > 0) {
        const long long n = 8ll * max_radius * max_radius * max_radius;
        carve_cmd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_vox, W, H, D, device_cmd, max_radius, err);
        CUDA_TRY(cudaGetLastError());
    }
    const int r2 = max_radius + (7 >> 1);
    const long long nd = 8ll * r2 * r2 * r2;
    depth_cmd_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, c->stream>>>(c->d_vox, W, H, D, device_cmd, max_radius, err);
    CUDA_TRY(cudaGetLastError());
    const long long nt = 4ll * (r2 + TRAV_REACH) * (r2 + TRAV_REACH) * (2 * r2 + 1);
    trav_cmd_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(c->d_vox, c->d_trav, W, H, D, device_cmd, max_radius, c->d_trav_bad);
    CUDA_TRY(cudaGetLastError());
    return VXRT_OK;
}

// 1 if a device-side edit command was refused since the context was created (radius outside [0, max_radius]); synchronises
extern "C" int vxrt_edit_cmd_error(vxrt_ctx* c) {
    CHECK_CTX(c);
    if (!c->d_yrange || !c->edit_err_zeroed) return 0;
    int e = 0;
    CUDA_TRY(cudaMemcpyAsync(&e, c->d_yrange + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return e;
}

extern "C" int vxrt_build_depth_field(vxrt_ctx* c) {                                     // render.cpp:273-286
    CHECK_CTX(c);
    if (!c->grid_loaded) return fail(VXRT_ERR_STATE, "build_depth_field before any grid upload");
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    EditBox b{0, 0, 0, W, H, D, 0, 0, 0, -1};
    const long long n = (long long)W * H * D;
    depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_vox, W, H, D, b);
    CUDA_TRY(cudaGetLastError());
    return trav_sync(c, 0, 0, 0, 0, 0, 0, true, true);               // (synchronises)
}

// ---- procedural levels ---------------------------------------------------------------------------
static int launch_trees(vxrt_ctx* c, int base_y, const int* surface, int relative_destroy) {
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    const int ntx = (W - 11) / 30, ntz = (D - 11) / 25;             // trees at x = 30..,  z = 25.. below w-10 / d-10
    if (ntx > 0 && ntz > 0) {
        trees_kernel<<<ntx * ntz, 128, 0, c->stream>>>(c->d_vox, W, H, D, ntx, base_y, surface, relative_destroy);
        CUDA_TRY(cudaGetLastError());
    }
    return VXRT_OK;
}

// render.cpp:349-352 (fill -1) + initVoxels() level.cpp:82-138, generalised from 512x96x512 to the context's grid
extern "C" int vxrt_generate_default_level(vxrt_ctx* c) {
    CHECK_CTX(c);
    const long long n = (long long)c->nvox;
    default_slabs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_vox, c->cfg.grid_w, c->cfg.grid_h, c->cfg.grid_d);
    CUDA_TRY(cudaGetLastError());
    int rc = launch_trees(c, 36, nullptr, 1);
    if (rc != VXRT_OK) return rc;
    rc = trav_sync(c, 0, 0, 0, 0, 0, 0, true, true);                  // (synchronises)
    if (rc != VXRT_OK) return rc;
    c->grid_loaded = true;
    return update_yrange(c, 0, c->nvox, true);
}

// config C4: integer fbm height field, the reference's material bands relative to the surface,
// the reference's trees standing on the surface.  No depth field (call vxrt_build_depth_field next).
extern "C" int vxrt_generate_terrain(vxrt_ctx* c, uint64_t seed) {
    CHECK_CTX(c);
    const int W = c->cfg.grid_w, H = c->cfg.grid_h, D = c->cfg.grid_d;
    int* d_surface = nullptr;
    CUDA_TRY(cudaMalloc(&d_surface, (size_t)W * D * sizeof(int)));
    terrain_surface_kernel<<<(W * D + 255) / 256, 256, 0, c->stream>>>(d_surface, W, H, D, seed);
    const long long n = (long long)c->nvox;
    terrain_slabs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_vox, d_surface, W, H, D);
    cudaError_t e = cudaGetLastError();
    int rc = (e == cudaSuccess) ? launch_trees(c, -1, d_surface, 0) : fail(VXRT_ERR_CUDA, cudaGetErrorString(e));
    cudaStreamSynchronize(c->stream);
    cudaFree(d_surface);
    if (rc != VXRT_OK) return rc;
    rc = trav_sync(c, 0, 0, 0, 0, 0, 0, true, true);
    if (rc != VXRT_OK) return rc;
    c->grid_loaded = true;
    return update_yrange(c, 0, c->nvox, true);
}

extern "C" int vxrt_terrain_height(uint64_t seed, int x, int z, int grid_h) { return terrain_height(seed, x, z, grid_h); }

// ---- frame -------------------------------------------------------------------------------------
extern "C" int vxrt_set_frame(vxrt_ctx* c, const vxrt_frame* f) {                        // render.cpp:289-296
    if (!c || !f) return fail(VXRT_ERR_INVALID, "null argument");
    c->frame = *f;
    return VXRT_OK;
}

extern "C" int vxrt_get_frame(vxrt_ctx* c, vxrt_frame* out) {
    if (!c || !out) return fail(VXRT_ERR_INVALID, "null argument");
    *out = c->frame;
    return VXRT_OK;
}

extern "C" int vxrt_init_local_lights(vxrt_ctx* c) {                                     // render.cpp:304-311
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    for (int i = 0; i < VXRT_MAX_LOCAL_LIGHTS; i++) {
        c->frame.lights[i][0] = -1.0f; c->frame.lights[i][1] = -1.0f; c->frame.lights[i][2] = -1.0f; c->frame.lights[i][3] = 0.0f;
    }
    return VXRT_OK;
}

extern "C" int vxrt_place_local_light(vxrt_ctx* c, float x, float y, float z, float diffuse) {   // render.cpp:375-385
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    for (int i = 0; i < VXRT_MAX_LOCAL_LIGHTS; i++) {
        float* L = c->frame.lights[i];
        if (L[0] < 0 || L[1] < 0 || L[2] < 0) { L[0] = x; L[1] = y; L[2] = z; L[3] = diffuse; return i; }
    }
    return VXRT_MAX_LOCAL_LIGHTS;
}

extern "C" int vxrt_resize(vxrt_ctx* c, int width, int height) {                         // render.cpp:404-411
    CHECK_CTX(c);
    if (width <= 0 || height <= 0) return fail(VXRT_ERR_INVALID, "resize: extents must be positive");
    if (c->p2p) return fail(VXRT_ERR_STATE, "resize: not available once a peer-memory target is set");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));      // queued read-backs still read the buffers freed below
    c->cfg.width = width; c->cfg.height = height;
    c->band_ms_n = 0;
    c->frame.aspect = (float)width / height;
    return alloc_frame_buffers(c);
}

// Launches the frame's kernels in `nbands` bands of whole tile rows.  host_dst != nullptr: each band's pixels are
// copied to host_dst (page-locked) on the copy stream as soon as the band's kernels finish, so the read-back of
// band b overlaps the rendering of band b+1.
// ---- launch orders (slowest block first), refreshed OFF the frame's stream --------------------------------------------------
// The one-block counting sorts take 37 us (32,400 tiles) + 70 us (64,800 shade units) for a whole 4K frame.  The order they
// produce is a hint for LATER frames, so they run on a side stream, behind the frame that produced the block times, into the
// back buffers; a later render call adopts the result once the sort's event has completed (pointer swap, no synchronisation).
static void adopt_launch_orders(vxrt_ctx* c) {
    if (!c->order_pending) return;
    const cudaError_t e = cudaEventQuery(c->ev_order_done);
    if (e == cudaErrorNotReady) { cudaGetLastError(); return; }
    c->order_pending = false;
    if (e != cudaSuccess) { cudaGetLastError(); return; }
    if (c->pend_tiles) { std::swap(c->d_tile_order, c->d_tile_order_back); c->have_tile_order = true; }
    if (c->pend_shade) { std::swap(c->d_shade_order, c->d_shade_order_back); c->have_shade_order = true; c->order_shade_threads = c->pend_shade_threads; }
}
// forget a sort in flight (the block times it reads are about to mean something else, or its buffers are about to go)
static void drop_pending_orders(vxrt_ctx* c) {
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaGetLastError(); }
    c->order_pending = false;
}
static int refresh_launch_orders(vxrt_ctx* c, bool tiles, bool shade) {
    if (c->order_pending) return VXRT_OK;                 // the previous refresh has not been adopted yet
    // The sort reads every block time twice (histogram pass, scatter pass) while the NEXT frame's kernels already overwrite them:
    // it must work on a snapshot, or the two passes disagree and the "order" is no permutation (tiles rendered twice / never).
    // The snapshots are small device-to-device copies on the frame's own stream (130 KB + 260 KB for a whole 4K frame).
    const int upt = TILE_PIX / c->shade_threads;
    if (tiles) CUDA_TRY(cudaMemcpyAsync(c->d_cost_snap, c->d_tile_cost, (size_t)c->map.nlocal * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (shade) CUDA_TRY(cudaMemcpyAsync(c->d_cost_snap + c->map.nlocal, c->d_shade_cost, (size_t)c->map.nlocal * upt * 4, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev_order_src, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_order_src, 0));
    if (tiles) {
        tile_order_kernel<<<1, 1024, 0, c->aux_stream>>>(c->d_cost_snap, c->d_tile_order_back, c->map.nlocal);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
    }
    if (shade) {
        tile_order_kernel<<<1, 1024, 0, c->aux_stream>>>(c->d_cost_snap + c->map.nlocal, c->d_shade_order_back, c->map.nlocal * upt);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
    }
    CUDA_TRY(cudaEventRecord(c->ev_order_done, c->aux_stream));
    c->order_pending = true; c->pend_tiles = tiles; c->pend_shade = shade; c->pend_shade_threads = c->shade_threads;
    return VXRT_OK;
}

static int render_bands(vxrt_ctx* c, int nbands, uint8_t* host_dst, uint32_t* dev_out = nullptr, bool fence_main = true, bool raster_out = false,
                        bool local_out = false) {
    const bool p2p_frame = c->p2p && !raster_out && !local_out;   // a host-frame render of a peer-memory context leaves the peer frame alone
    if (!dev_out) dev_out = c->d_rgba8;
    if (fence_main)                                       // synchronous paths: never overwrite a frame a pipelined copy still reads
        for (int slot = 0; slot < 2; slot++)
            if (c->slot_busy[slot]) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_slot[slot], 0));
    if (!c->grid_loaded) return fail(VXRT_ERR_STATE, "render before any grid upload");
    adopt_launch_orders(c);
    const GridView g = grid_view(c);
    FrameParams fp;
    memcpy(&fp, &c->frame, sizeof fp);
    // kernel variants: iteration counting on/off (the step-count view and the debug planes always need it), and the
    // reference's compile-time grid extents (512 x 96 x 512) vs runtime extents
    const bool count = c->stats_mode != 0 || c->d_dbg_hit != nullptr;
    const bool counted_full = c->stats_mode == 1 || (c->stats_mode == 0 && c->d_dbg_hit != nullptr);
    const bool count_primary = count || c->frame.view_depth_field == 1;
    const bool ref_dims = (g.w == GridViewRef::w && g.h == GridViewRef::h && g.d == GridViewRef::d);
    GridViewRef gr; gr.vox = g.vox; gr.ymin = g.ymin; gr.ymax = g.ymax;
    // the two passes may read different copies of the grid (vxrt_set_traversal, auto): the traversal grid or the reference-layout one
    const bool trav_shade = use_trav(c), trav_primary = use_trav_primary(c);
    GridView g_plain = g, g_trav = g;
    GridViewRef gr_plain = gr, gr_trav = gr;
    g_plain.vox = c->d_vox; gr_plain.vox = c->d_vox; g_trav.vox = c->d_trav; gr_trav.vox = c->d_trav;
    // bands: whole tile rows when this context owns the whole frame (raster rows stay contiguous), else tile ranges
    const int units = (c->cfg.world == 1) ? c->map.ty : c->map.nlocal;
    const int tiles_per_unit = (c->cfg.world == 1) ? c->map.tx : 1;
    if (c->map.rows) nbands = 1;                          // (a rank's strips hold whole tile rows: no contiguous tile ranges to band over)
    if (nbands > units) nbands = units;
    if (nbands > MAX_BANDS) nbands = MAX_BANDS;
    if (nbands < 1) nbands = 1;
    c->launches = 0;
    c->bands_used = nbands;
    // Peer-memory frames: the flag protocol rides on programmatic dependent launches, so that none of its kernels costs a launch
    // gap on the frame's critical path -- p2p_begin_kernel (back-pressure wait + counter reset) runs WHILE the first render kernel
    // traces its primary rays (that kernel executes griddepcontrol.wait before its first global write), and the kernel that
    // publishes this rank's completion flag is resident before the last render kernel ends.  Nothing may be queued between the
    // members of such a pair: the frame's events bracket the whole chain and the launch-order refreshes follow it.
    const bool pdl_chain = p2p_frame && c->p2p_pdl;
    if (p2p_frame && host_dst) return fail(VXRT_ERR_STATE, "render_frame_host is not available on a peer-memory context: use vxrt_p2p_wait_frame on the owner");
    if (pdl_chain) {
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        p2p_begin_kernel<<<1, 64, 0, c->stream>>>((const P2PShared*)c->p2p_base, c->p2p_seq, c->d_p2p_err, (unsigned int*)c->d_counters,
                                                  (int)(sizeof(Counters) * MAX_BANDS / sizeof(unsigned int)), c->d_chain_time);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
    } else {
        CUDA_TRY(cudaMemsetAsync(c->d_counters, 0, sizeof(Counters) * MAX_BANDS, c->stream));
        if (p2p_frame) {
            p2p_wait_consumed_kernel<<<1, 1, 0, c->stream>>>((const P2PShared*)c->p2p_base, c->p2p_seq, c->d_p2p_err);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
        }
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    }
    bool pdl_first = pdl_chain && c->p2p_begin_overlap;    // the next render kernel is the first of the frame and may overlap the wait
    bool refresh_tiles_later = false, refresh_shade_later = false, ev1_recorded = false;
    // auto: small shares whose primary pass reads the reference-layout grid (with the traversal grid its long rays no longer wait
    // for a new line on every step, and the sweep costs more than it returns: 1/8 of a 4K frame 0.190 -> 0.184 ms without it)
    const bool sweep = c->l2_prefetch == 1 || (c->l2_prefetch == 2 && c->map.nlocal <= 12000 && !use_trav_primary(c));
    if (sweep && (size_t)c->nvox * 4 <= (size_t)120 << 20) {              // only grids that fit the 126 MB L2
        // rows a ray can read: up to 7 above the highest solid row (beyond that the culling ends the ray), all rows otherwise
        int ytop = c->cfg.grid_h;
        if (c->use_culling && !counted_full && c->yrange[1] >= c->yrange[0]) ytop = std::min(c->cfg.grid_h, c->yrange[1] + 9);
        const int slab_ints = c->cfg.grid_w * c->cfg.grid_h;
        const int lines_per_slab = (c->cfg.grid_w * ytop + 31) / 32;
        const long long nlines = (long long)lines_per_slab * c->cfg.grid_d;
        l2_prefetch_kernel<<<148 * 8, 256, 0, c->stream>>>(trav_primary ? c->d_trav : c->d_vox, lines_per_slab, slab_ints, nlines, c->d_yrange ? c->d_yrange + 2 : nullptr);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        pdl_first = false;                                // (the sweep, a plain launch, already waited for p2p_begin_kernel)
    }
    // Banded read-back: the copy of a band overlaps the kernels of the bands after it, so the frame is in host memory
    // soonest when the cheapest bands render first (their copies hide behind the expensive ones) -- the order of the
    // previous banded frame's kernel times; the sky half of a frame costs a twentieth of the ground half.
    int band_seq[MAX_BANDS];
    for (int b = 0; b < nbands; b++) band_seq[b] = b;
    const bool timed_bands = host_dst != nullptr && nbands > 1;
    if (timed_bands && c->band_ms_n == nbands)
        std::stable_sort(band_seq, band_seq + nbands, [&](int a, int b) { return c->band_ms[a] < c->band_ms[b]; });
    for (int bi = 0; bi < nbands; bi++) {
        const int b = band_seq[bi];
        if (timed_bands) CUDA_TRY(cudaEventRecord(c->ev_band_start[b], c->stream));
        const int u0 = (int)((long long)units * b / nbands), u1 = (int)((long long)units * (b + 1) / nbands);
        const int tile0 = u0 * tiles_per_unit, ntile = (u1 - u0) * tiles_per_unit;
        if (ntile <= 0) continue;
        TileMap m = c->map;
        m.tile_base = tile0;
        Outputs o;
        o.rgba8 = p2p_frame ? (uint32_t*)(c->p2p_base + sizeof(P2PShared) + (c->p2p_seq & 1) * c->p2p_frame_bytes) : dev_out;
        o.raster = (c->cfg.world == 1 || p2p_frame || raster_out) ? 1 : 0;
        o.skip_dark = (c->use_culling && !counted_full) ? 1 : 0;
        o.hitq = c->d_hitq; o.hitpix = c->d_hitpix; o.tile_hits = c->d_tile_hits;
        o.counters = c->d_counters + b;
        // longest-tile-first launch order (whole-frame launches only; bands keep their contiguous tile ranges)
        o.tile_order = (nbands == 1 && c->use_tile_order && c->have_tile_order) ? c->d_tile_order : nullptr;
        o.tile_cost = (nbands == 1 && c->use_tile_order) ? c->d_tile_cost : nullptr;
        const int upt = TILE_PIX / c->shade_threads;                  // shade units per tile
        if (c->order_shade_threads != c->shade_threads) c->have_shade_order = false;
        o.shade_order = (nbands == 1 && c->use_tile_order && c->have_shade_order) ? c->d_shade_order : nullptr;
        o.shade_cost = (nbands == 1 && c->use_tile_order) ? c->d_shade_cost : nullptr;
        o.shade_unit_base = tile0 * upt;
        o.dbg_hit = c->d_dbg_hit; o.dbg_steps = c->d_dbg_steps; o.dbg_occl = c->d_dbg_occl; o.dbg_cast = c->d_dbg_cast;
        const dim3 grid(ntile), block(256);
        // kernel variants: <iteration counters, grid type (reference extents as compile-time constants / runtime extents), traversal grid>;
        // PDL: launched with programmatic stream serialization (the shade pass of an overlapped frame)
#define VXRT_LAUNCH1(KERNEL, GRID, BLOCK, PDL, GV)                                                                            \
        do {                                                                                                                  \
            cudaLaunchConfig_t lc;                                                                                            \
            memset(&lc, 0, sizeof lc);                                                                                        \
            lc.gridDim = GRID; lc.blockDim = BLOCK; lc.dynamicSmemBytes = 0; lc.stream = c->stream;                            \
            cudaLaunchAttribute la[1];                                                                                        \
            la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                    \
            la[0].val.programmaticStreamSerializationAllowed = 1;                                                             \
            lc.attrs = la; lc.numAttrs = (PDL) ? 1 : 0;                                                                       \
            CUDA_TRY(cudaLaunchKernelEx(&lc, KERNEL, GV, fp, m, o));                                                          \
        } while (0)
#define VXRT_LAUNCH(KERNEL, COUNT, TRAV, GRID, BLOCK, PDL)                                                                    \
        do {                                                                                                                  \
            if (ref_dims) {                                                                                                   \
                if (COUNT) { if (TRAV) VXRT_LAUNCH1((KERNEL<true, GridViewRef, true>), GRID, BLOCK, PDL, gr_trav);            \
                             else VXRT_LAUNCH1((KERNEL<true, GridViewRef, false>), GRID, BLOCK, PDL, gr_plain); }             \
                else       { if (TRAV) VXRT_LAUNCH1((KERNEL<false, GridViewRef, true>), GRID, BLOCK, PDL, gr_trav);           \
                             else VXRT_LAUNCH1((KERNEL<false, GridViewRef, false>), GRID, BLOCK, PDL, gr_plain); }            \
            } else {                                                                                                          \
                if (COUNT) { if (TRAV) VXRT_LAUNCH1((KERNEL<true, GridView, true>), GRID, BLOCK, PDL, g_trav);                \
                             else VXRT_LAUNCH1((KERNEL<true, GridView, false>), GRID, BLOCK, PDL, g_plain); }                 \
                else       { if (TRAV) VXRT_LAUNCH1((KERNEL<false, GridView, true>), GRID, BLOCK, PDL, g_trav);               \
                             else VXRT_LAUNCH1((KERNEL<false, GridView, false>), GRID, BLOCK, PDL, g_plain); }                \
            }                                                                                                                 \
        } while (0)
        // overlap (whole-frame launches of the lit view only): nothing may be queued between the two passes, so the event that
        // separates their times and the refresh of the primary launch order move behind the shade pass
        const bool fused = nbands == 1 && (c->fusion == 1 || (c->fusion == 2 && c->map.nlocal <= 12000));
        const bool overlap = !fused && nbands == 1 && c->frame.view_depth_field != 1 &&
                             (c->overlap == 1 || (c->overlap == 2 && c->map.nlocal <= 12000));
        o.overlap = overlap ? 1 : 0;
        o.tile_ready = c->d_tile_ready; o.overlap_err = c->d_overlap_err;
        o.frame_seq = ++c->frame_seq;
        c->last_fused = fused;
        o.pdl_wait = 0; o.pdl_trigger = 0; o.wide_blocks = 0;
        const bool shade_follows = !fused && c->frame.view_depth_field != 1;
        if (fused) {
            // one kernel per frame: each block traces its tile's primary rays and then shades its own hits (kernels.cuh frame_kernel)
            o.pdl_wait = pdl_first ? 1 : 0; o.pdl_trigger = pdl_chain ? 1 : 0;
            // wide blocks: the heaviest tiles (the head of the launch order) are rendered by two blocks each, two threads per hit
            // pixel (kernels.cuh shade_wide_*); production variants of the lit view only
            int nwide = 0;
            if (!count_primary && o.tile_order && c->wide_tiles > 0) nwide = std::min(c->wide_tiles, ntile / 8);
            o.wide_blocks = 2 * nwide;
            const dim3 fgrid((unsigned)(ntile + nwide));
            VXRT_LAUNCH(frame_kernel, count_primary, trav_primary, fgrid, block, pdl_first);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
            pdl_first = false;
            if (!pdl_chain) CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));   // (one kernel: ms_primary reads as the whole frame, ms_shadow as 0)
            refresh_tiles_later = o.tile_cost && c->map.nlocal >= 64 && (c->order_frame < 2 || (c->order_frame % 8) == 0);
        } else {
        o.pdl_wait = pdl_first ? 1 : 0; o.pdl_trigger = (pdl_chain && !shade_follows) ? 1 : 0;
        VXRT_LAUNCH(primary_kernel, count_primary, trav_primary, grid, block, pdl_first);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        pdl_first = false;
        o.pdl_wait = 0; o.pdl_trigger = (pdl_chain && shade_follows) ? 1 : 0;
        if (nbands == 1 && !overlap && !(pdl_chain && !shade_follows)) { CUDA_TRY(cudaEventRecord(c->ev[1], c->stream)); ev1_recorded = true; }
        // the launch orders are refreshed on the first two frames and then every 8th (block times are temporally coherent), behind
        // the frame and off its stream (refresh_launch_orders below)
        const bool refresh_order = c->order_frame < 2 || (c->order_frame % 8) == 0;
        refresh_tiles_later = o.tile_cost && c->map.nlocal >= 64 && refresh_order;
        if (c->frame.view_depth_field != 1) {
            const dim3 sblock(c->shade_threads), sgrid((unsigned)(((size_t)ntile * TILE_PIX + c->shade_threads - 1) / c->shade_threads));
            VXRT_LAUNCH(shade_kernel, count, trav_shade, sgrid, sblock, overlap);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
            if (overlap) {
                CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));      // (the passes overlap: ms_primary then reads as the whole frame, ms_shadow as 0)
                ev1_recorded = true;
            }
            refresh_shade_later = o.shade_cost && c->map.nlocal >= 64 && refresh_order;
        }
        }   // two passes
        if (host_dst) {
            size_t off, bytes;
            if (c->cfg.world == 1) {
                const int y0 = u0 * TILE_H, y1 = (u1 * TILE_H < c->cfg.height) ? u1 * TILE_H : c->cfg.height;
                off = (size_t)y0 * c->cfg.width * 4; bytes = (size_t)(y1 - y0) * c->cfg.width * 4;
            } else {
                off = (size_t)tile0 * TILE_PIX * 4; bytes = (size_t)ntile * TILE_PIX * 4;
            }
            CUDA_TRY(cudaEventRecord(c->ev_band[b], c->stream));
            CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_band[b], 0));
            CUDA_TRY(cudaMemcpyAsync(host_dst + off, (const uint8_t*)dev_out + off, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    if (pdl_chain) {
        // the completion flag directly behind the last render kernel (resident before that kernel ends), then the events and the
        // launch-order refreshes that had to stay out of the chain
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3(1); lc.blockDim = dim3(1); lc.stream = c->stream;
        cudaLaunchAttribute la[1];
        la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        la[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = la; lc.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&lc, p2p_signal_done_kernel, (P2PShared*)c->p2p_base, (int)c->cfg.rank, c->p2p_seq, c->d_chain_time));
        c->launches++;
        c->p2p_seq++;
        // no event records behind the chain: the owner's wait / release kernels (vxrt_p2p_wait_frame / _release_frame) chain on
        // programmatically as well; vxrt_get_stats takes the chain's span from the device clock (d_chain_time)
        c->chain_timed = true; c->chain_has_ev1 = ev1_recorded;
        if (nbands == 1 && c->use_tile_order) c->order_frame++;
    } else {
    c->chain_timed = false;
    if (nbands != 1) CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    if (nbands == 1 && c->use_tile_order) c->order_frame++;
    if (p2p_frame) {
        p2p_signal_done_kernel<<<1, 1, 0, c->stream>>>((P2PShared*)c->p2p_base, c->cfg.rank, c->p2p_seq, (unsigned long long*)nullptr);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        c->p2p_seq++;
    }
    }
    if (refresh_tiles_later || refresh_shade_later) {
        const int rc = refresh_launch_orders(c, refresh_tiles_later, refresh_shade_later);
        if (rc != VXRT_OK) return rc;
    }
    if (host_dst && fence_main) {                         // later work on the main stream must not overwrite pixels in flight
        CUDA_TRY(cudaEventRecord(c->ev_copy, c->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    }
    c->rendered = true;
    c->d_last_frame = (p2p_frame || raster_out) ? nullptr : dev_out;    // (frames stored into a peer / host frame are not kept here)
    return VXRT_OK;
}

extern "C" int vxrt_render(vxrt_ctx* c) {                                                // main.cpp:59
    CHECK_CTX(c);
    return render_bands(c, 1, nullptr);
}

extern "C" int vxrt_set_readback_bands(vxrt_ctx* c, int nbands) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    if (nbands < 1 || nbands > MAX_BANDS) return fail(VXRT_ERR_INVALID, "readback bands must be in [1,16]");
    c->readback_bands = nbands;
    return VXRT_OK;
}

extern "C" int vxrt_set_l2_prefetch(vxrt_ctx* c, int enabled) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    if (enabled < 0 || enabled > 2) return fail(VXRT_ERR_INVALID, "l2_prefetch: 0 off, 1 on, 2 auto");
    c->l2_prefetch = enabled;
    return VXRT_OK;
}

extern "C" int vxrt_set_culling(vxrt_ctx* c, int enabled) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    c->use_culling = enabled != 0;
    return VXRT_OK;
}

extern "C" int vxrt_set_traversal(vxrt_ctx* c, int mode) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    if (mode < 0 || mode > 2) return fail(VXRT_ERR_INVALID, "set_traversal: 0 off, 1 both passes, 2 auto");
    c->trav_mode = mode;
    return VXRT_OK;
}

// 1: the rays read the traversal grid; 0: switched off, or the grid holds values that cannot be encoded (then the plain kernels
// render from the reference-layout grid: same pixels)
extern "C" int vxrt_traversal_active(vxrt_ctx* c) { return (c && use_trav(c)) ? 1 : 0; }

extern "C" int vxrt_download_traversal(vxrt_ctx* c, int32_t* out, size_t count) {
    CHECK_CTX(c);
    if (!out || count != c->nvox) return fail(VXRT_ERR_INVALID, "download_traversal: count must equal grid_w*grid_h*grid_d");
    CUDA_TRY(cudaMemcpyAsync(out, c->d_trav, count * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

extern "C" int vxrt_set_fusion(vxrt_ctx* c, int mode) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    if (mode < 0 || mode > 2) return fail(VXRT_ERR_INVALID, "set_fusion: 0 two passes, 1 fused, 2 auto");
    if (mode != c->fusion) { drop_pending_orders(c); c->have_tile_order = false; c->have_shade_order = false; c->order_frame = 0; }   // block times mean something else now
    c->fusion = mode;
    return VXRT_OK;
}

extern "C" int vxrt_set_partition(vxrt_ctx* c, int mode) {
    CHECK_CTX(c);
    if (mode != 0 && mode != 1) return fail(VXRT_ERR_INVALID, "set_partition: 0 tiles, 1 tile rows");
    if ((mode == 1) == c->row_partition) return VXRT_OK;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    c->row_partition = mode == 1;
    return alloc_frame_buffers(c);       // other local tiles: hit slots, launch orders and the local frame start over
}

extern "C" int vxrt_set_wide_tiles(vxrt_ctx* c, int tiles) {
    CHECK_CTX(c);
    if (tiles < 0 || tiles > 64) return fail(VXRT_ERR_INVALID, "set_wide_tiles: 0 (off) .. 64");
    c->wide_tiles = tiles;
    return VXRT_OK;
}

// host-only: the partition's arithmetic (kernels.cuh tile_of / tile_owner) for tests and integrators; needs no device
extern "C" int vxrt_partition_tile(int width, int height, int rank, int world, int tile_rows, int local_tile) {
    if (width <= 0 || height <= 0 || world < 1 || rank < 0 || rank >= world || local_tile < 0) return -1;
    const TileMap m = make_map(width, height, rank, world, tile_rows != 0);
    if (local_tile >= m.nlocal) return -1;
    const int t = tile_of(m, local_tile);
    return t < m.ntiles ? t : -1;
}
extern "C" int vxrt_partition_owner(int width, int height, int world, int tile_rows, int tile) {
    if (width <= 0 || height <= 0 || world < 1 || tile < 0) return -1;
    const TileMap m = make_map(width, height, 0, world, tile_rows != 0);
    if (tile >= m.ntiles) return -1;
    return m.rows ? (tile / m.tx) % m.world : tile_owner(m, tile);
}

extern "C" int vxrt_frame_was_fused(vxrt_ctx* c) { return (c && c->last_fused) ? 1 : 0; }

extern "C" int vxrt_set_overlap(vxrt_ctx* c, int mode) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    if (mode < 0 || mode > 2) return fail(VXRT_ERR_INVALID, "set_overlap: 0 off, 1 on, 2 auto");
    c->overlap = mode;
    return VXRT_OK;
}

extern "C" int vxrt_set_tile_ordering(vxrt_ctx* c, int enabled) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    c->use_tile_order = enabled != 0;
    drop_pending_orders(c);
    if (!enabled) { c->have_tile_order = false; c->have_shade_order = false; }
    c->order_frame = 0;
    return VXRT_OK;
}

extern "C" int vxrt_set_stats(vxrt_ctx* c, int enabled) {
    if (!c) return fail(VXRT_ERR_INVALID, "null context");
    if (enabled < 0 || enabled > 2) return fail(VXRT_ERR_INVALID, "set_stats: 0 off, 1 reference casting rule, 2 executed work");
    c->stats_mode = enabled;
    return VXRT_OK;
}

extern "C" int vxrt_sync(vxrt_ctx* c) {
    CHECK_CTX(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

extern "C" int vxrt_render_frame_host(vxrt_ctx* c, const vxrt_frame* f, uint8_t* out) {
    if (!out) return fail(VXRT_ERR_INVALID, "render_frame_host: null output");
    int rc = vxrt_set_frame(c, f);
    if (rc != VXRT_OK) return rc;
    CHECK_CTX(c);
    // page-locked destination (vxrt_host_alloc / cudaHostRegister): DMA straight into it; pageable: stage through the
    // context's pinned buffer.  Either way the copy of band b overlaps the kernels of band b+1.
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    rc = render_bands(c, c->readback_bands, pinned ? out : c->h_frame);
    if (rc != VXRT_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->bands_used > 1) {                               // this frame's band times order the next frame's bands
        for (int b = 0; b < c->bands_used; b++)
            if (cudaEventElapsedTime(&c->band_ms[b], c->ev_band_start[b], c->ev_band[b]) != cudaSuccess) { cudaGetLastError(); c->band_ms[b] = 0.0f; }
        c->band_ms_n = c->bands_used;
    }
    if (!pinned) memcpy(out, c->h_frame, c->out_pixels * 4);
    return VXRT_OK;
}

// Pipelined variant: returns as soon as the work is queued.  Frames alternate between two device buffers, so the
// read-back of frame k (copy stream) overlaps the kernels of frame k+1 (main stream); at most two frames are in
// flight -- a third submit first waits for the oldest one's copy.  `out` must be page-locked (vxrt_host_alloc).
extern "C" int vxrt_submit_frame_host(vxrt_ctx* c, const vxrt_frame* f, uint8_t* out) {
    if (!out) return fail(VXRT_ERR_INVALID, "submit_frame_host: null output");
    int rc = vxrt_set_frame(c, f);
    if (rc != VXRT_OK) return rc;
    CHECK_CTX(c);
    if (c->p2p) return fail(VXRT_ERR_STATE, "submit_frame_host is not available on a peer-memory context");
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) return fail(VXRT_ERR_INVALID, "submit_frame_host needs a page-locked destination (vxrt_host_alloc)");
    if (!c->d_rgba8_alt) CUDA_TRY(cudaMalloc(&c->d_rgba8_alt, c->out_pixels * 4));
    const int slot = (int)(c->submit_seq & 1);
    if (c->slot_busy[slot]) {                              // the device buffer of this slot is still being copied out
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_slot[slot], 0));
    }
    rc = render_bands(c, 1, out, slot ? c->d_rgba8_alt : c->d_rgba8, /*fence_main=*/false);
    if (rc != VXRT_OK) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_slot[slot], c->copy_stream));
    c->slot_busy[slot] = true;
    c->submit_seq++;
    return VXRT_OK;
}

// blocks until every submitted frame is in host memory
extern "C" int vxrt_wait_frames(vxrt_ctx* c) {
    CHECK_CTX(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    c->slot_busy[0] = c->slot_busy[1] = false;
    return VXRT_OK;
}

// ---- results -----------------------------------------------------------------------------------
extern "C" int vxrt_read_rgba8(vxrt_ctx* c, uint8_t* out) {
    CHECK_CTX(c);
    if (!out) return fail(VXRT_ERR_INVALID, "read_rgba8: null output");
    if (!c->rendered) return fail(VXRT_ERR_STATE, "read_rgba8 before render");
    if (!c->d_last_frame) return fail(VXRT_ERR_STATE, "read_rgba8: the last frame was stored into a peer-memory / host frame, not into this context's buffer");
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));       // a pipelined read-back of that buffer may still be in flight
    CUDA_TRY(cudaMemcpyAsync(out, c->d_last_frame, c->out_pixels * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

extern "C" int vxrt_read_debug(vxrt_ctx* c, int32_t* hit_index, uint16_t* steps, uint32_t* occl_mask, uint32_t* cast_mask) {
    CHECK_CTX(c);
    if (!(c->cfg.flags & VXRT_FLAG_DEBUG_OUTPUTS)) return fail(VXRT_ERR_STATE, "read_debug needs VXRT_FLAG_DEBUG_OUTPUTS");
    if (!c->rendered) return fail(VXRT_ERR_STATE, "read_debug before render");
    const size_t npix = (size_t)c->cfg.width * c->cfg.height;
    if (hit_index) CUDA_TRY(cudaMemcpyAsync(hit_index, c->d_dbg_hit, npix * 4, cudaMemcpyDeviceToHost, c->stream));
    if (steps) CUDA_TRY(cudaMemcpyAsync(steps, c->d_dbg_steps, npix * 2, cudaMemcpyDeviceToHost, c->stream));
    if (occl_mask) CUDA_TRY(cudaMemcpyAsync(occl_mask, c->d_dbg_occl, npix * 4, cudaMemcpyDeviceToHost, c->stream));
    if (cast_mask) CUDA_TRY(cudaMemcpyAsync(cast_mask, c->d_dbg_cast, npix * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

extern "C" int vxrt_get_stats(vxrt_ctx* c, vxrt_stats* out) {
    CHECK_CTX(c);
    if (!out) return fail(VXRT_ERR_INVALID, "get_stats: null output");
    if (!c->rendered) return fail(VXRT_ERR_STATE, "get_stats before render");
    Counters hb[MAX_BANDS], h;
    CUDA_TRY(cudaMemcpyAsync(hb, c->d_counters, sizeof hb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memset(&h, 0, sizeof h);
    for (int b = 0; b < MAX_BANDS; b++) {
        h.hit_count += hb[b].hit_count; h.global_traced += hb[b].global_traced; h.rays_local += hb[b].rays_local;
        h.fetches_primary += hb[b].fetches_primary; h.fetches_shadow += hb[b].fetches_shadow; h.rays_dark += hb[b].rays_dark;
    }
    memset(out, 0, sizeof *out);
    // pixels this context rendered (padding tiles and clipped pixels excluded)
    uint64_t pix = 0;
    for (int j = 0; j < c->map.nlocal; j++) {
        const int t = tile_of(c->map, j);
        if (t >= c->map.ntiles) break;
        const int x0 = (t % c->map.tx) * TILE_W, y0 = (t / c->map.tx) * TILE_H;
        const int w = (c->map.width - x0 < TILE_W) ? c->map.width - x0 : TILE_W;
        const int hgt = (c->map.height - y0 < TILE_H) ? c->map.height - y0 : TILE_H;
        pix += (uint64_t)w * hgt;
    }
    out->rays_primary = pix;
    out->hit_pixels = h.hit_count;
    out->rays_global = (c->frame.view_depth_field == 1) ? 0 : (c->stats_mode == 2 ? h.global_traced : h.hit_count);
    out->rays_local = h.rays_local;
    out->fetches = h.fetches_primary + h.fetches_shadow;
    out->fetches_primary = h.fetches_primary;
    out->rays_dark = h.rays_dark;
    if (c->chain_timed) {
        // a peer-memory frame whose kernels are chained programmatically: the chain's span from the device clock (globaltimer, ns)
        unsigned long long tt[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(tt, c->d_chain_time, sizeof tt, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        out->ms_total = (tt[1] > tt[0]) ? (float)((double)(tt[1] - tt[0]) * 1e-6) : 0.0f;
        if (c->chain_has_ev1) { CUDA_TRY(cudaEventElapsedTime(&out->ms_primary, c->ev[0], c->ev[1])); }   // two passes: an event sits between them
        else out->ms_primary = out->ms_total;
        if (out->ms_primary > out->ms_total) out->ms_primary = out->ms_total;
        out->ms_shadow = out->ms_total - out->ms_primary;
    } else {
        CUDA_TRY(cudaEventElapsedTime(&out->ms_primary, c->ev[0], c->ev[1]));
        CUDA_TRY(cudaEventElapsedTime(&out->ms_shadow, c->ev[1], c->ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&out->ms_total, c->ev[0], c->ev[2]));
    }
    out->kernel_launches = c->launches;
    return VXRT_OK;
}

// diagnostics: SM cycles each local tile's primary block / each shade unit's block took in the last whole-frame launch (the
// figures the slowest-first launch orders are made from); either pointer may be null.  Sizes: vxrt_local_tiles() and
// vxrt_local_tiles() * (256 / shade threads per block) entries
extern "C" int vxrt_read_block_costs(vxrt_ctx* c, uint32_t* primary, uint32_t* shade, size_t shade_count) {
    CHECK_CTX(c);
    if (!c->rendered) return fail(VXRT_ERR_STATE, "read_block_costs before render");
    const size_t upt = (size_t)(TILE_PIX / c->shade_threads);
    if (shade && shade_count != (size_t)c->map.nlocal * upt) return fail(VXRT_ERR_INVALID, "read_block_costs: shade_count must be local tiles * units per tile");
    if (primary) CUDA_TRY(cudaMemcpyAsync(primary, c->d_tile_cost, (size_t)c->map.nlocal * 4, cudaMemcpyDeviceToHost, c->stream));
    if (shade) CUDA_TRY(cudaMemcpyAsync(shade, c->d_shade_cost, shade_count * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return VXRT_OK;
}

extern "C" int vxrt_cast_rays(vxrt_ctx* c, int32_t n, const float* starts, const float* dirs, const int32_t* dists,
                              int32_t* ret, float* out7) {
    CHECK_CTX(c);
    if (n < 0 || (n > 0 && (!starts || !dirs || !dists || !ret || !out7))) return fail(VXRT_ERR_INVALID, "cast_rays: bad argument");
    if (!c->grid_loaded) return fail(VXRT_ERR_STATE, "cast_rays before any grid upload");
    if (n == 0) return VXRT_OK;
    float *d_s = nullptr, *d_d = nullptr, *d_o = nullptr; int32_t *d_n = nullptr, *d_r = nullptr;
    auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_d); cudaFree(d_o); cudaFree(d_n); cudaFree(d_r); };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&d_s, (size_t)n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_d, (size_t)n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, (size_t)n * 28);
    if (e == cudaSuccess) e = cudaMalloc(&d_n, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_r, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_s, starts, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, dirs, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, dists, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        if (use_trav(c)) cast_rays_kernel<true><<<(n + 127) / 128, 128, 0, c->stream>>>(grid_view(c), n, d_s, d_d, d_n, d_r, d_o);
        else cast_rays_kernel<false><<<(n + 127) / 128, 128, 0, c->stream>>>(grid_view(c), n, d_s, d_d, d_n, d_r, d_o);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ret, d_r, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out7, d_o, (size_t)n * 28, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) return fail(VXRT_ERR_CUDA, std::string("cast_rays: ") + cudaGetErrorString(e));
    return VXRT_OK;
}

extern "C" int vxrt_selftest_division(vxrt_ctx* c, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    CHECK_CTX(c);
    if (!mismatches) return fail(VXRT_ERR_INVALID, "selftest_division: null output");
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 8));
    cudaMemsetAsync(d, 0, 8, c->stream);
    division_selftest_kernel<<<148 * 8, 256, 0, c->stream>>>(n, seed, d);
    unsigned long long h = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(VXRT_ERR_CUDA, std::string("selftest_division: ") + cudaGetErrorString(e));
    *mismatches = h;
    return VXRT_OK;
}

extern "C" int vxrt_selftest_reciprocal(vxrt_ctx* c, uint64_t* mismatches) {
    CHECK_CTX(c);
    if (!mismatches) return fail(VXRT_ERR_INVALID, "selftest_reciprocal: null output");
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 8));
    cudaMemsetAsync(d, 0, 8, c->stream);
    reciprocal_selftest_kernel<<<148 * 8, 256, 0, c->stream>>>(d);
    unsigned long long h = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(VXRT_ERR_CUDA, std::string("selftest_reciprocal: ") + cudaGetErrorString(e));
    *mismatches = h;
    return VXRT_OK;
}

extern "C" int vxrt_write_ppm(vxrt_ctx* c, const char* path) {
    CHECK_CTX(c);
    if (!path) return fail(VXRT_ERR_INVALID, "write_ppm: null path");
    if (c->cfg.world != 1) return fail(VXRT_ERR_STATE, "write_ppm needs a whole-frame context (world == 1)");
    const int W = c->cfg.width, H = c->cfg.height;
    std::vector<uint8_t> rgba((size_t)W * H * 4);
    int rc = vxrt_read_rgba8(c, rgba.data());
    if (rc != VXRT_OK) return rc;
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(VXRT_ERR_IO, std::string("cannot open ") + path);
    fprintf(fp, "P6\n%d %d\n255\n", W, H);
    std::vector<uint8_t> row((size_t)W * 3);
    for (int y = H - 1; y >= 0; y--) {                   // GL rows are bottom-up
        for (int x = 0; x < W; x++) for (int k = 0; k < 3; k++) row[3 * x + k] = rgba[4 * ((size_t)y * W + x) + k];
        fwrite(row.data(), 1, row.size(), fp);
    }
    fclose(fp);
    return VXRT_OK;
}

// ---- multi-GPU plumbing --------------------------------------------------------------------------
extern "C" size_t vxrt_local_tiles(vxrt_ctx* c) { return c ? (size_t)c->map.nlocal : 0; }
extern "C" size_t vxrt_local_bytes(vxrt_ctx* c) { return c ? (size_t)c->map.nlocal * TILE_PIX * 4 : 0; }
extern "C" void* vxrt_device_rgba8(vxrt_ctx* c) { return c ? (void*)c->d_rgba8 : nullptr; }
extern "C" void* vxrt_stream(vxrt_ctx* c) { return c ? (void*)c->stream : nullptr; }

// ---- peer-memory frame target ----------------------------------------------------------------------
static unsigned long long pci_identity(int device) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) { cudaGetLastError(); return 0ull; }
    return 1ull + (((unsigned long long)pr.pciDomainID << 16) | ((unsigned long long)pr.pciBusID << 8) | (unsigned long long)pr.pciDeviceID);
}

extern "C" int vxrt_p2p_export(vxrt_ctx* c, uint8_t handle[64]) {
    CHECK_CTX(c);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!handle) return fail(VXRT_ERR_INVALID, "p2p_export: null handle");
    if (c->p2p) return fail(VXRT_ERR_STATE, "peer-memory target already set");
    if (c->cfg.world > 16) return fail(VXRT_ERR_INVALID, "peer-memory target supports up to 16 ranks");
    c->p2p_frame_bytes = (size_t)c->cfg.width * c->cfg.height * 4;
    const size_t total = sizeof(P2PShared) + 2 * c->p2p_frame_bytes;
    CUDA_TRY(cudaMalloc(&c->p2p_base, total));
    CUDA_TRY(cudaMemset(c->p2p_base, 0, total));
    {   // header: what an importer's context must look like
        const unsigned long long hdr[4] = {(unsigned long long)c->cfg.width, (unsigned long long)c->cfg.height, (unsigned long long)c->cfg.world,
                                           pci_identity(c->cfg.device)};
        CUDA_TRY(cudaMemcpy(c->p2p_base + offsetof(P2PShared, width), hdr, sizeof hdr, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMalloc(&c->d_p2p_err, sizeof(int)));
    CUDA_TRY(cudaMemset(c->d_p2p_err, 0, sizeof(int)));
    CUDA_TRY(cudaMalloc(&c->d_chain_time, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(c->d_chain_time, 0, 2 * sizeof(unsigned long long)));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->p2p_base));
    memcpy(handle, &h, 64);
    c->p2p = true; c->p2p_owner = true; c->p2p_seq = 0;
    return VXRT_OK;
}

// an importer stores into the owner's frame with ITS OWN extents and rank: both must be what the owner allocated for
static int p2p_check_header(vxrt_ctx* c, const void* base) {
    if (c->cfg.world > 16 || c->cfg.rank >= 16) return fail(VXRT_ERR_INVALID, "peer-memory target supports up to 16 ranks");
    unsigned long long hdr[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpy(hdr, (const uint8_t*)base + offsetof(P2PShared, width), sizeof hdr, cudaMemcpyDeviceToHost));
    // The back-pressure wait may run WHILE this context's first render kernel already occupies the GPU (programmatic dependent
    // launch) only if the owner's release cannot depend on this GPU: a kernel with undispatched blocks keeps the kernels of other
    // streams from being dispatched, so on the owner's own device the release would wait behind it (a deadlock the bounded spin
    // ends after 4 s).  Other device than the owner's: safe.
    const unsigned long long mine = pci_identity(c->cfg.device);
    c->p2p_begin_overlap = hdr[3] != 0 && mine != 0 && hdr[3] != mine;
    if (hdr[0] != (unsigned long long)c->cfg.width || hdr[1] != (unsigned long long)c->cfg.height || hdr[2] != (unsigned long long)c->cfg.world)
        return fail(VXRT_ERR_INVALID, "peer-memory target: the owner's frame is " + std::to_string(hdr[0]) + "x" + std::to_string(hdr[1]) + " for " +
                    std::to_string(hdr[2]) + " ranks, this context renders " + std::to_string(c->cfg.width) + "x" + std::to_string(c->cfg.height) +
                    " as one of " + std::to_string(c->cfg.world));
    return VXRT_OK;
}

extern "C" int vxrt_p2p_import(vxrt_ctx* c, const uint8_t handle[64]) {
    CHECK_CTX(c);
    if (!handle) return fail(VXRT_ERR_INVALID, "p2p_import: null handle");
    if (c->p2p) return fail(VXRT_ERR_STATE, "peer-memory target already set");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    { const int rc = p2p_check_header(c, p); if (rc != VXRT_OK) { cudaIpcCloseMemHandle(p); return rc; } }
    c->p2p_base = (uint8_t*)p;
    c->p2p_frame_bytes = (size_t)c->cfg.width * c->cfg.height * 4;
    CUDA_TRY(cudaMalloc(&c->d_p2p_err, sizeof(int)));
    CUDA_TRY(cudaMemset(c->d_p2p_err, 0, sizeof(int)));
    CUDA_TRY(cudaMalloc(&c->d_chain_time, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(c->d_chain_time, 0, 2 * sizeof(unsigned long long)));
    c->p2p = true; c->p2p_owner = false; c->p2p_seq = 0;
    return VXRT_OK;
}

// same-process variant of import (several contexts in one process, peer access enabled between their devices)
extern "C" int vxrt_p2p_attach(vxrt_ctx* c, void* owner_base) {
    CHECK_CTX(c);
    if (!owner_base) return fail(VXRT_ERR_INVALID, "p2p_attach: null base");
    if (c->p2p) return fail(VXRT_ERR_STATE, "peer-memory target already set");
    { const int rc = p2p_check_header(c, owner_base); if (rc != VXRT_OK) return rc; }
    c->p2p_base = (uint8_t*)owner_base;
    c->p2p_frame_bytes = (size_t)c->cfg.width * c->cfg.height * 4;
    CUDA_TRY(cudaMalloc(&c->d_p2p_err, sizeof(int)));
    CUDA_TRY(cudaMemset(c->d_p2p_err, 0, sizeof(int)));
    CUDA_TRY(cudaMalloc(&c->d_chain_time, 2 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(c->d_chain_time, 0, 2 * sizeof(unsigned long long)));
    c->p2p = true; c->p2p_owner = false; c->p2p_attached = true; c->p2p_seq = 0;
    return VXRT_OK;
}
extern "C" void* vxrt_p2p_base(vxrt_ctx* c) { return (c && c->p2p) ? (void*)c->p2p_base : nullptr; }

extern "C" int vxrt_p2p_wait_frame(vxrt_ctx* c, void** frame) {
    CHECK_CTX(c);
    if (!c->p2p || !c->p2p_owner) return fail(VXRT_ERR_STATE, "p2p_wait_frame: not the owner of a peer-memory target");
    if (c->p2p_seq == 0) return fail(VXRT_ERR_STATE, "p2p_wait_frame before render");
    const unsigned long long seq = c->p2p_seq - 1;
    {   // programmatic stream serialization: directly behind this rank's own p2p_signal_done_kernel the wait is resident (and already
        // polling the other ranks' flags) before the frame's kernels end; it completes only when every flag -- this rank's included --
        // is there, and the release queued behind it waits for that completion
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3(1); lc.blockDim = dim3(32); lc.stream = c->stream;
        cudaLaunchAttribute la[1];
        la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        la[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = la; lc.numAttrs = c->p2p_pdl ? 1 : 0;
        CUDA_TRY(cudaLaunchKernelEx(&lc, p2p_wait_done_kernel, (const P2PShared*)c->p2p_base, (int)c->cfg.world, seq, c->d_p2p_err));
    }
    if (frame) *frame = c->p2p_base + sizeof(P2PShared) + (seq & 1) * c->p2p_frame_bytes;
    return VXRT_OK;
}

extern "C" int vxrt_p2p_release_frame(vxrt_ctx* c) {
    CHECK_CTX(c);
    if (!c->p2p || !c->p2p_owner) return fail(VXRT_ERR_STATE, "p2p_release_frame: not the owner of a peer-memory target");
    if (c->p2p_seq == 0) return fail(VXRT_ERR_STATE, "p2p_release_frame before render");
    {   // programmatic stream serialization: behind p2p_wait_done_kernel (or the consumer's last kernel) the release is resident
        // before that kernel ends; it waits for its completion (griddepcontrol.wait) before it publishes
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof lc);
        lc.gridDim = dim3(1); lc.blockDim = dim3(1); lc.stream = c->stream;
        cudaLaunchAttribute la[1];
        la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        la[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = la; lc.numAttrs = c->p2p_pdl ? 1 : 0;
        CUDA_TRY(cudaLaunchKernelEx(&lc, p2p_release_kernel, (P2PShared*)c->p2p_base, c->p2p_seq - 1));
    }
    CUDA_TRY(cudaGetLastError());
    return VXRT_OK;
}

// owner: queue "acquire every rank's flag -> copy the frame to page-locked host memory -> release the buffer" on the
// copy stream and return; the main stream is free to start the next frame (double-buffered target)
extern "C" int vxrt_p2p_readback(vxrt_ctx* c, uint8_t* out) {
    CHECK_CTX(c);
    if (!c->p2p || !c->p2p_owner) return fail(VXRT_ERR_STATE, "p2p_readback: not the owner of a peer-memory target");
    if (c->p2p_seq == 0) return fail(VXRT_ERR_STATE, "p2p_readback before render");
    if (!out) return fail(VXRT_ERR_INVALID, "p2p_readback: null output");
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) return fail(VXRT_ERR_INVALID, "p2p_readback needs a page-locked destination (vxrt_host_alloc)");
    const unsigned long long seq = c->p2p_seq - 1;
    CUDA_TRY(cudaEventRecord(c->ev_band[0], c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_band[0], 0));
    p2p_wait_done_kernel<<<1, 32, 0, c->copy_stream>>>((const P2PShared*)c->p2p_base, c->cfg.world, seq, c->d_p2p_err);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, c->p2p_base + sizeof(P2PShared) + (seq & 1) * c->p2p_frame_bytes, c->p2p_frame_bytes,
                             cudaMemcpyDeviceToHost, c->copy_stream));
    p2p_release_kernel<<<1, 1, 0, c->copy_stream>>>((P2PShared*)c->p2p_base, seq);
    CUDA_TRY(cudaGetLastError());
    return VXRT_OK;
}

extern "C" int vxrt_p2p_error(vxrt_ctx* c) {
    CHECK_CTX(c);
    if (!c->p2p) return 0;
    int e = 0;
    CUDA_TRY(cudaMemcpyAsync(&e, c->d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return e;
}

// ---- host frames ---------------------------------------------------------------------------------
struct vxrt_host_frame {
    uint8_t* base = nullptr;            // mapping: pixels, then (page-aligned) HostFrameFlags
    size_t pixel_bytes = 0, total = 0;
    int width = 0, height = 0;
    bool owner = false, registered = false;
    std::string name;
};
struct HostFrameFlags {
    unsigned long long done[64];        // done[rank] = last frame whose pixels of that rank are in the mapping
    unsigned long long released;        // last frame the display rank has finished with
};

static HostFrameFlags* host_frame_flags(vxrt_host_frame* hf) { return (HostFrameFlags*)(hf->base + hf->total - 4096); }

static int host_frame_map(const char* name, int width, int height, bool create, vxrt_host_frame** out) {
    if (!name || !out || width <= 0 || height <= 0) return fail(VXRT_ERR_INVALID, "host_frame: bad argument");
    *out = nullptr;
    const size_t pixel_bytes = (size_t)width * height * 4;
    const size_t total = ((pixel_bytes + 4095) / 4096) * 4096 + 4096;
    const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return fail(VXRT_ERR_IO, std::string("host_frame: shm_open failed for ") + name);
    if (create && ftruncate(fd, (off_t)total) != 0) { close(fd); shm_unlink(name); return fail(VXRT_ERR_IO, "host_frame: ftruncate failed"); }
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size != total) { close(fd); if (create) shm_unlink(name); return fail(VXRT_ERR_INVALID, "host_frame: the shared mapping has a different size (other extents?)"); }
    void* p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { if (create) shm_unlink(name); return fail(VXRT_ERR_IO, "host_frame: mmap failed"); }
    if (create) memset(p, 0, total);
    vxrt_host_frame* hf = new vxrt_host_frame();
    hf->base = (uint8_t*)p; hf->pixel_bytes = pixel_bytes; hf->total = total; hf->width = width; hf->height = height;
    hf->owner = create; hf->name = name;
    const cudaError_t e = cudaHostRegister(p, total, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        munmap(p, total);
        if (create) shm_unlink(name);
        delete hf;
        return fail(VXRT_ERR_CUDA, std::string("host_frame: cudaHostRegister: ") + cudaGetErrorString(e));
    }
    hf->registered = true;
    *out = hf;
    return VXRT_OK;
}

extern "C" int vxrt_host_frame_create(const char* name, int width, int height, vxrt_host_frame** out) { return host_frame_map(name, width, height, true, out); }
extern "C" int vxrt_host_frame_open(const char* name, int width, int height, vxrt_host_frame** out) { return host_frame_map(name, width, height, false, out); }
extern "C" uint8_t* vxrt_host_frame_pixels(vxrt_host_frame* hf) { return hf ? hf->base : nullptr; }

extern "C" void vxrt_host_frame_destroy(vxrt_host_frame* hf) {
    if (!hf) return;
    if (hf->registered) { cudaHostUnregister(hf->base); cudaGetLastError(); }
    munmap(hf->base, hf->total);
    if (hf->owner) shm_unlink(hf->name.c_str());
    delete hf;
}

// bounded host-side spin on a flag in the shared mapping
static bool host_spin_until(const volatile unsigned long long* flag, unsigned long long want, int timeout_ms) {
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; spins++) {
        if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) >= want) return true;
        if ((spins & 1023u) == 1023u) {
            if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) return false;
            std::this_thread::yield();
        }
    }
}

extern "C" int vxrt_render_to_host_frame(vxrt_ctx* c, const vxrt_frame* f, vxrt_host_frame* hf, uint64_t seq) {
    if (!hf || seq == 0) return fail(VXRT_ERR_INVALID, "render_to_host_frame: null host frame or seq 0");
    int rc = vxrt_set_frame(c, f);
    if (rc != VXRT_OK) return rc;
    CHECK_CTX(c);
    if (hf->width != c->cfg.width || hf->height != c->cfg.height) return fail(VXRT_ERR_INVALID, "render_to_host_frame: the host frame has other extents than the context");
    if (c->cfg.rank >= 64) return fail(VXRT_ERR_INVALID, "render_to_host_frame: at most 64 ranks");
    HostFrameFlags* fl = host_frame_flags(hf);
    // the previous frame of THIS host frame must have been released by the display rank before it is overwritten: before the
    // kernels are launched when they store into the host frame themselves; when they render into a local buffer first (tile-row
    // partition) it is enough to wait before the COPY is queued, and the kernels of this frame run while the display rank still
    // holds the frame before last.  Measured (frame in host memory, ms per 4K frame): 2 GPUs 0.806 -> 0.709, 4 GPUs 0.579 -> 0.484,
    // but 8 GPUs 0.319 -> 0.393 (profiles/r2_call2{1,2,3}_*): the late wait is used up to 4 ranks.
    auto wait_released = [&]() -> int {
        if (seq > 1 && !host_spin_until(&fl->released, seq - 1, 4000))
            return fail(VXRT_ERR_STATE, "render_to_host_frame: the display rank did not release the previous frame (4 s)");
        return VXRT_OK;
    };
    void* dpix = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer(&dpix, hf->base, 0));
    unsigned long long* dflag = (unsigned long long*)((uint8_t*)dpix + ((uint8_t*)&fl->done[c->cfg.rank] - hf->base));
    if (c->map.rows) {
        // Tile-row partition (vxrt_set_partition 1): this rank's pixels are whole 8-row strips of the raster frame.  The kernels
        // render them into a LOCAL strip buffer and ONE strided DMA (source pitch = a strip, destination pitch = world strips) moves
        // them into the shared host frame on the copy stream -- a DMA fills a PCIe link, 128-byte stores from kernels reach about
        // 40 % of it -- while the main stream already renders the next frame into the other local buffer.  The completion flag
        // follows the copy in the copy stream.
        const bool late_wait = c->cfg.world <= 4;
        if (!late_wait) {
            rc = wait_released();
            if (rc != VXRT_OK) return rc;
        }
        if (!c->d_rgba8_alt) CUDA_TRY(cudaMalloc(&c->d_rgba8_alt, c->out_pixels * 4));
        const int slot = (int)(c->submit_seq & 1);
        if (c->slot_busy[slot]) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_slot[slot], 0));     // its previous copy still reads it
        uint32_t* local = slot ? c->d_rgba8_alt : c->d_rgba8;
        rc = render_bands(c, 1, nullptr, local, /*fence_main=*/false, /*raster_out=*/false, /*local_out=*/true);
        if (rc != VXRT_OK) return rc;
        CUDA_TRY(cudaEventRecord(c->ev_band[0], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_band[0], 0));
        if (late_wait) {                                    // (the kernels are already running)
            rc = wait_released();
            if (rc != VXRT_OK) return rc;
        }
        const int W = c->cfg.width, H = c->cfg.height, world = c->cfg.world, rank = c->cfg.rank;
        const size_t strip = (size_t)W * TILE_H * 4;
        const int my_rows = (c->map.ty > rank) ? (c->map.ty - rank + world - 1) / world : 0;      // tile rows rank, rank + world, ...
        if (my_rows > 0) {
            const int last_row = rank + (my_rows - 1) * world;                                    // may be cut by the frame's lower edge
            const int last_h = (H - last_row * TILE_H < TILE_H) ? H - last_row * TILE_H : TILE_H;
            const int full = (last_h == TILE_H) ? my_rows : my_rows - 1;
            uint8_t* dst0 = hf->base + (size_t)rank * strip;
            if (full > 0)
                CUDA_TRY(cudaMemcpy2DAsync(dst0, (size_t)world * strip, local, strip, strip, (size_t)full, cudaMemcpyDeviceToHost, c->copy_stream));
            if (full < my_rows)
                CUDA_TRY(cudaMemcpyAsync(dst0 + (size_t)full * world * strip, (const uint8_t*)local + (size_t)full * strip, (size_t)W * last_h * 4,
                                         cudaMemcpyDeviceToHost, c->copy_stream));
        }
        host_flag_kernel<<<1, 1, 0, c->copy_stream>>>(dflag, (unsigned long long)seq);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        CUDA_TRY(cudaEventRecord(c->ev_slot[slot], c->copy_stream));
        c->slot_busy[slot] = true;
        c->submit_seq++;
        return VXRT_OK;
    }
    rc = wait_released();
    if (rc != VXRT_OK) return rc;
    rc = render_bands(c, 1, nullptr, (uint32_t*)dpix, /*fence_main=*/true, /*raster_out=*/true);
    if (rc != VXRT_OK) return rc;
    host_flag_kernel<<<1, 1, 0, c->stream>>>(dflag, (unsigned long long)seq);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return VXRT_OK;
}

extern "C" int vxrt_host_frame_wait(vxrt_host_frame* hf, int world, uint64_t seq, int timeout_ms) {
    if (!hf || world < 1 || world > 64) return fail(VXRT_ERR_INVALID, "host_frame_wait: bad argument");
    HostFrameFlags* fl = host_frame_flags(hf);
    for (int r = 0; r < world; r++)
        if (!host_spin_until(&fl->done[r], seq, timeout_ms)) return fail(VXRT_ERR_STATE, "host_frame_wait: rank " + std::to_string(r) + " did not deliver the frame in time");
    return VXRT_OK;
}

extern "C" int vxrt_host_frame_release(vxrt_host_frame* hf, uint64_t seq) {
    if (!hf) return fail(VXRT_ERR_INVALID, "host_frame_release: null host frame");
    __atomic_store_n(&host_frame_flags(hf)->released, (unsigned long long)seq, __ATOMIC_RELEASE);
    return VXRT_OK;
}

extern "C" int vxrt_assemble_tiles(vxrt_ctx* c, const void* gathered, void* dst, void* stream) {
    CHECK_CTX(c);
    if (!gathered || !dst) return fail(VXRT_ERR_INVALID, "assemble_tiles: null pointer");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const int n = c->cfg.width * c->cfg.height;
    TileMap m = c->map;
    assemble_kernel<<<(n + 255) / 256, 256, 0, s>>>((const uint32_t*)gathered, (uint32_t*)dst, m);
    CUDA_TRY(cudaGetLastError());
    return VXRT_OK;
}
