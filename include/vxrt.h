/*
 * vxrt.h -- C ABI of libvxrt.so, the B200 (sm_100a) implementation of voxel-rt's per-pixel hot path.
 *
 * The reference (Berry2460/voxel-rt) has no plugin API; its seam is src/render.hpp:30-41 plus the
 * OpenGL calls src/render.cpp makes (glBufferData / glBufferSubData / glUniform* / glDrawArrays).
 * Every entry point below names the reference interface it replaces (file:line under src/).
 * Plain pointers and sizes only; no C++/torch types.  All functions return 0 on success and a
 * negative vxrt_status on failure (the reference surfaces no errors at all; its render.hpp shim
 * may ignore them); vxrt_last_error() gives the message of the calling thread's last failure.
 *
 * Threading: like the reference's GL usage, one caller thread per context.
 * Ownership: upload calls COPY from the caller's (pageable or pinned) memory at call time, exactly
 * like glBufferData / glBufferSubData; the host array stays the caller's.
 *
 * Frame buffers use the GL window convention: row 0 is the BOTTOM row.
 */
#ifndef VXRT_H
#define VXRT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXRT_MAX_LOCAL_LIGHTS 16     /* render.hpp:11, fshader.glsl:6 */
#define VXRT_TILE_W 32               /* image-tile partition unit (multi-GPU sort-first split) */
#define VXRT_TILE_H 8

typedef enum {
    VXRT_OK = 0,
    VXRT_ERR_INVALID = -1,           /* bad argument */
    VXRT_ERR_CUDA = -2,              /* CUDA runtime failure (message has the cudaError string) */
    VXRT_ERR_NO_DEVICE = -3,         /* no usable sm_100 device: there is NO CPU fallback */
    VXRT_ERR_STATE = -4,             /* call out of order (e.g. render before any grid upload) */
    VXRT_ERR_IO = -5
} vxrt_status;

typedef struct vxrt_ctx vxrt_ctx;

/* flags */
#define VXRT_FLAG_DEBUG_OUTPUTS 1u   /* keep per-pixel hit index / step count / shadow masks (parity tests) */

typedef struct {
    int32_t grid_w, grid_h, grid_d;  /* voxel grid extents; reference: 512, 96, 512 (render.hpp:4-5) */
    int32_t width, height;           /* frame size in pixels (main.cpp:11-12, reshape render.cpp:404-411) */
    int32_t device;                  /* CUDA device ordinal */
    int32_t rank, world;             /* image-tile partition: this context renders one tile of every group of `world` consecutive tiles
                                        (rotated per tile row so that its tiles do not line up in columns; vxrt_set_partition) */
    uint32_t flags;
} vxrt_config;

/* == the shader's uniforms, fshader.glsl:20-26, as written by updateUniforms() render.cpp:289-296 */
typedef struct {
    float cam_pos[3];
    float cam_rotation[2];           /* uploaded by the reference, never read by its shader */
    float light_pos[3];
    float aspect;
    float rotate[16];                /* column-major mat4 (render.cpp:294, GL_FALSE) */
    int32_t view_depth_field;        /* 1 = step-count view (fshader.glsl:143-145) */
    float lights[VXRT_MAX_LOCAL_LIGHTS][4];   /* xyz + diffuse weight; slot inactive if any of xyz < 0 */
} vxrt_frame;

typedef struct {
    /* ray counts by the REFERENCE's casting rule: primary = pixels rendered by this context;
       global = one per hit pixel; local = one per (hit pixel, light) the reference shader would cast */
    uint64_t rays_primary, rays_global, rays_local;
    uint64_t fetches;                /* castRay iterations == voxel fetches of the reference algorithm */
    uint64_t fetches_primary;        /* ... of which by primary rays (the rest: shadow / light rays) */
    uint64_t rays_dark;              /* of rays_global + rays_local: rays toward a light the surface faces away from (N.L <= 0);
                                        their outcome cannot change the pixel and production frames do not trace them */
    uint64_t hit_pixels;
    float ms_primary, ms_shadow, ms_total;   /* CUDA-event times of the last vxrt_render */
    uint32_t kernel_launches;        /* kernels launched by the last vxrt_render */
} vxrt_stats;

/* ---- lifetime --------------------------------------------------------------------------------- */
/* replaces initRender()'s GL object / SSBO creation, render.cpp:313-372 */
int vxrt_create(const vxrt_config* cfg, vxrt_ctx** out);
void vxrt_destroy(vxrt_ctx* ctx);
const char* vxrt_last_error(void);
/* 1 iff a device of compute capability 10.x is present */
int vxrt_device_available(void);
/* page-locked host memory for frame read-back (vxrt_render_frame_host / vxrt_read_rgba8 copy straight into it);
   any other host pointer works too, through an internal staging buffer */
void* vxrt_host_alloc(size_t nbytes);
void vxrt_host_free(void* p);
/* FNV-1a-64 of a host buffer as SURVEY.md 8c fingerprints grids (prime 1099511628211, offset basis
   1469598103934665603) */
uint64_t vxrt_fnv1a64(const void* data, size_t nbytes);

/* ---- grid ------------------------------------------------------------------------------------- */
/* updateGeometry() render.cpp:199-202 / glBufferData render.cpp:368: whole grid, count must equal w*h*d */
int vxrt_upload_grid(vxrt_ctx* ctx, const int32_t* voxels, size_t count);
/* one glBufferSubData(offset=first*4, size=count*4, &voxels[first]) render.cpp:219 */
int vxrt_upload_range(vxrt_ctx* ctx, size_t first, size_t count, const int32_t* src);
/* updatePartialGeometry(start,end) render.cpp:204-223, with its row-skipping / wrapping behaviour, as ONE
   staged copy + scatter kernel; host_voxels is the caller's full grid (the reference's voxels[]).
   Rows that would run past the end of the buffer are dropped (GL_INVALID_VALUE in the reference).
   *rows_out (optional) receives the number of rows uploaded (the reference's glBufferSubData call count). */
int vxrt_update_partial(vxrt_ctx* ctx, const float start[3], const float end[3], const int32_t* host_voxels,
                        int32_t* rows_out);
/* a batch of glBufferSubData calls of one length -- what ONE updatePartialGeometry issues (render.cpp:214-221: 900 calls
   of 124 bytes per destruction) -- as one staged copy + scatter kernel: row r covers voxels [firsts[r], firsts[r]+row_len),
   its data is packed[r*row_len ...] (copied at call time).  Rows of one batch must not overlap unless they carry the
   same data.  Any row outside the buffer fails the whole call (GL_INVALID_VALUE), nothing is uploaded.  Used by the
   link-level GL shim (voxel-rt_b200/csrc/host/vxrt_glshim.cpp). */
int vxrt_upload_rows(vxrt_ctx* ctx, size_t rows, size_t row_len, const int64_t* firsts, const int32_t* packed);
int vxrt_download_grid(vxrt_ctx* ctx, int32_t* out, size_t count);
/* copy the device box [lo,hi) (clamped to the grid) into the caller's FULL-GRID array (host mirror sync after
   a device-side edit; the reference's CPU collision code reads voxels[], controls.cpp:10-19) */
int vxrt_download_box(vxrt_ctx* ctx, const int32_t lo[3], const int32_t hi[3], int32_t* host_voxels);
/* Grid files (the reference has no on-disk format, its level exists only as level.cpp's generator):
   a 64-byte little-endian header -- "VXRTGRD1", uint32 w, h, d, flags(0), uint64 count, uint64 vxrt_fnv1a64 of the
   payload, 24 reserved zero bytes -- followed by count int32 voxels in the reference's order (x + w*y + w*h*z).
   save streams the device grid to the file; load checks extents against the context and the fingerprint against the
   payload (VXRT_ERR_IO on a short / corrupt file, after which the context has no grid) and leaves the grid on the
   device as vxrt_upload_grid would.  voxel-rt_b200/gridfile.py reads and writes the same format on the host. */
int vxrt_save_grid(vxrt_ctx* ctx, const char* path);
int vxrt_load_grid(vxrt_ctx* ctx, const char* path);

/* placeVoxel render.cpp:256-262 / destroyVoxel render.cpp:265-271 applied to the device grid */
int vxrt_place_voxel(vxrt_ctx* ctx, int x, int y, int z, int32_t voxel);
int vxrt_destroy_voxel(vxrt_ctx* ctx, int x, int y, int z);
/* n placeVoxel calls as ONE staged copy + one kernel (the reference places voxels in bulk: placeBush level.cpp:21, placeTrunk
   level.cpp:74, initVoxels level.cpp:102,113,124); xyz = n * 3 ints, voxels = n values; cells outside the grid are ignored,
   where cells repeat the last entry wins -- exactly what the sequence of calls would leave */
int vxrt_place_voxels(vxrt_ctx* ctx, size_t n, const int32_t* xyz, const int32_t* voxels);
/* removeSphere(pos, radius) level.cpp:30-56 executed on the device grid (carve + fixDepthField over the
   radius+3 sphere); no host upload needed afterwards */
int vxrt_edit_remove_sphere(vxrt_ctx* ctx, int cx, int cy, int cz, int radius);
/* The same edit from a 16-byte command {cx, cy, cz, radius} (4 x int32) in DEVICE memory: queued on the context's stream
   (vxrt_stream()), the host never reads the command -- in the multi-GPU split (BASELINE configs[4]) the command is the target
   of ONE NCCL broadcast per edit queued on that stream, and every replica replays it without a host synchronisation.  The
   kernels are sized for max_radius (<= 64); a command whose radius lies outside [0, max_radius] is refused and
   vxrt_edit_cmd_error() (which synchronises) returns 1. */
int vxrt_edit_remove_sphere_cmd(vxrt_ctx* ctx, const int32_t* device_cmd, int max_radius);
int vxrt_edit_cmd_error(vxrt_ctx* ctx);
/* computeDepthField sweep over the whole grid, render.cpp:226-253,273-286 (out-of-grid neighbours = solid) */
int vxrt_build_depth_field(vxrt_ctx* ctx);

/* ---- procedural levels, generated on the device ------------------------------------------------- */
/* the reference's default level: fill -1 (render.cpp:349-352) + initVoxels() (level.cpp:82-138) incl. its
   origin-carving quirk (level.cpp:11,64); no depth field yet (the reference builds it in background threads) */
int vxrt_generate_default_level(vxrt_ctx* ctx);
/* synthetic terrain of config C4: integer-only 5-octave value-noise height field
   (lattice hash = splitmix64(seed ^ ix*73856093 ^ iz*19349663 ^ octave*83492791) >> 48), the reference's
   stone / dirt(8) / grass(3) bands and colour jitter relative to the surface, the reference's trees
   (x%30==0, z%25==0) standing on it.  No depth field. */
int vxrt_generate_terrain(vxrt_ctx* ctx, uint64_t seed);
/* surface height of that terrain at (x,z) for a grid of height grid_h (host-side, for camera placement) */
int vxrt_terrain_height(uint64_t seed, int x, int z, int grid_h);

/* ---- frame ------------------------------------------------------------------------------------ */
/* updateUniforms() render.cpp:289-296 */
int vxrt_set_frame(vxrt_ctx* ctx, const vxrt_frame* frame);
/* initLocalLights() render.cpp:304-311 / placeLocalLight() render.cpp:375-385 acting on the context's frame;
   place returns the slot used (>= 0) or VXRT_MAX_LOCAL_LIGHTS when all slots are taken (silently ignored) */
int vxrt_init_local_lights(vxrt_ctx* ctx);
int vxrt_place_local_light(vxrt_ctx* ctx, float x, float y, float z, float diffuse);
int vxrt_get_frame(vxrt_ctx* ctx, vxrt_frame* out);
/* reshape() render.cpp:404-411: new frame size; aspect := (float)width/height */
int vxrt_resize(vxrt_ctx* ctx, int width, int height);
/* glDrawArrays(GL_TRIANGLES,0,6) main.cpp:59: runs the per-pixel path for this context's tiles; asynchronous */
int vxrt_render(vxrt_ctx* ctx);
int vxrt_sync(vxrt_ctx* ctx);
/* mode 1: every frame starts with a streaming read of the part of the grid rays can reach, which allocates it in L2
   (grids up to 120 MB).  It pays off when frames start with a cold L2 and the context renders few pixels -- the
   critical path of its longest rays is otherwise a chain of ~1 us HBM round trips -- and costs ~15 us otherwise.
   mode 0: never; mode 2 (default): when this context renders at most 12,000 tiles (1080p, or a quarter of a 4K frame). */
int vxrt_set_l2_prefetch(vxrt_ctx* ctx, int mode);
/* enabled (default): production frames (counters off, no debug planes, not the step-count view) end a ray as a miss
   as soon as its cell lies beyond every grid row that holds a solid voxel, in its direction of travel -- the coarsest
   level of an occupancy hierarchy -- and do not trace shadow / light rays from surfaces that face away from the light
   (their term is multiplied by max(0, N.L) = 0).  The first-hit voxel and every pixel are unchanged (ray.cuh CULL,
   kernels.cuh skip_dark); ray statistics always count the rays the reference casts. */
int vxrt_set_culling(vxrt_ctx* ctx, int enabled);
/* The TRAVERSAL GRID (csrc/trav.cuh; north_star item 1): the device copy of the grid the rays read.  Same index and 4 bytes per
   cell as the reference layout; a -1 cell (empty, no depth-field jump) carries, per travel quadrant, how many further cells of
   its y layer (and of the layer above) are -1 cells inside the grid, and castRay takes those steps without index arithmetic,
   range test or load -- exact by adjacency: a step moves one cell along one axis whatever the float state says.  It is kept
   coherent by every upload / edit entry point (a rebuild over the cells whose words can change).  mode 0: rays read the
   reference-layout grid with the plain kernels; 1: both passes read the traversal grid; 2 (default, auto): the shade pass
   always, the primary pass when this context renders <= 12,000 tiles (a 1080p frame, a share of a 4K frame: there its runs
   shorten the critical path; on a whole 4K frame they cost the primary pass throughput).  Same pixels in every mode.
   vxrt_traversal_active: 1 if the next frame reads the traversal grid (0 also when the grid holds a value that cannot be
   encoded: a negative int with bit 30 clear other than through -1, which the reference never produces).
   vxrt_download_traversal: the words, for tests. */
int vxrt_set_traversal(vxrt_ctx* ctx, int mode);
int vxrt_traversal_active(vxrt_ctx* ctx);
int vxrt_download_traversal(vxrt_ctx* ctx, int32_t* out, size_t count);
/* enabled (default): the primary pass records how long each tile's block took and the next frame launches the
   slowest tiles first (shorter kernel tail; it matters when a GPU renders only a fraction of the frame).  Same pixels. */
int vxrt_set_tile_ordering(vxrt_ctx* ctx, int enabled);
/* Overlap of the two passes of a two-kernel frame: 0 off (default), 1 on, 2 auto (on when this context renders <= 12,000
   tiles).  Superseded by vxrt_set_fusion for small shares (measured on 1/8 of a 4K frame: 0.190 ms separate, 0.181 ms
   overlapped); kept as a switch.  On: the shade kernel is launched with programmatic stream serialization and the primary kernel releases it at once
   (griddepcontrol.launch_dependents), so shade blocks are scheduled into the SM capacity the primary pass's tail leaves idle;
   a shade block waits for ITS tile's ready flag (release / acquire, bounded spin) instead of the kernel boundary.  Same
   pixels.  While on, vxrt_get_stats cannot separate the passes: ms_primary reads as the whole frame, ms_shadow as 0. */
int vxrt_set_overlap(vxrt_ctx* ctx, int mode);
/* Fusion of the two passes: 0 two kernels per frame (primary pass, then shade pass over the hit slots), 1 one kernel per frame
   in which every block traces its tile's primary rays and then shades its own hits (hit records stay in shared memory),
   2 auto (default): fused when this context renders <= 12,000 tiles.  For a small share the frame time is set by serial
   chains (the longest primary ray, then the longest pixel's sequential light loop); with a kernel boundary between the
   passes the two chains add, fused they overlap.  On a whole 4K frame the two specialised kernels (40 / 48 registers, 6 / 5
   blocks per SM) have the higher throughput.  Same pixels.  While fused, vxrt_get_stats cannot separate the passes
   (ms_primary = whole frame, ms_shadow = 0), and vxrt_set_overlap has no effect. */
int vxrt_set_fusion(vxrt_ctx* ctx, int mode);
/* Which image tiles a context of a multi-GPU split renders (all ranks must choose the same): 0 (default) groups of `world`
   consecutive tiles, one to each rank, rotated by the tile row the group starts in; 1 whole TILE ROWS are the interleaved unit -- tile row r belongs to rank r % world -- so that a rank's pixels are
   contiguous 8-row strips of the raster frame.  With 1, vxrt_render_to_host_frame renders into a local strip buffer and moves the
   strips into the shared host frame with ONE strided DMA on the copy stream (a DMA fills a PCIe link; stores from kernels reach
   about 40 % of it), overlapped with the next frame; local buffers / the all-gather layout become [local strip][8][width].
   Re-allocates the per-frame buffers.  Same pixels. */
int vxrt_set_partition(vxrt_ctx* ctx, int mode);
/* Fused frames only: how many of the heaviest tiles (previous frames' block times) are rendered by TWO blocks each with two threads
   per hit pixel -- one evaluates the global shadow ray and the first half of the active lights, the other the second half without
   the early-out; the terms are combined in slot order afterwards (what a light adds does not depend on the multiplier,
   fshader.glsl:161-179), so the pixel is the same.  A small share of a frame lasts as long as its slowest block; this halves that
   block's sequential chain.  0 = off, default 8, at most 64.  Same pixels. */
int vxrt_set_wide_tiles(vxrt_ctx* ctx, int tiles);
/* The partition's arithmetic, host only (no device needed): the global tile (row-major, 32x8 pixels) that local tile `local_tile` of
   `rank` renders, or -1 for padding; and the rank that owns global tile `tile`.  tile_rows as in vxrt_set_partition. */
int vxrt_partition_tile(int width, int height, int rank, int world, int tile_rows, int local_tile);
int vxrt_partition_owner(int width, int height, int world, int tile_rows, int tile);
/* 1 when the last vxrt_render ran as one fused kernel (see vxrt_set_fusion), else 0. */
int vxrt_frame_was_fused(vxrt_ctx* ctx);
/* mode 0 (default) = the production kernels: no per-iteration counter, rays that cannot change a pixel are not traced
   (vxrt_set_culling); vxrt_get_stats then reads rays_local / fetches / rays_dark as 0 while hit_pixels, rays_primary,
   rays_global and the timings stay valid.  mode 1: every vxrt_render runs the counted kernel variants by the REFERENCE's
   casting rule (every ray fshader.glsl casts is marched to its end and counted; rays_dark = how many of them face away from
   their light; ~1.5x the frame time).  mode 2: counted variants that skip exactly what the production kernels skip:
   rays_global / rays_local = rays traced, fetches = DDA iterations executed, rays_dark = rays not traced.  Same pixels in
   every mode.  A debug context (VXRT_FLAG_DEBUG_OUTPUTS) counts by mode 1 unless mode 2 is set; the step-count view
   always counts its primary rays. */
int vxrt_set_stats(vxrt_ctx* ctx, int mode);
/* set_frame + render + device->host copy of the RGBA8 frame into out (width*height*4 bytes) + sync.
   world > 1: out receives this rank's tiles in gather layout (vxrt_local_bytes()). */
int vxrt_render_frame_host(vxrt_ctx* ctx, const vxrt_frame* frame, uint8_t* out);

/* Pipelined form of vxrt_render_frame_host for a stream of frames: queues set_frame + render + read-back into the
   page-locked buffer `out` and returns; frames alternate between two device buffers so that the read-back of one
   frame overlaps the kernels of the next.  `out` may be reused once vxrt_wait_frames returned (or two submits later).
   vxrt_wait_frames blocks until every submitted frame is complete in host memory. */
int vxrt_submit_frame_host(vxrt_ctx* ctx, const vxrt_frame* frame, uint8_t* out);
int vxrt_wait_frames(vxrt_ctx* ctx);
/* vxrt_render_frame_host renders the frame in `nbands` bands of tile rows (default 2, 1..16) and copies each band
   to the host while the next one renders.  Same pixels for every value. */
int vxrt_set_readback_bands(vxrt_ctx* ctx, int nbands);

/* ---- results ---------------------------------------------------------------------------------- */
int vxrt_read_rgba8(vxrt_ctx* ctx, uint8_t* out);            /* width*height*4 (world==1) else vxrt_local_bytes() */
/* parity outputs (need VXRT_FLAG_DEBUG_OUTPUTS); full-frame arrays, any pointer may be NULL:
   hit_index: primary castRay result (-1 miss); steps: primary stepCount; occl_mask bit0 = global-light ray
   occluded, bit(1+i) = local light i occluded; cast_mask: same bits, "ray was cast" */
int vxrt_read_debug(vxrt_ctx* ctx, int32_t* hit_index, uint16_t* steps, uint32_t* occl_mask, uint32_t* cast_mask);
int vxrt_get_stats(vxrt_ctx* ctx, vxrt_stats* out);
/* diagnostics: SM cycles each local tile's primary block and each shade unit's block took in the last whole-frame launch (what
   the slowest-first launch orders are made from).  primary: vxrt_local_tiles() entries; shade: shade_count = local tiles x
   (256 / threads per shade block, 2 by default) entries; either may be NULL */
int vxrt_read_block_costs(vxrt_ctx* ctx, uint32_t* primary, uint32_t* shade, size_t shade_count);
/* known-answer hook: castRay(start, dir, dist) fshader.glsl:59-129 for n independent rays (host arrays:
   starts/dirs n*3 floats, dists n ints); ret[n] = return value, out7[n*7] = hitPos[3] hitNormal[3] stepCount */
int vxrt_cast_rays(vxrt_ctx* ctx, int32_t n, const float* starts, const float* dirs, const int32_t* dists,
                   int32_t* ret, float* out7);
/* device self-test of the kernels' exact division-by-ray-direction (ray.cuh div_by) against IEEE division on n
   pseudo-random operand pairs inside its documented domain; *mismatches must come back 0 */
int vxrt_selftest_division(vxrt_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* device self-test of the kernels' step reciprocals 1 / |dir + 0.000001| (fshader.glsl:74-76; ray.cuh refined_rcp) against the
   IEEE reciprocal on EVERY float of [2^-40, 4); *mismatches must come back 0 */
int vxrt_selftest_reciprocal(vxrt_ctx* ctx, uint64_t* mismatches);
/* binary PPM (P6), rows flipped so the image is upright */
int vxrt_write_ppm(vxrt_ctx* ctx, const char* path);

/* ---- multi-GPU plumbing (device pointers; the collective itself is the caller's, e.g. NCCL all-gather) -- */
size_t vxrt_local_tiles(vxrt_ctx* ctx);                      /* tiles per rank (padded: ceil(tiles/world)) */
size_t vxrt_local_bytes(vxrt_ctx* ctx);                      /* local_tiles * TILE_W*TILE_H*4 */
void*  vxrt_device_rgba8(vxrt_ctx* ctx);                     /* device pointer of this rank's output buffer */
void*  vxrt_stream(vxrt_ctx* ctx);                           /* cudaStream_t the context launches on */
/* gathered: device pointer to world * vxrt_local_bytes() bytes (rank-major, as produced by an all-gather of
   vxrt_device_rgba8()); dst: device pointer to width*height*4 bytes; un-tiles into a raster frame on
   `stream` (NULL = the context's stream) */
int vxrt_assemble_tiles(vxrt_ctx* ctx, const void* gathered, void* dst, void* stream);

/* ---- multi-GPU without a gather: peer-memory frame target ------------------------------------------------
   One rank (the display rank) owns a double-buffered raster frame; every rank's kernels store their tiles'
   pixels straight into it over NVLink / NVSwitch while they render (no collective, no un-tile pass), then publish
   completion with a system-scope release that the owner acquires before it reads the frame.
     owner : vxrt_p2p_export(ctx, handle)  -> send the 64-byte handle to the other ranks (any transport)
     others: vxrt_p2p_import(ctx, handle)
     every frame, every rank: vxrt_set_frame + vxrt_render        (waits, on the stream, for the owner to have released
                                                                   the buffer it is about to overwrite)
     owner : vxrt_p2p_wait_frame(ctx, &dptr) -> dptr = complete width*height RGBA8 raster frame (stream-ordered)
             ... consume it on vxrt_stream() ...   vxrt_p2p_release_frame(ctx)
   vxrt_p2p_error: 0, or non-zero if a bounded device-side wait timed out (a rank stopped participating). */
int vxrt_p2p_export(vxrt_ctx* ctx, uint8_t handle[64]);
int vxrt_p2p_import(vxrt_ctx* ctx, const uint8_t handle[64]);
/* same-process variant of import: several contexts in one process; owner_base = vxrt_p2p_base(owner ctx) */
int vxrt_p2p_attach(vxrt_ctx* ctx, void* owner_base);
void* vxrt_p2p_base(vxrt_ctx* ctx);
int vxrt_p2p_wait_frame(vxrt_ctx* ctx, void** frame);
int vxrt_p2p_release_frame(vxrt_ctx* ctx);
/* owner, instead of wait_frame / release_frame when the consumer is host memory: queues acquire -> device-to-host copy
   into the page-locked buffer `out` -> release on the context's copy stream and returns, so the next frame renders
   while this one is read back; vxrt_wait_frames blocks until the copies are complete */
int vxrt_p2p_readback(vxrt_ctx* ctx, uint8_t* out);
int vxrt_p2p_error(vxrt_ctx* ctx);

/* ---- frames straight to host memory, any number of GPUs, no exchange ---------------------------
   When the consumer of the frame is the HOST (a glReadPixels after the draw, an encoder, a file), the ranks need not
   send their pixels to one GPU first: a host frame is ONE raster in page-locked host memory (POSIX shared memory,
   registered with CUDA by every process that opens it) and the render kernels of every rank store their own tiles'
   pixels straight into it, each GPU over its own PCIe link.
     display rank : vxrt_host_frame_create(name, w, h, &hf)        other ranks : vxrt_host_frame_open(name, w, h, &hf)
     every rank   : vxrt_render_to_host_frame(ctx, &frame, hf, seq)   (seq = 1, 2, 3, ...; returns when queued)
     display rank : vxrt_host_frame_wait(hf, world, seq, timeout_ms) -> vxrt_host_frame_pixels(hf) holds frame seq
                    ... consume ...   vxrt_host_frame_release(hf, seq)
   Alternate between two host frames to overlap the consumption of one frame with the rendering of the next;
   render_to_host_frame(seq) first waits (host side, bounded) until the display rank released frame seq - 1 of the
   same host frame.  The completion flags live behind the pixels in the same shared mapping. */
typedef struct vxrt_host_frame vxrt_host_frame;
int vxrt_host_frame_create(const char* name, int width, int height, vxrt_host_frame** out);
int vxrt_host_frame_open(const char* name, int width, int height, vxrt_host_frame** out);
uint8_t* vxrt_host_frame_pixels(vxrt_host_frame* hf);
int vxrt_render_to_host_frame(vxrt_ctx* ctx, const vxrt_frame* frame, vxrt_host_frame* hf, uint64_t seq);
int vxrt_host_frame_wait(vxrt_host_frame* hf, int world, uint64_t seq, int timeout_ms);
int vxrt_host_frame_release(vxrt_host_frame* hf, uint64_t seq);
void vxrt_host_frame_destroy(vxrt_host_frame* hf);

#ifdef __cplusplus
}
#endif
#endif
