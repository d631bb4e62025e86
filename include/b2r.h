/* b2r.h -- C ABI of the B200-native render hot paths (libb2r.so).
 *
 * Drop-in boundary for the two `void Draw()` functions of
 * ArchDD/CPP-Raytracer-Rasterizer:
 *     raytracer/Source/raytracer.cpp:547-606    (Draw -> ClosestIntersection -> DirectLight)
 *     rasteriser/Source/rasteriser.cpp:461-482  (Draw -> DrawPolygon -> ... -> PixelShader)
 * The reference has no FFI: Draw() takes no arguments and talks through
 * file-scope globals.  This header is what a binding of that path looks like
 * once the globals are made explicit: plain pointers and sizes, POD structs
 * whose fields are the reference's globals (cited per field), no C++ or torch
 * types.  Every function returns 0 on success or a negative B2R_E_* code and
 * never throws or exits; b2r_last_error() gives the text.
 *
 * Ownership: the context owns all device memory; the caller owns every host
 * pointer it passes; no host pointer is retained after a call returns.
 * Threading: one context per GPU and per host thread (externally
 * synchronised).  All calls are synchronous on return unless named *_async.
 */
#ifndef B2R_H
#define B2R_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_ABI_VERSION 1

#define B2R_MAX_LIGHTS 32         /* Light lights[32]: raytracer.cpp:48, rasteriser.cpp:50 */
#define B2R_RANDOM_POSITIONS 256  /* vec3 randomPositions[256]: raytracer.cpp:84 */

enum {
    B2R_OK = 0,
    B2R_E_INVALID = -1,    /* bad argument */
    B2R_E_CUDA = -2,       /* CUDA runtime error (text in b2r_last_error) */
    B2R_E_NO_SCENE = -3,   /* draw before b2r_set_triangles / b2r_set_frame */
    B2R_E_UNSUPPORTED = -4,
    B2R_E_IO = -5,
    B2R_E_CAPACITY = -6    /* a projected triangle exceeds the row limit (see b2r_ras_draw) */
};

/* == reference `class Light` (raytracer TestModel.h:35-45), 28 bytes */
typedef struct b2r_light {
    float position[3];
    float color[3];
    float intensity;
} b2r_light;

/* == reference `struct Intersection` (raytracer.cpp:91-96), 20 bytes */
typedef struct b2r_intersection {
    float position[3];
    float distance;
    int32_t triangleIndex;
} b2r_intersection;

/* The globals Draw() reads, made explicit. */
typedef struct b2r_frame_params {
    float cameraPos[3];        /* raytracer.cpp:70, rasteriser.cpp:39 */
    float cameraRot[9];        /* glm::mat3 memory order (column-major): raytracer.cpp:73, rasteriser.cpp:40 */
    float focalLength;         /* raytracer.cpp:69, rasteriser.cpp:41 */
    int32_t numLights;         /* NUM_LIGHTS: raytracer.cpp:47, rasteriser.cpp:49 */
    b2r_light lights[B2R_MAX_LIGHTS];
    float randomPositions[B2R_RANDOM_POSITIONS * 3]; /* raytracer.cpp:84 (soft-shadow jitter table; an INPUT) */
    int32_t aaEnabled;         /* AA_ENABLED raytracer.cpp:37 */
    int32_t aaSamples;         /* AA_SAMPLES raytracer.cpp:38 */
    int32_t softShadowsEnabled;/* SOFT_SHADOWS_ENABLED raytracer.cpp:40 */
    int32_t softShadowsSamples;/* SOFT_SHADOWS_SAMPLES raytracer.cpp:41 */
    float indirectLight[3];    /* raytracer.cpp:81 / indirectLightPowerPerArea rasteriser.cpp:47 */
    float dofFocalLength;      /* FOCAL_LENGTH raytracer.cpp:45 / rasteriser.cpp:31 */
    float currentReflectance[3]; /* rasteriser.cpp:46,466 */
    int32_t dofEnabled;        /* DOF_ENABLED raytracer.cpp:43 / rasteriser.cpp:29 */
    int32_t dofKernelSize;     /* DOF_KERNEL_SIZE raytracer.cpp:44 / rasteriser.cpp:30 */
    int32_t backfaceCulling;   /* BACKFACE_CULLING_ENABLED rasteriser.cpp:26 */
    int32_t frustumCulling;    /* FRUSTUM_CULLING_ENABLED rasteriser.cpp:27 */
} b2r_frame_params;

typedef struct b2r_ctx b2r_ctx;

/* ---- lifetime ----------------------------------------------------------- */
int b2r_abi_version(void);
/* One context per GPU; width/height are SCREEN_WIDTH/SCREEN_HEIGHT (raytracer.cpp:67-68). */
int b2r_create(b2r_ctx** out, int device, int width, int height);
int b2r_destroy(b2r_ctx* ctx);
/* Last error text of this context (or of b2r_create when ctx == NULL). */
const char* b2r_last_error(const b2r_ctx* ctx);
/* Fills frame params with the reference's start-up defaults for the raytracer
 * (which = 0: raytracer.cpp:33-81,116,162) or the rasteriser (which = 1:
 * rasteriser.cpp:22-50,104,115), scaled to the given screen size the way the
 * reference scales them (focalLength = H/2 resp. H). */
int b2r_default_frame_params(b2r_frame_params* out, int which, int width, int height);

/* ---- scene -------------------------------------------------------------- */
/* The global `vector<Triangle> triangles` (raytracer.cpp:28, rasteriser.cpp:64)
 * as the reference lays it out in memory: stride 60 (raytracer Triangle:
 * v0,v1,v2,normal,color) or 64 (rasteriser Triangle: + bool isCulled at byte
 * 60).  Copied to the device; the host pointer is not retained. */
int b2r_set_triangles(b2r_ctx* ctx, const void* triangles, int count, int stride_bytes);
/* Replaces only the isCulled flags (rasteriser.cpp:404-447 writes them each frame). */
int b2r_set_culled(b2r_ctx* ctx, const uint8_t* culled, int count);
int b2r_set_frame(b2r_ctx* ctx, const b2r_frame_params* params);

/* ---- raytracer: Draw() of raytracer.cpp:547-606 ------------------------- */
/* Renders rows [y0,y1) (0,H for the whole frame; a band for multi-GPU splits).
 * Includes the per-frame precondition of Update() (raytracer.cpp:335-339):
 * every pixel's Intersection starts at distance FLT_MAX.  Outputs (each may be
 * NULL) are FULL-FRAME host arrays indexed y*width+x; only rows [y0,y1) are
 * written:
 *   pixelColours          3 floats/pixel  (raytracer.cpp:88,600)
 *   closestIntersections  20 bytes/pixel  (raytracer.cpp:98); pixels never hit
 *                         get distance FLT_MAX, triangleIndex -1, position 0
 *   focalDistances        1 float/pixel   (raytracer.cpp:87,249); 0 where never hit */
int b2r_rt_draw(b2r_ctx* ctx, int y0, int y1, float* pixelColours,
                b2r_intersection* closestIntersections, float* focalDistances);

/* ---- rasteriser: Draw() of rasteriser.cpp:461-482 ----------------------- */
/* Includes the per-frame clears of Update() (rasteriser.cpp:183-192).  Outputs
 * (each may be NULL), full-frame host arrays, rows [y0,y1) written:
 *   depthBuffer     1 float/pixel (rasteriser.cpp:52), 0 where nothing drawn
 *   pixelColours    3 floats/pixel (rasteriser.cpp:69,588)
 *   focalDistances  1 float/pixel (rasteriser.cpp:68,565), 0 where nothing drawn
 *   winnerIndex     index of the triangle that owns the depth-buffer value
 *                   (the reference keeps none; oracle patch P4), -1 if none */
int b2r_ras_draw(b2r_ctx* ctx, int y0, int y1, float* depthBuffer, float* pixelColours,
                 float* focalDistances, int32_t* winnerIndex);
/* The culling block of Update() (rasteriser.cpp:385-447) on the device: writes
 * Triangle::isCulled for the current frame params.  culledOut may be NULL. */
int b2r_ras_cull(b2r_ctx* ctx, uint8_t* culledOut);

/* ---- resolve: CalculateDOF() + PutPixelSDL ------------------------------ */
/* raytracer.cpp:608-656 == rasteriser.cpp:484-529, SDLauxiliary.h:70-81.
 * Converts the pixelColours of the last draw on this context to the 32-bit
 * XRGB surface the reference fills (interior pixels only; the 1-pixel border
 * stays 0).  Applies the depth-of-field blur when params.dofEnabled.
 * surface: width*height uint32 (0x00RRGGBB), host memory.
 * "Last draw" means a host-buffer draw (b2r_rt_draw / b2r_ras_draw with pixelColours, b2r_rt_frame, b2r_ras_frame):
 * only those leave a frame in the context's own buffers.  After a *_device_async, *_part or *_bgr8 draw, or a draw
 * without pixelColours, the resolve calls return B2R_E_NO_SCENE instead of an older image. */
int b2r_resolve_surface(b2r_ctx* ctx, uint32_t* surface);
/* Same pixels as 24-bit bottom-up BGR rows padded to 4 bytes (the payload
 * SDL_SaveBMP writes, raytracer.cpp:175).  bgr: b2r_bmp_payload_bytes(). */
int b2r_resolve_bgr8(b2r_ctx* ctx, uint8_t* bgr);
size_t b2r_bmp_payload_bytes(int width, int height);
/* Headless replacement of SDL_SaveBMP(screen, path). */
int b2r_write_bmp(const char* path, const uint8_t* bgr_payload, int width, int height);

/* One call per frame, the Draw() drop-ins: draw + resolve, copy out only what is asked for.
 * surface may be NULL.  Extra outputs as in b2r_rt_draw / b2r_ras_draw, may be NULL. */
int b2r_rt_frame(b2r_ctx* ctx, uint32_t* surface, float* pixelColours,
                 b2r_intersection* closestIntersections, float* focalDistances);
int b2r_ras_frame(b2r_ctx* ctx, uint32_t* surface, float* depthBuffer, float* pixelColours,
                  float* focalDistances, int32_t* winnerIndex);

/* Optional: page-lock a caller-owned output buffer (screen->pixels, pixelColours, ...) once, so that the copies
 * of b2r_rt_frame / b2r_ras_frame / b2r_*_draw run at full PCIe speed and overlap the kernels.  Without it the
 * calls work the same, only slower (pageable memory is copied through the driver's staging buffer).  Idempotent. */
int b2r_pin_host_buffer(b2r_ctx* ctx, void* host, size_t bytes);
int b2r_unpin_host_buffer(b2r_ctx* ctx, void* host);

/* ---- device-resident entry points (no host copies) ---------------------- */
/* For callers that keep frames in HBM (multi-GPU band gathers, benchmarks).
 * All pointers are DEVICE pointers on the context's GPU, full-frame arrays as
 * above, may be NULL.  Work is enqueued on the context's stream;
 * b2r_synchronize() waits for it.  cuda_stream (a cudaStream_t cast to void*)
 * replaces the context's own stream when non-NULL at b2r_set_stream().  A context is driven by one host thread
 * and its draws are ordered on that one stream (they share scheduler words and scratch buffers); use one context
 * per stream for concurrent frames. */
int b2r_set_stream(b2r_ctx* ctx, void* cuda_stream);
void* b2r_get_stream(b2r_ctx* ctx);
int b2r_synchronize(b2r_ctx* ctx);  /* also reports B2R_E_CAPACITY of asynchronous rasteriser draws since the last call */
int b2r_rt_draw_device_async(b2r_ctx* ctx, int y0, int y1, float* d_pixelColours,
                             b2r_intersection* d_closestIntersections, float* d_focalDistances);
/* Rasteriser draws size their intermediate buffers from counters of the frame itself.  Small scenes (triangles x band
 * height <= 2M) use fixed worst-case slots: the call only enqueues, and a projected triangle beyond the row/coordinate
 * limits is reported by the next b2r_synchronize (B2R_E_CAPACITY).  Larger scenes read the counters back ONCE per
 * (scene, isCulled flags, frame params, band): the first draw of a new state blocks the host for one stream
 * synchronisation and reports B2R_E_CAPACITY itself; later draws of the same state only enqueue. */
int b2r_ras_draw_device_async(b2r_ctx* ctx, int y0, int y1, float* d_depthBuffer,
                              float* d_pixelColours, float* d_focalDistances, int32_t* d_winnerIndex);
/* Draw() for rows [y0,y1) in one call: trace (or rasterise) + shade + CalculateDOF/PutPixelSDL into d_surface (full-frame
 * uint32 array).  Without depth of field the surface is written by the trace / shade kernel itself (no second pass;
 * d_pixelColours / d_focalDistances may be NULL).  With depth of field the rows are resolved by a second kernel
 * from d_pixelColours / d_focalDistances (required), whose neighbouring rows must already be up to date. */
int b2r_rt_frame_device_async(b2r_ctx* ctx, int y0, int y1, uint32_t* d_surface, float* d_pixelColours,
                              b2r_intersection* d_closestIntersections, float* d_focalDistances);
/* One frame split over `nparts` GPUs with the exchange inside the trace kernel: this call draws tile rows (8 pixel
 * rows) part, part + nparts, ... of the whole frame -- interleaved, so every part carries the same mix of cheap and
 * expensive rows -- and stores each resolved pixel into all n surfaces: d_surfaces[0] is normally the local one, the
 * others peer-mapped buffers of the other GPUs (b2r_shared_open), written over NVLink.  After every part has run
 * (and the caller has ordered the frames, e.g. with a barrier) each surface holds the full frame.  Not available
 * with depth of field (B2R_E_UNSUPPORTED).  d_pixelColours etc. are optional full-frame arrays (local rows only). */
int b2r_rt_frame_split_device_async(b2r_ctx* ctx, int part, int nparts, uint32_t* const* d_surfaces, int n,
                                    float* d_pixelColours, b2r_intersection* d_closestIntersections,
                                    float* d_focalDistances);
/* Gather form of the split: this part's pixels go to ONE surface, normally the root GPU's (local on the root,
 * peer-mapped elsewhere), so the bytes a part sends fall as 1/nparts.  When d_arrive is not NULL, the last thread
 * block of the launch adds 1 to that 32-bit word -- usually in the root's memory, over NVLink -- after all of the
 * launch's stores are visible system-wide; the root orders frames with b2r_stream_wait_value32(word, parts * frames)
 * on its own stream: no collective and no host round trip. */
int b2r_rt_frame_gather_device_async(b2r_ctx* ctx, int part, int nparts, uint32_t* d_root_surface, uint32_t* d_arrive);
/* Makes the context's stream wait (on the GPU front end) until *d_word >= value. */
int b2r_stream_wait_value32(b2r_ctx* ctx, const uint32_t* d_word, uint32_t value);
/* Host side of a split frame: draws tile rows part, part + nparts, ... and copies exactly those rows into the
 * caller's FULL-FRAME host surface (width*height uint32; page-lock it once with b2r_pin_host_buffer).  Each GPU of
 * a split thus ships height/nparts rows over its own PCIe link.  The _async form returns after enqueueing. */
int b2r_rt_frame_part(b2r_ctx* ctx, int part, int nparts, uint32_t* surface);
int b2r_rt_frame_part_async(b2r_ctx* ctx, int part, int nparts, uint32_t* surface);
/* Whole raytracer frame straight to the 24-bit BMP payload (b2r_bmp_payload_bytes; what SDL_SaveBMP writes,
 * raytracer.cpp:175): 3 bytes per pixel cross the bus instead of 4.  bgr: host memory.  Returns after enqueueing. */
int b2r_rt_frame_bgr8_async(b2r_ctx* ctx, uint8_t* bgr);
/* Rows [y0,y1) of the rasteriser's Draw() into the caller's full-frame host surface.  Returns after enqueueing;
 * b2r_synchronize also reports B2R_E_CAPACITY. */
int b2r_ras_frame_part_async(b2r_ctx* ctx, int y0, int y1, uint32_t* surface);
int b2r_ras_frame_device_async(b2r_ctx* ctx, int y0, int y1, uint32_t* d_surface, float* d_depthBuffer,
                               float* d_pixelColours, float* d_focalDistances, int32_t* d_winnerIndex);
/* Resolve rows [y0,y1) of d_pixelColours (+ d_focalDistances when DOF is on) into d_surface. */
int b2r_resolve_surface_device_async(b2r_ctx* ctx, int y0, int y1, const float* d_pixelColours,
                                     const float* d_focalDistances, uint32_t* d_surface);

/* ---- introspection ------------------------------------------------------ */
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
unsigned long long b2r_launch_count(const b2r_ctx* ctx);
/* Counters of the last draw: see B2R_STAT_* indices; out must hold B2R_STAT_COUNT values. */
enum {
    B2R_STAT_PRIMARY_RAYS = 0,  /* ClosestIntersection calls from Draw (raytracer.cpp:580) */
    B2R_STAT_SHADOW_RAYS = 1,   /* ClosestIntersection calls the reference makes from DirectLight (raytracer.cpp:310) for this draw */
    B2R_STAT_EXACT_TESTS = 2,   /* ray/triangle pairs that reached the exact (reference-order) evaluation */
    B2R_STAT_RAS_TRIANGLES = 3, /* triangles drawn (not culled) */
    B2R_STAT_RAS_ROWS = 4,      /* polygon rows produced by ComputePolygonRows */
    B2R_STAT_RAS_DEPTH_TESTS = 5, /* on-screen depth tests (rasteriser.cpp:606) */
    B2R_STAT_SHADOW_RAYS_EVALUATED = 6, /* of SHADOW_RAYS, those whose DirectLight term was computed; the others belong to
                                           sub-samples whose hit did not replace the pixel's carried Intersection
                                           (raytracer.cpp:243): DirectLight has the same argument as for the sub-sample
                                           before and its value is reused, bit for bit */
    B2R_STAT_COUNT = 8
};
int b2r_get_stats(b2r_ctx* ctx, unsigned long long* out);
/* Counters cost atomics, so they are off by default; when on, every draw also fills them. */
int b2r_enable_stats(b2r_ctx* ctx, int on);
/* Tuning/diagnostic switches (B2R_OPT_*); results are identical for every setting. */
enum {
    B2R_OPT_RT_FILTER = 0,   /* 1 (default): conservative FMA filter ahead of the exact reference-order test; 0: exact test on every ray/triangle pair */
    B2R_OPT_RT_VARIANT = 1,  /* 0 all culling levels (default); 1 per-ray filter only; 2 force the large-scene path; 3 no shadow-candidate cache; 4 host-buffer draws use one launch per sub-band instead of stream wait-value operations; 5 DirectLight evaluated for every hit sub-sample instead of once per carried Intersection */
    B2R_OPT_RAS_VARIANT = 2, /* 0 sort-last pipeline (default); 1 the same, never using the fixed-slot (no readback) large-triangle path of small scenes; 2 screen-tile pipeline (binning + per-tile raster/shade in shared memory); 3 the same without fixed slots */
    B2R_OPT_DOF_VARIANT = 3  /* 0 shared-memory tiled kernel when dofKernelSize == 8 (default); 1 always the generic kernel */
};
int b2r_set_option(b2r_ctx* ctx, int option, int value);
/* Device FP32 FFMA throughput microbenchmark (TFLOP/s), the raytracer's roofline denominator. */
int b2r_measure_fp32_peak(b2r_ctx* ctx, double* tflops, double* seconds);
/* Self-test of the rasteriser's shared-reciprocal IEEE division (several a/b with one b, used for pos/pos.z of
 * VertexShader rasteriser.cpp:538-541, the step divisions of Interpolate :622-624 and Bresenham :648-649, and
 * pos3d/zinv of PixelShader :557): n pseudo-random operand pairs (all exponents, zeros, denormals, infinities, NaN,
 * small integers) against div.rn.f32, bit for bit.  *mismatches must come back 0; firstBad4 (optional) receives
 * a, b, expected, got of the first mismatch. */
int b2r_selftest_division(b2r_ctx* ctx, unsigned long long n, unsigned seed, unsigned long long* mismatches, float* firstBad4);

/* ---- single-frame split across GPUs: resolve fused with the band exchange ---- */
/* One process per GPU.  Each rank renders its row band (y0,y1 of the draw calls) and resolves it
 * DIRECTLY into the surface buffers of its peers with stores over NVLink (peer-mapped pointers), so
 * the gather of framebuffer bands needs no separate collective -- only a barrier afterwards.
 *   b2r_shared_alloc        cudaMalloc on this context's GPU + a 64-byte inter-process handle
 *   b2r_shared_open/close   map a peer's allocation into this process (cudaIpcOpenMemHandle)
 *   b2r_resolve_surface_multi_device_async
 *                           rows [y0,y1): one kernel, every pixel written to each of n (<= 8)
 *                           destination surfaces (local or peer-mapped) */
#define B2R_IPC_HANDLE_BYTES 64
#define B2R_MAX_PEERS 8
int b2r_shared_alloc(b2r_ctx* ctx, size_t bytes, void** d_ptr, void* handle_out /* 64 bytes */);
int b2r_shared_free(b2r_ctx* ctx, void* d_ptr);
int b2r_shared_open(b2r_ctx* ctx, const void* handle /* 64 bytes */, void** d_ptr);
int b2r_shared_close(b2r_ctx* ctx, void* d_ptr);
/* Device-to-device copy on the context's stream (e.g. out of a shared buffer into caller-owned memory). */
int b2r_copy_device_async(b2r_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
int b2r_resolve_surface_multi_device_async(b2r_ctx* ctx, int y0, int y1, const float* d_pixelColours,
                                           const float* d_focalDistances, uint32_t* const* d_surfaces, int n);

/* ---- several GPUs behind one Draw(): a group of contexts in ONE process ------------------------------ */
/* The reference's Draw() parallelises image rows with OpenMP (raytracer.cpp:557); a group does the same over GPUs.
 * b2r_group_create makes one context per listed device (same screen size); scene and frame params are replicated.
 *   b2r_group_rt_frame    Draw() of the raytracer for one frame: device i traces tile rows i, i+n, ... (interleaved,
 *                         so every device carries the same mix of rows) and copies its rows into the caller's host
 *                         surface over its own PCIe link.  No collective: the host surface is the meeting point.
 *   b2r_group_ras_frame   Draw() of the rasteriser, sort-first: the triangle list is replicated, device i draws the
 *                         i-th contiguous row band.
 *   b2r_group_rt_frames   an animation: frame f of `frames` is rendered whole by device f mod n;
 *                         each finished frame is either copied to surfaces + f*width*height (32-bit XRGB) or, when
 *                         bmp_pattern is given (a printf pattern with one %d), written as the 24-bit BMP
 *                         SDL_SaveBMP would write (raytracer.cpp:175) by writer threads while the GPUs go on.
 *                         For the duration of the call every device holds a ring of up to 8 page-locked frame
 *                         buffers (2 frames on the GPU, the others being written; fewer if page-locked memory is short).
 * A group is driven by one host thread.  Errors: negative B2R_E_* code, text from b2r_group_last_error. */
typedef struct b2r_group b2r_group;
int b2r_group_create(b2r_group** out, const int* devices, int n, int width, int height);
int b2r_group_destroy(b2r_group* g);
int b2r_group_size(const b2r_group* g);
b2r_ctx* b2r_group_ctx(b2r_group* g, int i);  /* member context i (options, statistics); owned by the group */
const char* b2r_group_last_error(const b2r_group* g);
int b2r_group_set_triangles(b2r_group* g, const void* triangles, int count, int stride_bytes);
int b2r_group_set_frame(b2r_group* g, const b2r_frame_params* params);
int b2r_group_rt_frame(b2r_group* g, uint32_t* surface);
int b2r_group_ras_frame(b2r_group* g, uint32_t* surface);
int b2r_group_rt_frames(b2r_group* g, const b2r_frame_params* frames, int nframes, uint32_t* surfaces,
                        const char* bmp_pattern);

/* ---- sub-stage entry points ------------------------------------------------------------------------- */
/* The reference's callee functions (raytracer.cpp:105-107, rasteriser.cpp:87-94) on caller-provided inputs:
 * batched, host pointers, synchronous; each runs exactly that stage in reference-order arithmetic with the current
 * scene (b2r_set_triangles) and frame params (b2r_set_frame).  Pixel = the reference's 24-byte struct
 * {int x, y; float zinv; vec3 pos3d} (rasteriser TestModel.h:34-53).
 *   ClosestIntersection(start, dir, triangles, closest&, isLight, x, y) -> bool   raytracer.cpp:202-257
 *       io[k] is the running closest intersection (in/out); hit[k] the return value; focal[k] the value the call
 *       would store in focalDistances (0 if it stored none); isLight may be NULL (all false).
 *   DirectLight(const Intersection&) -> vec3                                       raytracer.cpp:265-327
 *   VertexShader(const Vertex&, Pixel&)                                            rasteriser.cpp:532-546
 *   Interpolate(Pixel a, Pixel b, vector<Pixel>& result)  (result.size() == n)     rasteriser.cpp:615-637
 *   ComputePolygonRows(vertexPixels[3], left&, right&)    (*rows = ROWS)           rasteriser.cpp:674-735
 *   PixelShader(const Pixel&, color, normal)  -> the pixelColours / focalDistances values it would store  :549-589 */
int b2r_rt_closest_intersection_batch(b2r_ctx* ctx, int n, const float* starts3, const float* dirs3,
                                      const int32_t* isLight, b2r_intersection* io, int32_t* hit, float* focal);
int b2r_rt_direct_light_batch(b2r_ctx* ctx, int n, const b2r_intersection* hits, float* out3);
int b2r_ras_vertex_shader_batch(b2r_ctx* ctx, int n, const float* verts3, void* pixels24);
int b2r_ras_interpolate(b2r_ctx* ctx, const void* a24, const void* b24, int n, void* out24);
int b2r_ras_compute_polygon_rows(b2r_ctx* ctx, const void* vertexPixels3x24, void* left24, void* right24, int maxRows,
                                 int* rows);
int b2r_ras_pixel_shader_batch(b2r_ctx* ctx, int n, const void* pixels24, const float* colors3, const float* normals3,
                               float* outColours3, float* outFocal);

/* ---- host-side scene helpers (no GPU involved) ---------------------------- */
/* LoadTestModel (raytracer TestModel.h:51-192 == rasteriser TestModel.h:151-292): the 30-triangle
 * Cornell box, written as reference Triangle records of the given stride (60 or 64).  Returns the
 * triangle count or a negative error. */
int b2r_scene_cornell_box(void* out, int capacity, int stride_bytes);
/* Uniform k*k tessellation of every input triangle, parent order kept:
 * P(i,j) = A + (i/k)(B-A) + (j/k)(C-A); per j then i: "up" (P(i,j),P(i+1,j),P(i,j+1)) and, if
 * i+j < k-1, "down" (P(i+1,j),P(i+1,j+1),P(i,j+1)); normals recomputed like the Triangle ctor
 * (TestModel.h:26-31).  out may be NULL to query the count.  Returns the output count. */
long long b2r_scene_tessellate(const void* in, int count, int in_stride, int k, void* out, int out_stride);
/* ASCII STL mesh with the reference loader's semantics (rasteriser/Source/LoadSTL.cpp:17-81: vertices after every
 * "outer" line, (float)atof, all coordinates * -0.05f, colour 0.5, normal recomputed).  out may be NULL to count.
 * Returns the facet count, or a negative error (B2R_E_IO, B2R_E_CAPACITY). */
long long b2r_scene_load_stl(const char* path, void* out, long long capacity, int stride_bytes);
/* cameraRot as Update() builds it from yaw (raytracer.cpp:377-382, rasteriser.cpp:378-383);
 * rot11 is the preset [1][1] element: 1.0f (raytracer.cpp:162) or 1.01f (rasteriser.cpp:115). */
int b2r_camera_rot_from_yaw(float yaw, float rot11, float* rot9_colmajor);
/* Orbit animation camera: yaw = frame*2pi/nframes,
 * cameraPos = -radius*forward, rotation as above with rot11 = 1. */
int b2r_orbit_camera(int frame, int nframes, float radius, float* cameraPos3, float* rot9_colmajor);
/* The soft-shadow jitter table AddLight() derives from glibc rand() after srand(seed)
 * (raytracer.cpp:186-190,260-263) for light 0 at lightPos; entries 16..255 are 0. */
int b2r_jitter_table(unsigned seed, const float* lightPos3, float* out768);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H */
