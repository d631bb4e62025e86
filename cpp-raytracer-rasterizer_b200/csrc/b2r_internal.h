// b2r_internal.h -- shared between the C-ABI layer (b2r_api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/b2r.h"

// Debug build (nvcc -DB2R_DEBUG_BOUNDS, __graft_entry__.build_debug(): libb2r_dbg.so): every index into a scheduler
// word, band counter, per-warp list, row-record slot or tile array is checked; a violation prints the site and traps,
// which the host sees as a CUDA error.  compute-sanitizer is not available on the GPU pool, so the randomised soak is
// run once per round against this build instead (profiles/r02_soak_debug_bounds.log).  Release builds: no code.
#ifdef B2R_DEBUG_BOUNDS
#include <stdio.h>
#define B2R_BOUND(i, n)                                                                                          \
    do {                                                                                                         \
        if (!((unsigned long long)(long long)(i) < (unsigned long long)(long long)(n))) {                        \
            printf("B2R_BOUND %s:%d: index %lld not in [0, %lld)\n", __FILE__, __LINE__, (long long)(i), (long long)(n)); \
            __trap();                                                                                            \
        }                                                                                                        \
    } while (0)
#else
#define B2R_BOUND(i, n) ((void)0)
#endif

namespace b2r {

// ---------------------------------------------------------------------------
// Per-frame constants, laid out once by b2r_set_frame() on the host (exact
// reference-order arithmetic where bits matter) and uploaded to HBM in one
// copy.  Every vec3 is padded to float4 so the kernels can use 128-bit loads.
// ---------------------------------------------------------------------------
constexpr int kMaxOrigins = 1 + B2R_RANDOM_POSITIONS;  // camera + every light sample

struct DevFrame {
    float cam[4];        // cameraPos
    float R[12];         // cameraRot, column-major, 9 used
    float Rinv[12];      // glm::inverse(cameraRot) (rasteriser.cpp:559), 9 used
    float focal;         // focalLength
    float dofFocal;      // FOCAL_LENGTH
    float primaryDmax;   // bound on |cameraRot*d|_inf over the frame (filter margin)
    int   nLights;
    int   samples;       // shadow samples per light (1 or SOFT_SHADOWS_SAMPLES)
    int   aaN;           // sub-samples per axis (1 or AA_SAMPLES)
    int   nOrigins;      // 1 + nLights*samples
    int   dofEnabled;
    int   dofKernel;
    int   pad0[3];
    float indirect[4];
    float reflectance[4];
    float lightPos[B2R_MAX_LIGHTS][4];
    float lightColor[B2R_MAX_LIGHTS][4];    // color*intensity                 (rasteriser.cpp:577)
    float lightPower[B2R_MAX_LIGHTS][4];    // (color*intensity)/float(samples) (raytracer.cpp:282,296)
    float origin[kMaxOrigins][4];           // [0] camera; [1 + k*samples + s] light sample position
};

// Scene-static per-triangle record for the raytracer: 5 float4 = 80 bytes.
//   q0 = (v0.x, v0.y, v0.z, e1.x)   q1 = (e1.y, e1.z, e2.x, e2.y)
//   q2 = (e2.z, n.x, n.y, n.z)      n  = cross(e1,e2)           (raytracer.cpp:216-217,225)
//   q3 = (nh.x, nh.y, nh.z, col.r)  nh = normalize(normal)      (raytracer.cpp:300)
//   q4 = (col.g, col.b, 0, 0)
constexpr int kGeomQuads = 5;
// Per (ray origin, triangle): 2 float4 of exact constants  (be2.xyz, e1e2b), (e1b.xyz, 0)
// (raytracer.cpp:218,226-227,231) and 3 float4 of conservative filter forms (see rt_trace.cu).

// Frame constants the raytracer's per-sample code needs, passed by value in the kernel-parameter bank so they live
// in uniform registers instead of per-thread registers (the light-sample origins stay in DevFrame).
struct RtFrame {
    float cam[3], focal;
    float R[9], dofFocal;
    float Rf[3];       // cameraRot column 2 * focalLength, the third product of cameraRot * vec3(dx, dy, focalLength)
    float indirect[3];
    float light0[3], power0[3];  // single-light variant: origin[1] and lightPower[0]
    int aaN, nLights, samples, nOrigins;
};

struct RtLaunch {
    RtFrame fr;
    const float4* geom;      // T * kGeomQuads
    const float4* xconst;    // large scenes only: nO * T * 2 exact (origin,triangle) constants in HBM
    const float4* fconst;    // large scenes only: nO * T * 3 filter forms in HBM
    const DevFrame* frame;
    int T;
    int W, H, y0, y1;
    int tilesX, numTiles;
    float* colours;                     // may be null
    b2r_intersection* closest;          // may be null
    float* focal;                       // may be null
    uint32_t* surface;                  // may be null: the resolved XRGB surface (only without depth of field)
    uint32_t* peerSurface[B2R_MAX_PEERS - 1];  // further copies of the surface, e.g. peer-mapped buffers of other GPUs
    int nPeerSurfaces;                  // (single-frame multi-GPU split: the exchange is part of the trace kernel)
    int tileRowStride, tileRowOffset;   // this launch draws tile rows offset, offset + stride, ... (1, 0 = all)
    unsigned long long* stats;          // device counters (B2R_STAT_*), null when stats are off
    unsigned* bandDone;  // may be null: per sub-band, the number of finished warp tiles (host-buffer draw: the copy
                         // stream waits on these words and copies a sub-band out while the kernel is still tracing)
    int bandTileRows;    // tile rows (8 pixel rows each) per sub-band
    unsigned* arrive;     // may be null: word (usually in another GPU's memory) that the last CTA of this launch increments
                          // once all of the launch's stores are visible system-wide (single-frame split: gather to a root)
    unsigned* arriveCtr;  // local word, zero between launches: CTAs that have finished (used with arrive)
    unsigned* sched;   // 2 words, zero between launches: next warp tile to hand out, warps that have finished
    int batch;         // warp tiles per scheduler fetch (set by the launcher)
    float rcpTilesX;   // 1.0f / tilesX (tile index -> tile row without an integer division)
    float rcpNN;       // 1.0f / (aaN * aaN), used in place of the division by N*N when aaN is a power of two (same bits)
    int tileOrder;     // 0: tiles handed out top to bottom (needed when sub-bands are copied out in row order); 1: bottom to top
    int useFilter;
    int reuseLight;    // 1 (default): DirectLight only when the pixel's carried Intersection changed (see rt_trace.cu)
    int shadowCache;   // set by launch_rt_trace_shade: per-warp shadow-candidate cache in shared memory
};

// Frame constants the rasteriser kernels need, passed by value in the kernel-parameter bank.
struct RasFrame {
    float cam[3], focal;
    float R[9], dofFocal;
    float Rinv[9];
    float halfW, halfH;       // SCREEN_WIDTH / 2.0f, SCREEN_HEIGHT / 2.0f (rasteriser.cpp:544-545)
    int nLights;
    float reflectance[3], indirect[3];
    float lightPos[B2R_MAX_LIGHTS][3];
    float lightColor[B2R_MAX_LIGHTS][3];  // color*intensity (rasteriser.cpp:577)
};

struct RasLaunch {
    const unsigned char* raw;  // reference Triangle records, 64 bytes each (60-byte scenes are repacked at upload)
    int stride;                // 64
    const unsigned char* culled;  // one byte per triangle (may be null => nothing culled)
    int T;
    const DevFrame* frame;
    RasFrame fr;
    int W, H, y0, y1;
    float* depth;
    float* colours;
    float* focal;
    int32_t* winner;
    uint32_t* surface;         // may be null: the resolved XRGB surface (only without depth of field)
    unsigned long long* stats;
    int bandSlots;             // large-triangle path with fixed-capacity slots (set by launch_ras_draw)
};

struct Ctx;
constexpr int kFrameRing = 8;

// kernels (each returns cudaGetLastError() of its launches)
cudaError_t launch_tri_prep(Ctx* c, cudaStream_t s);
cudaError_t launch_rt_trace_shade(Ctx* c, const RtLaunch& a, cudaStream_t s);
cudaError_t launch_ras_draw(Ctx* c, const RasLaunch& a, cudaStream_t s);
cudaError_t launch_ras_cull(Ctx* c, unsigned char* d_culled, cudaStream_t s);
cudaError_t launch_ras_repack(Ctx* c, cudaStream_t s);  // 64-byte copy of a 60-byte scene for the rasteriser's 128-bit loads
// After the stream has been synchronised: cudaErrorInvalidValue if a draw since the last call hit the row/coordinate
// limits (-> B2R_E_CAPACITY), cudaSuccess otherwise.
cudaError_t ras_take_error(Ctx* c);
cudaError_t launch_resolve_surface(Ctx* c, int y0, int y1, const float* d_colours, const float* d_focal,
                                   uint32_t* d_surface, cudaStream_t s);
cudaError_t launch_resolve_surface_multi(Ctx* c, int y0, int y1, const float* d_colours, const float* d_focal,
                                         uint32_t* const* d_surfaces, int n, cudaStream_t s);
cudaError_t launch_surface_to_bgr8(Ctx* c, const uint32_t* d_surface, uint8_t* d_bgr, cudaStream_t s);
cudaError_t run_fp32_peak(Ctx* c, double* tflops, double* seconds);

// Growable device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct Ctx {
    int device = 0;
    int W = 0, H = 0;
    int smCount = 148;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;   // host-buffer draws: D2H of one sub-band overlaps the next sub-band's kernels
    cudaEvent_t partDone[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // scene
    int T = 0;
    int stride = 0;
    DevBuf raw;      // reference Triangle records as given
    DevBuf raw64;    // the same as 64-byte records when the caller's stride is 60 (rasteriser kernels use 128-bit loads)
    DevBuf culled;   // T bytes
    DevBuf geom;     // T * 80 bytes (raytracer)
    bool haveScene = false;

    // frame
    b2r_frame_params params{};
    DevFrame hostFrame{};
    DevFrame* pinnedFrame = nullptr;      // kFrameRing pinned staging slots
    cudaEvent_t frameUploaded[8] = {};
    int frameSlot = 0;
    DevBuf frame;
    bool haveFrame = false;

    // outputs kept on the device between draw and resolve / host copies
    DevBuf colours, closest, focal, depth, winner, surface, bgr;
    // rasteriser intermediates
    DevBuf rasTri, rasRows, rasRefs, rasScratch, rasJobs, rasPartials;
    DevBuf rasSLScratch, rasKeys, rasSmall;  // sort-last pipeline: counters + scan scratch, key buffer, small-triangle row records  // triangle words + large-triangle records, edge samples, tile lists, counters
    DevBuf subScratch;  // staging of the sub-stage entry points
    struct KernelInfo {
        const void* fn;
        size_t smem;
        int perSM;
    };
    std::vector<KernelInfo> rtKernelCache;  // raytracer variants already configured (shared-memory opt-in, CTAs/SM)
    DevBuf rtSched;   // raytracer: warp-tile scheduler words (RtLaunch::sched), then the band counters (::bandDone)
    void* waitValue32 = nullptr;  // cuStreamWaitValue32 when the driver offers stream memory operations
    bool memOpsProbed = false;
    DevBuf rtX, rtF;  // raytracer, scenes too large for shared memory: per-frame (origin,triangle) constants
    size_t rasKeysClean = 0;          // bytes of rasKeys known to be zero (left so by the last shade pass)
    void* rasKeysCleanPtr = nullptr;
    bool rasSmallAttr = false;        // ras_small's shared-memory opt-in has been set on this context's device
    bool rasSLCtrDirty = true;        // the sort-last pipeline's per-frame counters need a clear before the next draw
    void* rasErrCtr = nullptr;        // counters (RasCounters) of the pipeline whose capacity flag is pending
    bool rasTileAttr = false;         // ras_tile's shared-memory opt-in has been set on this context's device
    size_t rasTilesClean = 0;         // tile counts known to be zero for a grid of this many tiles (left so by ras_tile)
    // Buffer sizes of the rasteriser's large-scene path, read back once per (scene, culling flags, frame params, band):
    // rasGen counts the changes of the first three.
    unsigned long long rasGen = 1;
    struct RasSizes {
        bool valid = false;
        unsigned long long gen = 0;
        int y0 = 0, y1 = 0;
        unsigned nBig = 0, bigRows = 0, bigSamples = 0, totalRefs = 0;
    } rasSizes, rasSizesSL;  // screen-tile pipeline, sort-last pipeline
    // pinned staging for host-pointer entry points
    void* pinned = nullptr;
    size_t pinnedCap = 0;

    DevBuf stats;
    bool statsOn = false;
    unsigned long long launches = 0;
    int optRtFilter = 1, optRtVariant = 0, optRasVariant = 0, optDofVariant = 0;
    int lastDraw = -1;  // 0 raytracer, 1 rasteriser
    bool rasCtrDirty = true;     // the rasteriser's per-frame counters need a clear before the next draw
    bool rasErrPending = false;  // an asynchronous rasteriser draw left its capacity flag to be checked (see ras_take_error)
    // what the context's own buffers hold after the last host-buffer draw: the fused raytracer frame may leave only
    // the resolved surface (no pixelColours); b2r_resolve_* then start from it
    bool coloursValid = false, surfaceValid = false;
};

}  // namespace b2r
