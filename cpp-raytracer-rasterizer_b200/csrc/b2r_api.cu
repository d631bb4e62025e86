// b2r_api.cu -- the extern "C" layer declared in include/b2r.h.
//
// Owns the context (device buffers, stream, per-frame constants) and maps the
// reference's implicit Draw() contract (globals in, globals out) onto explicit
// calls.  No CPU rendering path exists here: every draw is a CUDA launch.
#include <float.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "b2r_internal.h"
#include "exact.cuh"

using namespace b2r;

namespace {

thread_local std::string g_createError;

int fail(Ctx* c, int code, const char* what) {
    if (c) c->err = what;
    return code;
}

int cuda_fail(Ctx* c, cudaError_t e, const char* what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    if (c) c->err = buf;
    else g_createError = buf;
    cudaGetLastError();  // clear the sticky-less error state
    return B2R_E_CUDA;
}

#define CU(call, what)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(c, e__, what); \
    } while (0)

int bind(Ctx* c) {
    if (!c) return B2R_E_INVALID;
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaSetDevice");
    return B2R_OK;
}

// glm::inverse(mat3) in GLM's operation order (glm/detail/type_mat3x3.inl:36-57); m, out column-major.
void inverse3(const float* m, float* o) {
#define M(c, r) m[3 * (c) + (r)]
    const float ood = xdiv(1.0f, xadd(xsub(xmul(+M(0, 0), xsub(xmul(M(1, 1), M(2, 2)), xmul(M(2, 1), M(1, 2)))),
                                           xmul(M(1, 0), xsub(xmul(M(0, 1), M(2, 2)), xmul(M(2, 1), M(0, 2))))),
                                      xmul(M(2, 0), xsub(xmul(M(0, 1), M(1, 2)), xmul(M(1, 1), M(0, 2))))));
    o[3 * 0 + 0] = xmul(+xsub(xmul(M(1, 1), M(2, 2)), xmul(M(2, 1), M(1, 2))), ood);
    o[3 * 1 + 0] = xmul(-xsub(xmul(M(1, 0), M(2, 2)), xmul(M(2, 0), M(1, 2))), ood);
    o[3 * 2 + 0] = xmul(+xsub(xmul(M(1, 0), M(2, 1)), xmul(M(2, 0), M(1, 1))), ood);
    o[3 * 0 + 1] = xmul(-xsub(xmul(M(0, 1), M(2, 2)), xmul(M(2, 1), M(0, 2))), ood);
    o[3 * 1 + 1] = xmul(+xsub(xmul(M(0, 0), M(2, 2)), xmul(M(2, 0), M(0, 2))), ood);
    o[3 * 2 + 1] = xmul(-xsub(xmul(M(0, 0), M(2, 1)), xmul(M(2, 0), M(0, 1))), ood);
    o[3 * 0 + 2] = xmul(+xsub(xmul(M(0, 1), M(1, 2)), xmul(M(1, 1), M(0, 2))), ood);
    o[3 * 1 + 2] = xmul(-xsub(xmul(M(0, 0), M(1, 2)), xmul(M(1, 0), M(0, 2))), ood);
    o[3 * 2 + 2] = xmul(+xsub(xmul(M(0, 0), M(1, 1)), xmul(M(1, 0), M(0, 1))), ood);
#undef M
}

int ensure_pinned(Ctx* c, size_t bytes) {
    if (bytes <= c->pinnedCap) return B2R_OK;
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->pinnedCap = 0;
    CU(cudaMallocHost(&c->pinned, bytes), "cudaMallocHost");
    c->pinnedCap = bytes;
    return B2R_OK;
}

int reset_stats(Ctx* c) {
    if (!c->statsOn) return B2R_OK;
    CU(c->stats.reserve(sizeof(unsigned long long) * B2R_STAT_COUNT), "stats alloc");
    CU(cudaMemsetAsync(c->stats.p, 0, sizeof(unsigned long long) * B2R_STAT_COUNT, c->stream), "stats clear");
    return B2R_OK;
}

int check_band(Ctx* c, int y0, int y1) {
    if (y0 < 0 || y1 > c->H || y0 > y1) return fail(c, B2R_E_INVALID, "row band outside [0,H]");
    if (!c->haveScene) return fail(c, B2R_E_NO_SCENE, "b2r_set_triangles has not been called");
    if (!c->haveFrame) return fail(c, B2R_E_NO_SCENE, "b2r_set_frame has not been called");
    return B2R_OK;
}

const char* const kRasCapacityText =
    "rasteriser: a projected triangle exceeds 2^22 rows or 2^24 pixels in a coordinate "
    "(a vertex at or behind the camera plane); the reference cannot draw it either";

// Rasteriser draws of small scenes are plain launch sequences without a readback; their capacity flag is looked at
// when the caller synchronises (host-buffer calls do so themselves, asynchronous callers through b2r_synchronize).
int ras_deferred_error(Ctx* c) {
    const cudaError_t e = ras_take_error(c);
    if (e == cudaErrorInvalidValue) return fail(c, B2R_E_CAPACITY, kRasCapacityText);
    if (e != cudaSuccess) return cuda_fail(c, e, "rasteriser error flag");
    return B2R_OK;
}

// D2H of rows [y0,y1) of a full-frame array with `bpp` bytes per pixel.
int copy_rows_out(Ctx* c, void* host, const void* dev, int y0, int y1, size_t bpp) {
    if (!host || y1 <= y0) return B2R_OK;
    const size_t off = (size_t)y0 * c->W * bpp, bytes = (size_t)(y1 - y0) * c->W * bpp;
    CU(cudaMemcpyAsync((char*)host + off, (const char*)dev + off, bytes, cudaMemcpyDeviceToHost, c->stream), "D2H copy");
    return B2R_OK;
}

}  // namespace

extern "C" {

int b2r_abi_version(void) { return B2R_ABI_VERSION; }

const char* b2r_last_error(const b2r_ctx* ctx) {
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    return c ? c->err.c_str() : g_createError.c_str();
}

int b2r_create(b2r_ctx** out, int device, int width, int height) {
    Ctx* c = nullptr;
    if (!out || width <= 0 || height <= 0 || width > 32768 || height > 32768) {
        g_createError = "b2r_create: bad arguments";
        return B2R_E_INVALID;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return cuda_fail(nullptr, e == cudaSuccess ? cudaErrorNoDevice : e,
                         "b2r_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= count) {
        g_createError = "b2r_create: device index out of range";
        return B2R_E_INVALID;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    c = new (std::nothrow) Ctx();
    if (!c) return B2R_E_CUDA;
    c->device = device;
    c->W = width;
    c->H = height;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete c;
        return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    }
    c->smCount = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    c->stream = c->ownStream;
    if ((e = cudaMallocHost((void**)&c->pinnedFrame, sizeof(DevFrame) * kFrameRing)) != cudaSuccess ||
        (e = cudaMallocHost(&c->pinned, 4096)) != cudaSuccess || (e = c->frame.reserve(sizeof(DevFrame))) != cudaSuccess) {
        b2r_destroy(reinterpret_cast<b2r_ctx*>(c));
        return cuda_fail(nullptr, e, "b2r_create: allocation");
    }
    c->pinnedCap = 4096;
    *out = reinterpret_cast<b2r_ctx*>(c);
    return B2R_OK;
}

int b2r_destroy(b2r_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_OK;
    cudaSetDevice(c->device);
    if (c->ownStream) cudaStreamSynchronize(c->ownStream);
    DevBuf* bufs[] = {&c->raw, &c->culled, &c->geom, &c->frame, &c->colours, &c->closest, &c->focal, &c->depth,
                      &c->winner, &c->surface, &c->bgr, &c->rasTri, &c->rasRows, &c->rasRefs, &c->rasScratch, &c->rasJobs, &c->rasPartials, &c->raw64, &c->rasSLScratch, &c->rasKeys, &c->rasSmall, &c->rtX, &c->rtF, &c->rtSched, &c->subScratch, &c->stats};
    for (DevBuf* b : bufs) b->release();
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->pinnedFrame) cudaFreeHost(c->pinnedFrame);
    for (int i = 0; i < kFrameRing; ++i)
        if (c->frameUploaded[i]) cudaEventDestroy(c->frameUploaded[i]);
    if (c->copyStream) {
        cudaStreamDestroy(c->copyStream);
        for (int i = 0; i < 4; ++i) cudaEventDestroy(c->partDone[i]);
    }
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
    return B2R_OK;
}

int b2r_set_stream(b2r_ctx* ctx, void* cuda_stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->ownStream;
    return B2R_OK;
}

void* b2r_get_stream(b2r_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    return c ? reinterpret_cast<void*>(c->stream) : nullptr;
}

int b2r_synchronize(b2r_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    CU(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
    return ras_deferred_error(c);
}

int b2r_set_triangles(b2r_ctx* ctx, const void* triangles, int count, int stride) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (count < 0 || (count > 0 && !triangles) || (stride != 60 && stride != 64))
        return fail(c, B2R_E_INVALID, "b2r_set_triangles: stride must be 60 (raytracer Triangle) or 64 (rasteriser Triangle)");
    CU(cudaStreamSynchronize(c->stream), "sync before scene upload");
    const size_t bytes = (size_t)count * stride;
    CU(c->raw.reserve(bytes + 64), "scene alloc");
    CU(c->culled.reserve((size_t)count + 64), "scene alloc");
    CU(c->geom.reserve((size_t)count * kGeomQuads * 16 + 64), "scene alloc");
    if (count) CU(cudaMemcpyAsync(c->raw.p, triangles, bytes, cudaMemcpyHostToDevice, c->stream), "scene upload");
    c->T = count;
    c->stride = stride;
    c->rasGen++;
    // isCulled: byte 60 of the 64-byte rasteriser Triangle; the raytracer Triangle has none
    if (count) {
        if (int rc = ensure_pinned(c, (size_t)count + 64)) return rc;
        unsigned char* m = (unsigned char*)c->pinned;
        for (int i = 0; i < count; ++i) m[i] = (stride == 64) ? (((const unsigned char*)triangles)[(size_t)i * 64 + 60] != 0) : 0;
        CU(cudaMemcpyAsync(c->culled.p, m, (size_t)count, cudaMemcpyHostToDevice, c->stream), "culled upload");
    }
    CU(launch_tri_prep(c, c->stream), "tri_prep_kernel");
    CU(launch_ras_repack(c, c->stream), "ras_repack_kernel");
    CU(cudaStreamSynchronize(c->stream), "scene upload");
    c->haveScene = true;
    return B2R_OK;
}

int b2r_set_culled(b2r_ctx* ctx, const uint8_t* culled, int count) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!c->haveScene || count != c->T || (count > 0 && !culled)) return fail(c, B2R_E_INVALID, "b2r_set_culled: count must equal the triangle count");
    if (count == 0) return B2R_OK;
    CU(cudaStreamSynchronize(c->stream), "sync");
    if (int rc = ensure_pinned(c, (size_t)count + 64)) return rc;
    memcpy(c->pinned, culled, (size_t)count);
    c->rasGen++;
    CU(cudaMemcpyAsync(c->culled.p, c->pinned, (size_t)count, cudaMemcpyHostToDevice, c->stream), "culled upload");
    CU(cudaStreamSynchronize(c->stream), "culled upload");
    return B2R_OK;
}

int b2r_set_frame(b2r_ctx* ctx, const b2r_frame_params* p) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!p) return fail(c, B2R_E_INVALID, "b2r_set_frame: null params");
    if (p->numLights < 0 || p->numLights > B2R_MAX_LIGHTS) return fail(c, B2R_E_INVALID, "numLights outside [0,32]");
    const int N = p->aaEnabled ? p->aaSamples : 1;                       // raytracer.cpp:551-554
    const int samples = p->softShadowsEnabled ? p->softShadowsSamples : 1;  // raytracer.cpp:272-275
    if (N < 1 || N > 64) return fail(c, B2R_E_INVALID, "aaSamples outside [1,64]");
    if (samples < 1) return fail(c, B2R_E_INVALID, "softShadowsSamples < 1");
    if (samples != 1 && p->numLights * samples > B2R_RANDOM_POSITIONS)
        return fail(c, B2R_E_INVALID, "numLights*softShadowsSamples exceeds randomPositions[256] (raytracer.cpp:84)");
    if (p->dofEnabled && (p->dofKernelSize < 1 || p->dofKernelSize > 64)) return fail(c, B2R_E_INVALID, "dofKernelSize outside [1,64]");

    DevFrame& f = c->hostFrame;
    memset(&f, 0, sizeof f);
    for (int i = 0; i < 3; ++i) {
        f.cam[i] = p->cameraPos[i];
        f.indirect[i] = p->indirectLight[i];
        f.reflectance[i] = p->currentReflectance[i];
    }
    for (int i = 0; i < 9; ++i) f.R[i] = p->cameraRot[i];
    inverse3(p->cameraRot, f.Rinv);  // rasteriser.cpp:559 (per pixel in the reference; same bits every time)
    f.focal = p->focalLength;
    f.dofFocal = p->dofFocalLength;
    f.nLights = p->numLights;
    f.samples = samples;
    f.aaN = N;
    f.nOrigins = 1 + p->numLights * samples;
    f.dofEnabled = p->dofEnabled;
    f.dofKernel = p->dofKernelSize;
    // bound on |cameraRot * (x1 - W/2, y1 - H/2, focalLength)|_inf over every sub-sample of the frame
    {
        const double ax = c->W / 2.0 + 2.0, ay = c->H / 2.0 + 2.0, az = fabs((double)p->focalLength);
        double dmax = 0.0;
        for (int r = 0; r < 3; ++r) {
            double v = fabs((double)p->cameraRot[r]) * ax + fabs((double)p->cameraRot[3 + r]) * ay +
                       fabs((double)p->cameraRot[6 + r]) * az;
            if (v > dmax) dmax = v;
        }
        f.primaryDmax = (float)(dmax * 1.0001);
        if (!(f.primaryDmax > 0.f) || !isfinite(f.primaryDmax)) f.primaryDmax = FLT_MAX;  // filter never rejects
    }
    for (int i = 0; i < 3; ++i) f.origin[0][i] = p->cameraPos[i];
    for (int k = 0; k < p->numLights; ++k) {
        const b2r_light& L = p->lights[k];
        for (int i = 0; i < 3; ++i) {
            f.lightPos[k][i] = L.position[i];
            f.lightColor[k][i] = xmul(L.color[i], L.intensity);                 // raytracer.cpp:282, rasteriser.cpp:577
            f.lightPower[k][i] = xdiv(f.lightColor[k][i], (float)samples);      // raytracer.cpp:296
        }
        for (int s = 0; s < samples; ++s) {
            // raytracer.cpp:284-291: the jitter table when soft shadows are on, else the light itself
            const float* src = (samples != 1) ? &p->randomPositions[3 * (k * p->softShadowsSamples + s)] : L.position;
            for (int i = 0; i < 3; ++i) f.origin[1 + k * samples + s][i] = src[i];
        }
    }
    if (!c->haveFrame || memcmp(&c->params, p, sizeof *p) != 0) c->rasGen++;  // same params: same rasteriser buffer sizes
    c->params = *p;
    // pinned staging ring: a slot is reused only after its own upload has completed, so consecutive frames never
    // wait for the GPU (the copy into c->frame is stream-ordered after the kernels that still read the old frame)
    const int slot = c->frameSlot;
    c->frameSlot = (slot + 1) % kFrameRing;
    if (c->frameUploaded[slot]) CU(cudaEventSynchronize(c->frameUploaded[slot]), "wait for the staging slot");
    else CU(cudaEventCreateWithFlags(&c->frameUploaded[slot], cudaEventDisableTiming), "cudaEventCreate");
    DevFrame* stage = c->pinnedFrame + slot;
    memcpy(stage, &f, sizeof f);
    const size_t used = offsetof(DevFrame, origin) + sizeof(float) * 4 * (size_t)f.nOrigins;
    CU(cudaMemcpyAsync(c->frame.p, stage, used, cudaMemcpyHostToDevice, c->stream), "frame upload");
    CU(cudaEventRecord(c->frameUploaded[slot], c->stream), "cudaEventRecord");
    c->haveFrame = true;
    return B2R_OK;
}

int b2r_set_option(b2r_ctx* ctx, int option, int value) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    switch (option) {
        case B2R_OPT_RT_FILTER: c->optRtFilter = value ? 1 : 0; return B2R_OK;
        case B2R_OPT_RT_VARIANT: c->optRtVariant = value; return B2R_OK;
        case B2R_OPT_RAS_VARIANT: c->optRasVariant = value; c->rasGen++; return B2R_OK;
        case B2R_OPT_DOF_VARIANT: c->optDofVariant = value; return B2R_OK;
    }
    return fail(c, B2R_E_INVALID, "unknown option");
}

int b2r_enable_stats(b2r_ctx* ctx, int on) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    c->statsOn = on != 0;
    if (c->statsOn) {
        CU(c->stats.reserve(sizeof(unsigned long long) * B2R_STAT_COUNT), "stats alloc");
        CU(cudaMemsetAsync(c->stats.p, 0, sizeof(unsigned long long) * B2R_STAT_COUNT, c->stream), "stats clear");
    }
    return B2R_OK;
}

int b2r_get_stats(b2r_ctx* ctx, unsigned long long* out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!out) return fail(c, B2R_E_INVALID, "null out");
    memset(out, 0, sizeof(unsigned long long) * B2R_STAT_COUNT);
    if (!c->statsOn || !c->stats.p) return B2R_OK;
    CU(cudaStreamSynchronize(c->stream), "sync");
    CU(cudaMemcpy(out, c->stats.p, sizeof(unsigned long long) * B2R_STAT_COUNT, cudaMemcpyDeviceToHost), "stats D2H");
    return B2R_OK;
}

unsigned long long b2r_launch_count(const b2r_ctx* ctx) {
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    return c ? c->launches : 0ull;
}

int b2r_measure_fp32_peak(b2r_ctx* ctx, double* tflops, double* seconds) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    double t = 0, s = 0;
    CU(run_fp32_peak(c, &t, &s), "fp32_peak_kernel");
    if (tflops) *tflops = t;
    if (seconds) *seconds = s;
    return B2R_OK;
}

// ---- raytracer ------------------------------------------------------------------
constexpr int kMaxCopyBands = 32;

// Optional extras of one raytracer launch: further surface copies and the tile-row interleave of a multi-GPU split.
struct RtSplit {
    uint32_t* const* peers = nullptr;
    int nPeers = 0;
    int stride = 1, offset = 0;
    unsigned* arrive = nullptr;  // incremented once when the launch's pixels have landed (gather to a root)
};

// stream memory operations: cuStreamWaitValue32 through the runtime's loader (absent on some driver set-ups)
static int probe_mem_ops(Ctx* c) {
    if (c->memOpsProbed) return B2R_OK;
    c->memOpsProbed = true;
    cudaDriverEntryPointQueryResult qr;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess && fn)
        c->waitValue32 = fn;
    cudaGetLastError();
    return B2R_OK;
}

static int ensure_copy_stream(Ctx* c) {
    if (c->copyStream) return B2R_OK;
    CU(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking), "cudaStreamCreate (copy)");
    for (int i = 0; i < 4; ++i) CU(cudaEventCreateWithFlags(&c->partDone[i], cudaEventDisableTiming), "cudaEventCreate");
    return B2R_OK;
}

static int rt_launch_band(Ctx* c, int y0, int y1, float* d_col, b2r_intersection* d_clo, float* d_foc,
                          uint32_t* d_surf = nullptr, int bandTileRows = 0, const RtSplit* split = nullptr) {
    c->lastDraw = 0;
    c->coloursValid = c->surfaceValid = false;  // set again by the host-buffer draws once their frame is complete
    if (y1 == y0) return B2R_OK;
    RtLaunch a;
    a.geom = c->geom.as<float4>();
    a.xconst = a.fconst = nullptr;
    a.frame = c->frame.as<DevFrame>();
    {
        const DevFrame& hf = c->hostFrame;
        RtFrame& r = a.fr;
        for (int i = 0; i < 3; ++i) r.cam[i] = hf.cam[i], r.indirect[i] = hf.indirect[i];
        for (int i = 0; i < 9; ++i) r.R[i] = hf.R[i];
        r.focal = hf.focal;
        r.dofFocal = hf.dofFocal;
        for (int i = 0; i < 3; ++i) {
            volatile float prod = hf.R[6 + i] * hf.focal;  // one IEEE single multiplication, as the reference's m[2][i]*v.z
            r.Rf[i] = prod;
        }
        r.aaN = hf.aaN;
        r.nLights = hf.nLights;
        r.samples = hf.samples;
        r.nOrigins = hf.nOrigins;
    }
    a.T = c->T;
    a.W = c->W;
    a.H = c->H;
    a.y0 = y0;
    a.y1 = y1;
    a.tilesX = (c->W + 31) / 32;
    a.rcpTilesX = 1.0f / (float)a.tilesX;
    a.rcpNN = 1.0f / (float)(c->hostFrame.aaN * c->hostFrame.aaN);
    a.tileRowStride = split ? split->stride : 1;
    a.tileRowOffset = split ? split->offset : 0;
    a.nPeerSurfaces = split ? split->nPeers : 0;
    for (int i = 0; i < B2R_MAX_PEERS - 1; ++i) a.peerSurface[i] = (split && i < split->nPeers) ? split->peers[i] : nullptr;
    {
        const int tileRows = (y1 - y0 + 7) / 8;  // of the band; this launch takes every stride-th one
        const int mine = tileRows > a.tileRowOffset ? (tileRows - a.tileRowOffset + a.tileRowStride - 1) / a.tileRowStride : 0;
        a.numTiles = a.tilesX * mine;
        if (mine == 0 && !(split && split->arrive)) return B2R_OK;
    }
    a.colours = d_col;
    a.closest = d_clo;
    a.focal = d_foc;
    a.surface = d_surf;
    a.stats = c->statsOn ? c->stats.as<unsigned long long>() : nullptr;
    if (!c->rtSched.p) {  // the kernel leaves the two scheduler words zero again when it finishes
        CU(c->rtSched.reserve(64 + 4 * kMaxCopyBands), "scheduler alloc");
        CU(cudaMemsetAsync(c->rtSched.p, 0, 64 + 4 * kMaxCopyBands, c->stream), "scheduler clear");
    }
    a.sched = c->rtSched.as<unsigned>();
    a.arrive = split ? split->arrive : nullptr;
    a.arriveCtr = c->rtSched.as<unsigned>() + 2;
    a.bandDone = bandTileRows > 0 ? c->rtSched.as<unsigned>() + 16 : nullptr;
    a.bandTileRows = bandTileRows > 0 ? bandTileRows : 1;
    // Bottom to top unless sub-bands are copied out in row order behind the tracing: the rows a launch ends with set
    // the length of its tail (warps run out of tiles one warp tile apart), and the top rows of a frame -- ceiling, far
    // walls -- are the cheap ones (a quarter of config 3 alone: 234 -> 223 us, a half: 408 -> 399 us).
    a.tileOrder = a.bandDone ? 0 : 1;
    a.useFilter = c->optRtFilter;
    a.reuseLight = c->optRtVariant == 5 ? 0 : 1;
    cudaError_t e = launch_rt_trace_shade(c, a, c->stream);
    if (e == cudaErrorInvalidConfiguration)
        return fail(c, B2R_E_UNSUPPORTED, "raytracer: more than ~100,000 triangles per scene are not supported by the brute-force tracer");
    if (e != cudaSuccess) return cuda_fail(c, e, "rt_trace_shade_kernel");
    return B2R_OK;
}

int b2r_rt_draw_device_async(b2r_ctx* ctx, int y0, int y1, float* d_col, b2r_intersection* d_clo, float* d_foc) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (int rc = check_band(c, y0, y1)) return rc;
    if (int rc = reset_stats(c)) return rc;
    return rt_launch_band(c, y0, y1, d_col, d_clo, d_foc);
}

// Host-buffer draw.  Large frames are cut into a few sub-bands so that the device-to-host copy of one
// sub-band overlaps the tracing of the next (second stream + events); the arithmetic per pixel is unchanged.
static int rt_draw_host(Ctx* c, int y0, int y1, float* col, b2r_intersection* clo, float* foc, uint32_t* surface) {
    if (int rc = check_band(c, y0, y1)) return rc;
    if (int rc = reset_stats(c)) return rc;
    const size_t n = (size_t)c->W * c->H;
    // without depth of field the trace kernel writes the surface itself (and pixelColours only if asked for)
    const bool fused = surface && !c->params.dofEnabled;
    const bool needCol = col || (surface && !fused);
    if (needCol) CU(c->colours.reserve(n * 12), "alloc pixelColours");
    if (clo) CU(c->closest.reserve(n * 20), "alloc closestIntersections");
    const bool needFocal = foc || (surface && c->params.dofEnabled);
    if (needFocal) CU(c->focal.reserve(n * 4), "alloc focalDistances");
    if (surface) CU(c->surface.reserve(n * 4), "alloc surface");
    const int rows = y1 - y0;
    const bool anyOut = surface || col || clo || foc;
    // the depth-of-field window reads rows of neighbouring sub-bands, so it resolves only after the whole draw
    int parts = (!anyOut || (surface && c->params.dofEnabled)) ? 1 : (rows >= 1024 ? 4 : (rows >= 256 ? 2 : 1));
    if (parts > 1)
        if (int rc = ensure_copy_stream(c)) return rc;
    cudaStream_t drawStream = c->stream;
    if (parts > 1) probe_mem_ops(c);
    if (parts > 1 && c->waitValue32 && c->optRtVariant != 4) {
        // One launch for the whole band.  The kernel counts finished warp tiles per sub-band (tiles are handed out
        // in row order); the copy stream waits -- on the GPU, no host involvement -- until a sub-band's count is
        // complete and copies its rows out while the rest of the frame is still being traced.  The copy engine is
        // about twice as fast as the tracing at config 3, so with 16 sub-bands it stays right behind the frontier
        // and only the last sub-band's copy is exposed.
        typedef int (*WaitFn)(cudaStream_t, unsigned long long, unsigned, unsigned);
        const int tileRows = (rows + 7) / 8, tilesX = (c->W + 31) / 32;
        const int perBand = (tileRows + 15) / 16, nb = (tileRows + perBand - 1) / perBand;
        if (!c->rtSched.p) {
            CU(c->rtSched.reserve(64 + 4 * kMaxCopyBands), "scheduler alloc");
            CU(cudaMemsetAsync(c->rtSched.p, 0, 64 + 4 * kMaxCopyBands, drawStream), "scheduler clear");
        }
        unsigned* done = c->rtSched.as<unsigned>() + 16;
        CU(cudaMemsetAsync(done, 0, 4 * kMaxCopyBands, drawStream), "band counters clear");
        CU(cudaEventRecord(c->partDone[0], drawStream), "cudaEventRecord");
        CU(cudaStreamWaitEvent(c->copyStream, c->partDone[0], 0), "cudaStreamWaitEvent");
        if (int rc = rt_launch_band(c, y0, y1, needCol ? c->colours.as<float>() : nullptr,
                                    clo ? c->closest.as<b2r_intersection>() : nullptr,
                                    needFocal ? c->focal.as<float>() : nullptr, fused ? c->surface.as<uint32_t>() : nullptr, perBand))
            return rc;
        c->stream = c->copyStream;  // copy_rows_out issues on c->stream
        int rc = 0;
        for (int b = 0; b < nb && !rc; ++b) {
            const int t0 = b * perBand, t1 = (b + 1) * perBand < tileRows ? (b + 1) * perBand : tileRows;
            const int a0 = y0 + t0 * 8, a1 = b == nb - 1 ? y1 : y0 + t1 * 8;
            const unsigned want = (unsigned)(t1 - t0) * (unsigned)tilesX * 8u;
            if (reinterpret_cast<WaitFn>(c->waitValue32)(c->copyStream, (unsigned long long)(uintptr_t)(done + b), want,
                                                         0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) {
                c->stream = drawStream;
                return fail(c, B2R_E_CUDA, "cuStreamWaitValue32 failed");
            }
            rc = copy_rows_out(c, surface, c->surface.p, a0, a1, 4);  // (depth of field never gets here: parts == 1)
            if (!rc) rc = copy_rows_out(c, col, c->colours.p, a0, a1, 12);
            if (!rc) rc = copy_rows_out(c, clo, c->closest.p, a0, a1, 20);
            if (!rc) rc = copy_rows_out(c, foc, c->focal.p, a0, a1, 4);
        }
        c->stream = drawStream;
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->copyStream), "raytracer draw (copies)");
        CU(cudaStreamSynchronize(c->stream), "raytracer draw");
        c->coloursValid = needCol;
        c->surfaceValid = surface && y0 == 0 && y1 == c->H;
        return B2R_OK;
    }
    // Only the last sub-band's copy is exposed (the others overlap the next sub-band's tracing), so the sub-bands
    // shrink towards the end; cuts fall on multiples of the 8-row tile height.
    static const int kCut4[5] = {0, 32, 60, 84, 100}, kCut2[3] = {0, 64, 100};
    auto cut = [&](int p) {
        if (p >= parts) return y1;
        const int pct = parts == 4 ? kCut4[p] : (parts == 2 ? kCut2[p] : 0);
        return y0 + (int)((long long)rows * pct / 100) / 8 * 8;
    };
    for (int p = 0; p < parts; ++p) {
        const int a0 = cut(p), a1 = cut(p + 1);
        if (int rc = rt_launch_band(c, a0, a1, needCol ? c->colours.as<float>() : nullptr,
                                    clo ? c->closest.as<b2r_intersection>() : nullptr,
                                    needFocal ? c->focal.as<float>() : nullptr, fused ? c->surface.as<uint32_t>() : nullptr))
            return rc;
        if (surface && !fused)
            CU(launch_resolve_surface(c, a0, a1, c->colours.as<float>(), c->focal.as<float>(), c->surface.as<uint32_t>(), drawStream),
               "resolve_surface_kernel");
        if (parts > 1) {
            CU(cudaEventRecord(c->partDone[p], drawStream), "cudaEventRecord");
            CU(cudaStreamWaitEvent(c->copyStream, c->partDone[p], 0), "cudaStreamWaitEvent");
            c->stream = c->copyStream;  // copy_rows_out issues on c->stream
        }
        int rc = copy_rows_out(c, surface, c->surface.p, a0, a1, 4);
        if (!rc) rc = copy_rows_out(c, col, c->colours.p, a0, a1, 12);
        if (!rc) rc = copy_rows_out(c, clo, c->closest.p, a0, a1, 20);
        if (!rc) rc = copy_rows_out(c, foc, c->focal.p, a0, a1, 4);
        c->stream = drawStream;
        if (rc) return rc;
    }
    if (parts > 1) CU(cudaStreamSynchronize(c->copyStream), "raytracer draw (copies)");
    CU(cudaStreamSynchronize(c->stream), "raytracer draw");
    c->coloursValid = needCol;
    c->surfaceValid = surface && y0 == 0 && y1 == c->H;
    return B2R_OK;
}

int b2r_rt_draw(b2r_ctx* ctx, int y0, int y1, float* col, b2r_intersection* clo, float* foc) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    return rt_draw_host(c, y0, y1, col, clo, foc, nullptr);
}

int b2r_rt_frame_device_async(b2r_ctx* ctx, int y0, int y1, uint32_t* d_surface, float* d_col, b2r_intersection* d_clo,
                              float* d_foc) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (int rc = check_band(c, y0, y1)) return rc;
    if (int rc = reset_stats(c)) return rc;
    if (!d_surface) return rt_launch_band(c, y0, y1, d_col, d_clo, d_foc);
    if (!c->params.dofEnabled) return rt_launch_band(c, y0, y1, d_col, d_clo, d_foc, d_surface);
    // depth of field: the blur window reads neighbouring rows, so the caller's arrays must hold them already
    if (!d_col || !d_foc) return fail(c, B2R_E_INVALID, "b2r_rt_frame_device_async: depth of field needs d_pixelColours and d_focalDistances");
    if (int rc = rt_launch_band(c, y0, y1, d_col, d_clo, d_foc)) return rc;
    CU(launch_resolve_surface(c, y0, y1, d_col, d_foc, d_surface, c->stream), "resolve_surface_kernel");
    return B2R_OK;
}

int b2r_rt_frame_split_device_async(b2r_ctx* ctx, int part, int nparts, uint32_t* const* d_surfaces, int n, float* d_col,
                                    b2r_intersection* d_clo, float* d_foc) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (nparts < 1 || part < 0 || part >= nparts || !d_surfaces || n < 1 || n > B2R_MAX_PEERS)
        return fail(c, B2R_E_INVALID, "rt_frame_split: bad arguments (0 <= part < nparts, 1..8 destination surfaces)");
    for (int i = 0; i < n; ++i)
        if (!d_surfaces[i]) return fail(c, B2R_E_INVALID, "rt_frame_split: null destination");
    if (c->params.dofEnabled)
        return fail(c, B2R_E_UNSUPPORTED, "rt_frame_split: depth of field needs the neighbouring rows; draw, exchange pixelColours, then resolve");
    if (int rc = check_band(c, 0, c->H)) return rc;
    if (int rc = reset_stats(c)) return rc;
    RtSplit sp;
    sp.peers = d_surfaces + 1;
    sp.nPeers = n - 1;
    sp.stride = nparts;
    sp.offset = part;
    return rt_launch_band(c, 0, c->H, d_col, d_clo, d_foc, d_surfaces[0], 0, &sp);
}

int b2r_rt_frame_gather_device_async(b2r_ctx* ctx, int part, int nparts, uint32_t* d_root_surface, uint32_t* d_arrive) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (nparts < 1 || part < 0 || part >= nparts || !d_root_surface)
        return fail(c, B2R_E_INVALID, "rt_frame_gather: bad arguments (0 <= part < nparts, a destination surface)");
    if (c->params.dofEnabled)
        return fail(c, B2R_E_UNSUPPORTED, "rt_frame_gather: depth of field needs the neighbouring rows; draw, exchange pixelColours, then resolve");
    if (int rc = check_band(c, 0, c->H)) return rc;
    if (int rc = reset_stats(c)) return rc;
    RtSplit sp;
    sp.stride = nparts;
    sp.offset = part;
    sp.arrive = reinterpret_cast<unsigned*>(d_arrive);
    return rt_launch_band(c, 0, c->H, nullptr, nullptr, nullptr, d_root_surface, 0, &sp);
}

int b2r_stream_wait_value32(b2r_ctx* ctx, const uint32_t* d_word, uint32_t value) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!d_word) return fail(c, B2R_E_INVALID, "stream_wait_value32: null word");
    probe_mem_ops(c);
    if (!c->waitValue32) return fail(c, B2R_E_UNSUPPORTED, "stream memory operations (cuStreamWaitValue32) are not available");
    typedef int (*WaitFn)(cudaStream_t, unsigned long long, unsigned, unsigned);
    if (reinterpret_cast<WaitFn>(c->waitValue32)(c->stream, (unsigned long long)(uintptr_t)d_word, value, 0u /* GEQ */) != 0)
        return fail(c, B2R_E_CUDA, "cuStreamWaitValue32 failed");
    return B2R_OK;
}

// One part of a frame split over several GPUs, host side: draws tile rows part, part + nparts, ... (8 pixel rows each)
// and copies exactly those rows into the caller's full-frame host surface -- every GPU over its own PCIe link.
int b2r_rt_frame_part(b2r_ctx* ctx, int part, int nparts, uint32_t* surface) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (int rc = b2r_rt_frame_part_async(ctx, part, nparts, surface)) return rc;
    CU(cudaStreamSynchronize(c->stream), "rt_frame_part");
    return B2R_OK;
}

int b2r_rt_frame_part_async(b2r_ctx* ctx, int part, int nparts, uint32_t* surface) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (nparts < 1 || part < 0 || part >= nparts || !surface) return fail(c, B2R_E_INVALID, "rt_frame_part: bad arguments");
    if (c->params.dofEnabled)
        return fail(c, B2R_E_UNSUPPORTED, "rt_frame_part: depth of field needs the neighbouring rows");
    if (int rc = check_band(c, 0, c->H)) return rc;
    if (int rc = reset_stats(c)) return rc;
    const size_t n = (size_t)c->W * c->H;
    CU(c->surface.reserve(n * 4), "alloc surface");
    RtSplit sp;
    sp.stride = nparts;
    sp.offset = part;
    // rows of tile row t: [8t, 8t+8); this part owns t = part, part + nparts, ...: 2-D copies with pitch nparts * 8 rows
    const int tileRows = (c->H + 7) / 8, tilesX = (c->W + 31) / 32;
    const int mine = tileRows > part ? (tileRows - part + nparts - 1) / nparts : 0;
    const size_t rowBytes = (size_t)c->W * 4, chunk = 8 * rowBytes, pitch = (size_t)nparts * chunk;
    // local tile rows [t0,t1) of this part -> host, on stream s
    auto copy_tile_rows = [&](int t0, int t1, cudaStream_t s) -> cudaError_t {
        if (t1 <= t0) return cudaSuccess;
        const int lastTile = part + (t1 - 1) * nparts;
        const int lastRows = std::min(8, c->H - lastTile * 8);  // the frame's last tile row may be short
        const int full = lastRows == 8 ? t1 - t0 : t1 - t0 - 1;
        const size_t first = (size_t)(part + t0 * nparts) * chunk;
        cudaError_t e = cudaSuccess;
        if (full > 0)
            e = cudaMemcpy2DAsync((char*)surface + first, pitch, (const char*)c->surface.p + first, pitch, chunk, (size_t)full,
                                  cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && full < t1 - t0) {
            const size_t off = (size_t)lastTile * chunk;
            e = cudaMemcpyAsync((char*)surface + off, (const char*)c->surface.p + off, (size_t)lastRows * rowBytes,
                                cudaMemcpyDeviceToHost, s);
        }
        return e;
    };
    if (int rc = probe_mem_ops(c)) return rc;
    if (mine >= 16 && c->waitValue32 && c->optRtVariant != 4) {
        // As in b2r_rt_frame: one launch; the kernel counts finished warp tiles per sub-band of its own tile rows and a
        // second stream copies a sub-band out as soon as its count is complete, while the rest is still being traced.
        if (int rc = ensure_copy_stream(c)) return rc;
        typedef int (*WaitFn)(cudaStream_t, unsigned long long, unsigned, unsigned);
        // up to 8 sub-bands of at least 4 tile rows: only the last one's copy is exposed, every one costs a stream wait
        // and a copy call on the host (an eighth of a 4K frame: 2 sub-bands 0.37 ms per frame, 8 sub-bands 0.34 ms)
        const int wantBands = std::max(1, std::min(8, mine / 4));
        const int perBand = (mine + wantBands - 1) / wantBands, nb = (mine + perBand - 1) / perBand;
        if (!c->rtSched.p) {
            CU(c->rtSched.reserve(64 + 4 * kMaxCopyBands), "scheduler alloc");
            CU(cudaMemsetAsync(c->rtSched.p, 0, 64 + 4 * kMaxCopyBands, c->stream), "scheduler clear");
        }
        unsigned* done = c->rtSched.as<unsigned>() + 16;
        CU(cudaMemsetAsync(done, 0, 4 * kMaxCopyBands, c->stream), "band counters clear");
        CU(cudaEventRecord(c->partDone[0], c->stream), "cudaEventRecord");
        CU(cudaStreamWaitEvent(c->copyStream, c->partDone[0], 0), "cudaStreamWaitEvent");
        if (int rc = rt_launch_band(c, 0, c->H, nullptr, nullptr, nullptr, c->surface.as<uint32_t>(), perBand, &sp)) return rc;
        for (int b = 0; b < nb; ++b) {
            const int t0 = b * perBand, t1 = std::min(mine, (b + 1) * perBand);
            const unsigned want = (unsigned)(t1 - t0) * (unsigned)tilesX * 8u;
            if (reinterpret_cast<WaitFn>(c->waitValue32)(c->copyStream, (unsigned long long)(uintptr_t)(done + b), want,
                                                         0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
                return fail(c, B2R_E_CUDA, "cuStreamWaitValue32 failed");
            CU(copy_tile_rows(t0, t1, c->copyStream), "D2H copy (tile rows)");
        }
        // the context's stream completes after the copies: b2r_synchronize covers both
        CU(cudaEventRecord(c->partDone[1], c->copyStream), "cudaEventRecord");
        CU(cudaStreamWaitEvent(c->stream, c->partDone[1], 0), "cudaStreamWaitEvent");
    } else {
        if (int rc = rt_launch_band(c, 0, c->H, nullptr, nullptr, nullptr, c->surface.as<uint32_t>(), 0, &sp)) return rc;
        CU(copy_tile_rows(0, mine, c->stream), "D2H copy (tile rows)");
    }
    c->surfaceValid = false;
    return B2R_OK;
}

int b2r_rt_frame(b2r_ctx* ctx, uint32_t* surface, float* col, b2r_intersection* clo, float* foc) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    return rt_draw_host(c, 0, c->H, col, clo, foc, surface);
}

// ---- rasteriser -------------------------------------------------------------------
static int ras_launch_band(Ctx* c, int y0, int y1, float* d_dep, float* d_col, float* d_foc, int32_t* d_win,
                           uint32_t* d_surf) {
    c->lastDraw = 1;
    c->coloursValid = c->surfaceValid = false;  // set again by the host-buffer draws once their frame is complete
    if (y1 == y0) return B2R_OK;
    RasLaunch a;
    a.raw = c->stride == 64 ? c->raw.as<unsigned char>() : c->raw64.as<unsigned char>();
    a.stride = 64;
    a.culled = c->culled.as<unsigned char>();
    a.T = c->T;
    a.frame = c->frame.as<DevFrame>();
    {
        const DevFrame& f = c->hostFrame;
        RasFrame& r = a.fr;
        for (int i = 0; i < 3; ++i) {
            r.cam[i] = f.cam[i];
            r.reflectance[i] = f.reflectance[i];
            r.indirect[i] = f.indirect[i];
        }
        for (int i = 0; i < 9; ++i) {
            r.R[i] = f.R[i];
            r.Rinv[i] = f.Rinv[i];
        }
        r.focal = f.focal;
        r.dofFocal = f.dofFocal;
        r.halfW = xdiv((float)c->W, 2.0f);
        r.halfH = xdiv((float)c->H, 2.0f);
        r.nLights = f.nLights;
        for (int k = 0; k < B2R_MAX_LIGHTS; ++k)
            for (int i = 0; i < 3; ++i) {
                r.lightPos[k][i] = f.lightPos[k][i];
                r.lightColor[k][i] = f.lightColor[k][i];
            }
    }
    a.W = c->W;
    a.H = c->H;
    a.y0 = y0;
    a.y1 = y1;
    a.depth = d_dep;
    a.colours = d_col;
    a.focal = d_foc;
    a.winner = d_win;
    a.surface = d_surf;
    a.stats = c->statsOn ? c->stats.as<unsigned long long>() : nullptr;
    cudaError_t e = launch_ras_draw(c, a, c->stream);
    if (e == cudaErrorInvalidValue) return fail(c, B2R_E_CAPACITY, kRasCapacityText);
    if (e != cudaSuccess) return cuda_fail(c, e, "rasteriser kernels");
    return B2R_OK;
}

int b2r_ras_draw_device_async(b2r_ctx* ctx, int y0, int y1, float* d_dep, float* d_col, float* d_foc, int32_t* d_win) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (int rc = check_band(c, y0, y1)) return rc;
    if (int rc = reset_stats(c)) return rc;
    return ras_launch_band(c, y0, y1, d_dep, d_col, d_foc, d_win, nullptr);
}

int b2r_ras_frame_device_async(b2r_ctx* ctx, int y0, int y1, uint32_t* d_surface, float* d_dep, float* d_col, float* d_foc,
                               int32_t* d_win) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (int rc = check_band(c, y0, y1)) return rc;
    if (int rc = reset_stats(c)) return rc;
    if (!d_surface || !c->params.dofEnabled) return ras_launch_band(c, y0, y1, d_dep, d_col, d_foc, d_win, d_surface);
    if (!d_col || !d_foc) return fail(c, B2R_E_INVALID, "b2r_ras_frame_device_async: depth of field needs d_pixelColours and d_focalDistances");
    if (int rc = ras_launch_band(c, y0, y1, d_dep, d_col, d_foc, d_win, nullptr)) return rc;
    CU(launch_resolve_surface(c, y0, y1, d_col, d_foc, d_surface, c->stream), "resolve_surface_kernel");
    return B2R_OK;
}

static int ras_draw_host(Ctx* c, int y0, int y1, float* dep, float* col, float* foc, int32_t* win, uint32_t* surface) {
    if (int rc = check_band(c, y0, y1)) return rc;
    const size_t n = (size_t)c->W * c->H;
    // without depth of field the shade kernel writes the surface itself (and pixelColours only if asked for)
    const bool fused = surface && !c->params.dofEnabled;
    const bool needCol = col || (surface && !fused);
    if (needCol) CU(c->colours.reserve(n * 12), "alloc pixelColours");
    if (dep) CU(c->depth.reserve(n * 4), "alloc depthBuffer");
    if (win) CU(c->winner.reserve(n * 4), "alloc winnerIndex");
    if (surface) CU(c->surface.reserve(n * 4), "alloc surface");
    const bool needFocal = foc || (surface && c->params.dofEnabled);
    if (needFocal) CU(c->focal.reserve(n * 4), "alloc focalDistances");
    if (int rc = reset_stats(c)) return rc;
    if (int rc = ras_launch_band(c, y0, y1, dep ? c->depth.as<float>() : nullptr, needCol ? c->colours.as<float>() : nullptr,
                                 needFocal ? c->focal.as<float>() : nullptr, win ? c->winner.as<int32_t>() : nullptr,
                                 fused ? c->surface.as<uint32_t>() : nullptr))
        return rc;
    if (surface) {
        if (!fused)
            CU(launch_resolve_surface(c, y0, y1, c->colours.as<float>(), c->focal.as<float>(), c->surface.as<uint32_t>(), c->stream),
               "resolve_surface_kernel");
        if (int rc = copy_rows_out(c, surface, c->surface.p, y0, y1, 4)) return rc;
    }
    if (int rc = copy_rows_out(c, dep, c->depth.p, y0, y1, 4)) return rc;
    if (int rc = copy_rows_out(c, col, c->colours.p, y0, y1, 12)) return rc;
    if (int rc = copy_rows_out(c, foc, c->focal.p, y0, y1, 4)) return rc;
    if (int rc = copy_rows_out(c, win, c->winner.p, y0, y1, 4)) return rc;
    CU(cudaStreamSynchronize(c->stream), "rasteriser draw");
    if (int rc = ras_deferred_error(c)) return rc;
    c->coloursValid = needCol;
    c->surfaceValid = surface && y0 == 0 && y1 == c->H;
    return B2R_OK;
}

int b2r_ras_draw(b2r_ctx* ctx, int y0, int y1, float* dep, float* col, float* foc, int32_t* win) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    return ras_draw_host(c, y0, y1, dep, col, foc, win, nullptr);
}

int b2r_ras_frame(b2r_ctx* ctx, uint32_t* surface, float* dep, float* col, float* foc, int32_t* win) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    return ras_draw_host(c, 0, c->H, dep, col, foc, win, surface);
}

int b2r_ras_cull(b2r_ctx* ctx, uint8_t* culledOut) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!c->haveScene || !c->haveFrame) return fail(c, B2R_E_NO_SCENE, "b2r_ras_cull needs a scene and frame params");
    CU(launch_ras_cull(c, c->culled.as<unsigned char>(), c->stream), "ras_cull_kernel");
    c->rasGen++;
    if (culledOut && c->T)
        CU(cudaMemcpyAsync(culledOut, c->culled.p, (size_t)c->T, cudaMemcpyDeviceToHost, c->stream), "culled D2H");
    CU(cudaStreamSynchronize(c->stream), "ras_cull");
    return B2R_OK;
}

// ---- resolve ----------------------------------------------------------------------
int b2r_resolve_surface_device_async(b2r_ctx* ctx, int y0, int y1, const float* d_col, const float* d_foc, uint32_t* d_surface) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (y0 < 0 || y1 > c->H || y0 > y1 || !d_col || !d_surface) return fail(c, B2R_E_INVALID, "resolve: bad arguments");
    if (c->params.dofEnabled && !d_foc) return fail(c, B2R_E_INVALID, "resolve: depth of field needs focalDistances");
    CU(launch_resolve_surface(c, y0, y1, d_col, d_foc, d_surface, c->stream), "resolve_surface_kernel");
    return B2R_OK;
}

int b2r_resolve_surface_multi_device_async(b2r_ctx* ctx, int y0, int y1, const float* d_col, const float* d_foc,
                                           uint32_t* const* d_surfaces, int n) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (y0 < 0 || y1 > c->H || y0 > y1 || !d_col || !d_surfaces || n < 1 || n > B2R_MAX_PEERS)
        return fail(c, B2R_E_INVALID, "resolve_multi: bad arguments (1..8 destination surfaces)");
    for (int i = 0; i < n; ++i)
        if (!d_surfaces[i]) return fail(c, B2R_E_INVALID, "resolve_multi: null destination");
    if (c->params.dofEnabled && !d_foc) return fail(c, B2R_E_INVALID, "resolve: depth of field needs focalDistances");
    CU(launch_resolve_surface_multi(c, y0, y1, d_col, d_foc, d_surfaces, n, c->stream), "resolve_surface_kernel");
    return B2R_OK;
}

// ---- inter-process shared device buffers (CUDA IPC) for the fused band exchange -----------------
int b2r_shared_alloc(b2r_ctx* ctx, size_t bytes, void** d_ptr, void* handle_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!d_ptr || !handle_out || bytes == 0) return fail(c, B2R_E_INVALID, "shared_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == B2R_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes), "cudaMalloc (shared)");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(c, e, "cudaIpcGetMemHandle");
    }
    memcpy(handle_out, &h, sizeof h);
    *d_ptr = p;
    return B2R_OK;
}

int b2r_shared_free(b2r_ctx* ctx, void* d_ptr) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (d_ptr) CU(cudaFree(d_ptr), "cudaFree (shared)");
    return B2R_OK;
}

int b2r_shared_open(b2r_ctx* ctx, const void* handle, void** d_ptr) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!handle || !d_ptr) return fail(c, B2R_E_INVALID, "shared_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    *d_ptr = p;
    return B2R_OK;
}

// Page-lock a caller-owned host buffer (e.g. screen->pixels, pixelColours) so the copies of the frame calls run at
// full PCIe speed and overlap the kernels; pageable memory goes through the driver's staging buffer instead.
int b2r_pin_host_buffer(b2r_ctx* ctx, void* host, size_t bytes) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!host || bytes == 0) return fail(c, B2R_E_INVALID, "pin_host_buffer: null buffer");
    cudaError_t e = cudaHostRegister(host, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return B2R_OK;
    }
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaHostRegister");
    return B2R_OK;
}

int b2r_unpin_host_buffer(b2r_ctx* ctx, void* host) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!host) return B2R_OK;
    cudaError_t e = cudaHostUnregister(host);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return B2R_OK;
    }
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaHostUnregister");
    return B2R_OK;
}

int b2r_copy_device_async(b2r_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!d_dst || !d_src) return fail(c, B2R_E_INVALID, "copy_device: null pointer");
    if (bytes) CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, c->stream), "cudaMemcpyAsync D2D");
    return B2R_OK;
}

int b2r_shared_close(b2r_ctx* ctx, void* d_ptr) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr), "cudaIpcCloseMemHandle");
    return B2R_OK;
}

int b2r_resolve_surface(b2r_ctx* ctx, uint32_t* surface) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!surface) return fail(c, B2R_E_INVALID, "null surface");
    const bool haveSurface = c->surfaceValid && !c->coloursValid;  // fused raytracer frame: already resolved
    if (c->lastDraw < 0 || (!haveSurface && !c->coloursValid))
        return fail(c, B2R_E_NO_SCENE, "resolve: the last draw on this context left no frame here (it was a device-pointer "
                                       "draw, a partial band without pixelColours, or there was none)");
    if (c->params.dofEnabled && !c->focal.p) return fail(c, B2R_E_INVALID, "resolve: the last draw did not produce focalDistances");
    const size_t n = (size_t)c->W * c->H;
    CU(c->surface.reserve(n * 4), "alloc surface");
    if (!haveSurface)
        CU(launch_resolve_surface(c, 0, c->H, c->colours.as<float>(), c->focal.as<float>(), c->surface.as<uint32_t>(), c->stream),
           "resolve_surface_kernel");
    CU(cudaMemcpyAsync(surface, c->surface.p, n * 4, cudaMemcpyDeviceToHost, c->stream), "surface D2H");
    CU(cudaStreamSynchronize(c->stream), "resolve");
    return B2R_OK;
}

size_t b2r_bmp_payload_bytes(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    return (size_t)((width * 3 + 3) & ~3) * (size_t)height;
}

int b2r_resolve_bgr8(b2r_ctx* ctx, uint8_t* bgr) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!bgr) return fail(c, B2R_E_INVALID, "null bgr");
    const bool haveSurface = c->surfaceValid && !c->coloursValid;  // fused raytracer frame: already resolved
    if (c->lastDraw < 0 || (!haveSurface && !c->coloursValid))
        return fail(c, B2R_E_NO_SCENE, "resolve: the last draw on this context left no frame here (it was a device-pointer "
                                       "draw, a partial band without pixelColours, or there was none)");
    if (c->params.dofEnabled && !c->focal.p) return fail(c, B2R_E_INVALID, "resolve: the last draw did not produce focalDistances");
    const size_t n = (size_t)c->W * c->H, payload = b2r_bmp_payload_bytes(c->W, c->H);
    CU(c->surface.reserve(n * 4), "alloc surface");
    CU(c->bgr.reserve(payload), "alloc bgr");
    if (!haveSurface)
        CU(launch_resolve_surface(c, 0, c->H, c->colours.as<float>(), c->focal.as<float>(), c->surface.as<uint32_t>(), c->stream),
           "resolve_surface_kernel");
    CU(launch_surface_to_bgr8(c, c->surface.as<uint32_t>(), c->bgr.as<uint8_t>(), c->stream), "surface_to_bgr8_kernel");
    CU(cudaMemcpyAsync(bgr, c->bgr.p, payload, cudaMemcpyDeviceToHost, c->stream), "bgr D2H");
    CU(cudaStreamSynchronize(c->stream), "resolve");
    return B2R_OK;
}

// Draw() of the raytracer for the whole frame straight to the BMP payload: trace (+ resolve), BGR conversion on the
// device, one D2H copy of 3 bytes per pixel; returns after enqueueing (b2r_synchronize waits).
int b2r_rt_frame_bgr8_async(b2r_ctx* ctx, uint8_t* bgr) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!bgr) return fail(c, B2R_E_INVALID, "null bgr");
    if (int rc = check_band(c, 0, c->H)) return rc;
    if (int rc = reset_stats(c)) return rc;
    const size_t n = (size_t)c->W * c->H, payload = b2r_bmp_payload_bytes(c->W, c->H);
    CU(c->surface.reserve(n * 4), "alloc surface");
    CU(c->bgr.reserve(payload), "alloc bgr");
    if (!c->params.dofEnabled) {
        if (int rc = rt_launch_band(c, 0, c->H, nullptr, nullptr, nullptr, c->surface.as<uint32_t>())) return rc;
    } else {
        CU(c->colours.reserve(n * 12), "alloc pixelColours");
        CU(c->focal.reserve(n * 4), "alloc focalDistances");
        if (int rc = rt_launch_band(c, 0, c->H, c->colours.as<float>(), nullptr, c->focal.as<float>())) return rc;
        CU(launch_resolve_surface(c, 0, c->H, c->colours.as<float>(), c->focal.as<float>(), c->surface.as<uint32_t>(), c->stream),
           "resolve_surface_kernel");
    }
    CU(launch_surface_to_bgr8(c, c->surface.as<uint32_t>(), c->bgr.as<uint8_t>(), c->stream), "surface_to_bgr8_kernel");
    CU(cudaMemcpyAsync(bgr, c->bgr.p, payload, cudaMemcpyDeviceToHost, c->stream), "bgr D2H");
    c->coloursValid = c->surfaceValid = false;
    return B2R_OK;
}

// Rows [y0,y1) of the rasteriser's Draw() into the caller's full-frame host surface; returns after enqueueing.
int b2r_ras_frame_part_async(b2r_ctx* ctx, int y0, int y1, uint32_t* surface) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (int rc = bind(c)) return rc;
    if (!surface) return fail(c, B2R_E_INVALID, "null surface");
    if (int rc = check_band(c, y0, y1)) return rc;
    if (c->params.dofEnabled)
        return fail(c, B2R_E_UNSUPPORTED, "ras_frame_part: depth of field needs the neighbouring rows");
    if (int rc = reset_stats(c)) return rc;
    CU(c->surface.reserve((size_t)c->W * c->H * 4), "alloc surface");
    if (int rc = ras_launch_band(c, y0, y1, nullptr, nullptr, nullptr, nullptr, c->surface.as<uint32_t>())) return rc;
    c->coloursValid = c->surfaceValid = false;
    return copy_rows_out(c, surface, c->surface.p, y0, y1, 4);
}

// Headless SDL_SaveBMP: BITMAPFILEHEADER + BITMAPINFOHEADER (54 bytes), 24 bpp, bottom-up.
int b2r_write_bmp(const char* path, const uint8_t* payload, int width, int height) {
    if (!path || !payload || width <= 0 || height <= 0) return B2R_E_INVALID;
    const uint32_t size = (uint32_t)b2r_bmp_payload_bytes(width, height);
    unsigned char h[54];
    memset(h, 0, sizeof h);
    auto put32 = [&](int off, uint32_t v) { h[off] = v & 0xFF; h[off + 1] = (v >> 8) & 0xFF; h[off + 2] = (v >> 16) & 0xFF; h[off + 3] = (v >> 24) & 0xFF; };
    h[0] = 'B';
    h[1] = 'M';
    put32(2, 54u + size);
    put32(10, 54u);
    put32(14, 40u);
    put32(18, (uint32_t)width);
    put32(22, (uint32_t)height);
    h[26] = 1;   // planes
    h[28] = 24;  // bits per pixel
    put32(34, size);
    FILE* fp = fopen(path, "wb");
    if (!fp) return B2R_E_IO;
    const bool ok = fwrite(h, 1, 54, fp) == 54 && fwrite(payload, 1, size, fp) == size;
    return (fclose(fp) == 0 && ok) ? B2R_OK : B2R_E_IO;
}

}  // extern "C"
