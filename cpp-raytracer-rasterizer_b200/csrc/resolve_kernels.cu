// resolve_kernels.cu -- the frame tail both reference programs share:
//   CalculateDOF()   raytracer.cpp:608-656 == rasteriser.cpp:484-529
//   PutPixelSDL()    SDLauxiliary.h:70-81  (Uint8(clamp(255*c, 0, 255)), truncation)
// and the 24-bit bottom-up payload SDL_SaveBMP writes (raytracer.cpp:175).
// One thread per pixel; HBM-bound (12-16 B read, 4 B written per pixel).
// Also: the FP32 FFMA throughput microbenchmark that gives the raytracer's
// roofline denominator (MEASURED_PEAKS.json has no FP32 entry).
#include <type_traits>

#include "b2r_internal.h"
#include "exact.cuh"
#include "pixel_pack.cuh"

namespace b2r {

// Out-of-range policy for the DOF window (the reference reads without bounds checks,
// raytracer.cpp:637): the flattened index is used as-is inside [0, W*H) (columns wrap into
// the neighbouring row exactly like the reference) and contributes 0 outside the array.
// Destinations of one resolve: the local surface and/or peer-mapped surfaces of other GPUs (stores go over
// NVLink), which fuses the framebuffer-band exchange of a single-frame multi-GPU split into this kernel.
struct SurfDst {
    uint32_t* p[B2R_MAX_PEERS];
    int n;
};

__global__ void __launch_bounds__(256) resolve_surface_kernel(const float* __restrict__ colours,
                                                              const float* __restrict__ focal, int W, int H,
                                                              int y0, int dof, int K, const SurfDst dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = y0 + blockIdx.y;
    if (x >= W) return;
    uint32_t out = 0u;  // the 1-pixel border is never written by the reference: stays black
    if (inside_border(x, y, W, H)) {
        const long long c = (long long)y * W + x;
        float fr = 0.f, fg = 0.f, fb = 0.f;
        if (dof) {
            const float totalPixels = (float)(K * K);                       // :614
            const int zlo = (int)ceilf(K / -2.0f), zhi = (int)ceilf(K / 2.0f);  // :626
            const float a = fabsf(focal[c]);
            const float m = (1.0f < a) ? 1.0f : a;                          // std::min(abs(fd), 1.0f)
            const float wCentre = xsub(1.0f, xmul(m, xdiv(xsub(totalPixels, 1.0f), totalPixels)));  // :632
            const float wOther = xmul(m, xdiv(1.0f, totalPixels));                                   // :634
            const long long n = (long long)W * H;
            for (int z = zlo; z < zhi; ++z)
                for (int z2 = zlo; z2 < zhi; ++z2) {
                    const float wgt = (z == 0 && z2 == 0) ? wCentre : wOther;
                    const long long q = (long long)(y + z) * W + (x + z2);
                    if (q < 0 || q >= n) continue;
                    fr = xadd(fr, xmul(colours[3 * q], wgt));               // :637
                    fg = xadd(fg, xmul(colours[3 * q + 1], wgt));
                    fb = xadd(fb, xmul(colours[3 * q + 2], wgt));
                }
        } else {
            fr = colours[3 * c];                                            // :643
            fg = colours[3 * c + 1];
            fb = colours[3 * c + 2];
        }
        out = pack_xrgb(fr, fg, fb);
    }
    const size_t o = (size_t)y * W + x;
#pragma unroll
    for (int d = 0; d < B2R_MAX_PEERS; ++d)
        if (d < dst.n) dst.p[d][o] = out;
}

// ---- depth of field with the reference's 8x8 window (DOF_KERNEL_SIZE 8, raytracer.cpp:45) -------------------
// The generic kernel above issues 192 scalar loads per pixel and is bound by the load/store unit.  Here a CTA of
// 256 threads resolves a 128 x 8 pixel tile: the (128+8) x (8+7) window of pixelColours is staged once in shared
// memory exactly as it lies in HBM (rgb interleaved; each window row is one contiguous run of the flattened array,
// which also reproduces the reference's wrap of out-of-row columns into the neighbouring rows), with 128-bit
// loads and stores.  Out-of-array taps are staged as 0: adding 0*w leaves the running sum unchanged (it starts at
// +0 and never becomes -0), which equals the reference's skipped taps.  Every thread resolves 4 adjacent pixels
// and reads each window row with nine 128-bit shared loads.  Per pixel and channel the 64 products are added in the
// reference's order (z rows outer, z2 columns inner, :626-639), unfused.  Needs W % 4 == 0 and a 16-byte aligned
// pixelColours (the launcher falls back to the generic kernel otherwise).
constexpr int kDofTileW = 128, kDofTileH = 8, kDofCols = kDofTileW + 8, kDofRows = kDofTileH + 7;
constexpr int kDofRowFloats = 3 * kDofCols, kDofRowQuads = kDofRowFloats / 4;
static_assert(kDofRowFloats % 4 == 0, "window rows are copied in 16-byte pieces");

__global__ void __launch_bounds__(256) resolve_dof8_kernel(const float* __restrict__ colours,
                                                           const float* __restrict__ focal, int W, int H, int y0,
                                                           int y1, const SurfDst dst) {
    __shared__ __align__(16) float tile[kDofRows * kDofRowFloats];
    const int tx0 = blockIdx.x * kDofTileW, ty0 = y0 + blockIdx.y * kDofTileH;
    const long long nFloats = 3LL * W * H;
    // stage: window row r <-> frame row ty0 - 4 + r, starting at frame column tx0 - 4 (all loads of a thread are
    // issued before its first store, so their latencies overlap)
    constexpr int kQuads = kDofRows * kDofRowQuads, kPerThread = (kQuads + 255) / 256;
    float4 stage[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int e = threadIdx.x + 256 * i;
        const int r = e / kDofRowQuads, k = e - r * kDofRowQuads;
        const long long F = 3 * ((long long)(ty0 - 4 + r) * W + (tx0 - 4)) + 4 * k;  // multiple of 4: W % 4 == 0
        stage[i] = (e < kQuads && F >= 0 && F < nFloats) ? *reinterpret_cast<const float4*>(colours + F)
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int e = threadIdx.x + 256 * i;
        if (e < kQuads) reinterpret_cast<float4*>(tile)[e] = stage[i];
    }
    __syncthreads();
    const int lx = (threadIdx.x & 31) * 4, ly = threadIdx.x >> 5;  // 32 x 8 threads, 4 pixels each
    const int y = ty0 + ly, xb = tx0 + lx;
    if (y >= y1) return;
    float wC[4], wO[4];
    bool live[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = xb + p;
        live[p] = x < W && inside_border(x, y, W, H);
        float a = live[p] ? fabsf(focal[(long long)y * W + x]) : 0.f;
        const float m = (1.0f < a) ? 1.0f : a;                 // std::min(abs(fd), 1.0f)
        wC[p] = xsub(1.0f, xmul(m, xdiv(xsub(64.0f, 1.0f), 64.0f)));  // :632
        wO[p] = xmul(m, xdiv(1.0f, 64.0f));                           // :634
    }
    float acc[4][3];
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0.f;
    // one window row: frame row y - 4 + z = window row ly + z; CENTRE marks the row that holds the centre tap
    auto add_row = [&](int z, auto centre) {
        constexpr bool CENTRE = decltype(centre)::value;
        const float4* row = reinterpret_cast<const float4*>(tile + (ly + z) * kDofRowFloats + 3 * lx);
        float v[36];  // 12 window columns x rgb
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float4 t = row[j];
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int z2 = 0; z2 < 8; ++z2) {  // frame column x - 4 + z2 = window column lx + p + z2
                    const float w = (CENTRE && z2 == 4) ? wC[p] : wO[p];
                    acc[p][ch] = xadd(acc[p][ch], xmul(v[3 * (p + z2) + ch], w));  // :637
                }
    };
#pragma unroll 1
    for (int z = 0; z < 4; ++z) add_row(z, std::false_type{});
    add_row(4, std::true_type{});
#pragma unroll 1
    for (int z = 5; z < 8; ++z) add_row(z, std::false_type{});
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = xb + p;
        if (x >= W) continue;
        const uint32_t out = live[p] ? pack_xrgb(acc[p][0], acc[p][1], acc[p][2]) : 0u;
        const size_t o = (size_t)y * W + x;
#pragma unroll
        for (int d = 0; d < B2R_MAX_PEERS; ++d)
            if (d < dst.n) dst.p[d][o] = out;
    }
}

cudaError_t launch_resolve_surface_multi(Ctx* c, int y0, int y1, const float* d_colours, const float* d_focal,
                                         uint32_t* const* d_surfaces, int n, cudaStream_t s) {
    if (y1 <= y0 || n <= 0) return cudaSuccess;
    SurfDst dst;
    dst.n = n;
    for (int d = 0; d < B2R_MAX_PEERS; ++d) dst.p[d] = d < n ? d_surfaces[d] : nullptr;
    if (c->params.dofEnabled && c->params.dofKernelSize == 8 && c->optDofVariant != 1 && c->W % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(d_colours) & 15u) == 0) {
        dim3 g8((c->W + kDofTileW - 1) / kDofTileW, (y1 - y0 + kDofTileH - 1) / kDofTileH);
        resolve_dof8_kernel<<<g8, 256, 0, s>>>(d_colours, d_focal, c->W, c->H, y0, y1, dst);
        c->launches++;
        return cudaGetLastError();
    }
    dim3 grid((c->W + 255) / 256, y1 - y0);
    resolve_surface_kernel<<<grid, 256, 0, s>>>(d_colours, d_focal, c->W, c->H, y0,
                                                c->params.dofEnabled ? 1 : 0, c->params.dofKernelSize, dst);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_resolve_surface(Ctx* c, int y0, int y1, const float* d_colours, const float* d_focal,
                                   uint32_t* d_surface, cudaStream_t s) {
    return launch_resolve_surface_multi(c, y0, y1, d_colours, d_focal, &d_surface, 1, s);
}

// XRGB surface -> bottom-up BGR rows padded to 4 bytes (the payload of SDL_SaveBMP's 24-bit file).
// One thread per 4 pixels: four 32-bit loads (one 128-bit load when the row allows it), three 32-bit stores -- every
// access coalesced, 7 bytes of traffic per pixel.  Pixels at x >= W read as 0, which is also the value of the padding
// bytes; words beyond the padded row are not written.
__global__ void __launch_bounds__(256) surface_to_bgr8_kernel(const uint32_t* __restrict__ surface, int W, int H,
                                                              int pitch, uint8_t* __restrict__ bgr) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels = 12 payload bytes = 3 words
    const int y = blockIdx.y;
    const int rowWords = pitch >> 2;
    if (3 * g >= rowWords) return;
    const uint32_t* src = surface + (size_t)y * W + 4 * (size_t)g;
    uint32_t p0, p1, p2, p3;
    if (4 * g + 3 < W && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(surface) & 15) == 0) {  // 16-byte aligned rows
        const uint4 q = *reinterpret_cast<const uint4*>(src);
        p0 = q.x; p1 = q.y; p2 = q.z; p3 = q.w;
    } else {
        p0 = 4 * g < W ? src[0] : 0u;
        p1 = 4 * g + 1 < W ? src[1] : 0u;
        p2 = 4 * g + 2 < W ? src[2] : 0u;
        p3 = 4 * g + 3 < W ? src[3] : 0u;
    }
    p0 &= 0xFFFFFFu; p1 &= 0xFFFFFFu; p2 &= 0xFFFFFFu; p3 &= 0xFFFFFFu;  // byte 0 = B, 1 = G, 2 = R
    uint32_t* dst = reinterpret_cast<uint32_t*>(bgr + (size_t)(H - 1 - y) * pitch) + 3 * g;
    dst[0] = p0 | (p1 << 24);                       // B0 G0 R0 B1
    if (3 * g + 1 < rowWords) dst[1] = (p1 >> 8) | (p2 << 16);   // G1 R1 B2 G2
    if (3 * g + 2 < rowWords) dst[2] = (p2 >> 16) | (p3 << 8);   // R2 B3 G3 R3
}

cudaError_t launch_surface_to_bgr8(Ctx* c, const uint32_t* d_surface, uint8_t* d_bgr, cudaStream_t s) {
    const int pitch = (c->W * 3 + 3) & ~3;
    const int groups = (pitch / 4 + 2) / 3;
    dim3 grid((groups + 255) / 256, c->H);
    surface_to_bgr8_kernel<<<grid, 256, 0, s>>>(d_surface, c->W, c->H, pitch, d_bgr);
    c->launches++;
    return cudaGetLastError();
}

// ---- FP32 peak: 8 independent FFMA chains per thread --------------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains alive
}

cudaError_t run_fp32_peak(Ctx* c, double* tflops, double* seconds) {
    cudaError_t e;
    if ((e = c->subScratch.reserve(1 << 20)) != cudaSuccess) return e;
    const int blocks = c->smCount * 8, threads = 256, iters = 4096;
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    double best = 0.0, bestSec = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(ev0, c->stream);
        fp32_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->subScratch.as<float>(), iters, 0.999f, 0.001f);
        cudaEventRecord(ev1, c->stream);
        if ((e = cudaEventSynchronize(ev1)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        double tf = flops / ((double)ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) {
            best = tf;
            bestSec = (double)ms * 1e-3;
        }
        c->launches++;
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (e != cudaSuccess) return e;
    *tflops = best;
    *seconds = bestSec;
    return cudaGetLastError();
}

}  // namespace b2r
