// substage_kernels.cu -- the reference's callee functions as batched entry points.
//
// Draw() is the drop-in boundary, but the reference also declares its stages as free functions
// (raytracer.cpp:105-107, rasteriser.cpp:87-94).  These entry points run exactly those stages on
// caller-provided inputs -- one CUDA thread per item, reference-order arithmetic, no culling -- so each stage can be
// checked on its own against the reference's own function (tests/test_substages_gpu.py uses vectors produced by
// the reference, tests/golden/sub_*.npz) and so that host code written against the stage functions keeps working
// (host/raytracer_dropin.h, host/rasteriser_dropin.h).
#include <float.h>
#include <string.h>

#include "b2r_internal.h"
#include "exact.cuh"
#include "ras_device.cuh"
#include "rt_device.cuh"

namespace b2r {

// ClosestIntersection (raytracer.cpp:202-257) for one ray: brute force over every triangle in index order.
__device__ __forceinline__ bool closest_intersection_bruteforce(const float4* __restrict__ geom, int T, V3 start, V3 dir,
                                                                b2r_intersection& c, bool isLight, float dofFocal,
                                                                float& focalSlot) {
    bool any = false;
    const V3 nd = neg3(dir);  // :229
    for (int i = 0; i < T; ++i) {
        const TriG t = load_geom(geom + (size_t)i * kGeomQuads);
        const OriginTri o = origin_constants(t.v0, t.e1, t.e2, t.n, start);
        V3 pos;
        float dist;
        if (exact_hit_core(t, o.be2, o.e1b, o.nb, start, nd, pos, dist)) {
            if (c.distance >= dist) {  // :243
                c.position[0] = pos.x;
                c.position[1] = pos.y;
                c.position[2] = pos.z;
                c.distance = dist;
                c.triangleIndex = i;
                if (!isLight) focalSlot = xsub(dist, dofFocal);  // :248-249
            }
            any = true;
        }
    }
    return any;
}

__global__ void closest_intersection_kernel(const float4* __restrict__ geom, int T, const DevFrame* __restrict__ f, int n,
                                            const float* __restrict__ starts, const float* __restrict__ dirs,
                                            const int32_t* __restrict__ isLight, b2r_intersection* __restrict__ io,
                                            int32_t* __restrict__ hit, float* __restrict__ focal) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    b2r_intersection c = io[k];
    float slot = 0.f;
    const bool light = isLight && isLight[k];
    const bool any = closest_intersection_bruteforce(geom, T, mk3(starts[3 * k], starts[3 * k + 1], starts[3 * k + 2]),
                                                     mk3(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]), c, light,
                                                     f->dofFocal, slot);
    io[k] = c;
    if (hit) hit[k] = any ? 1 : 0;
    if (focal) focal[k] = slot;
}

// DirectLight (raytracer.cpp:265-327) for one Intersection.
__global__ void direct_light_kernel(const float4* __restrict__ geom, int T, const DevFrame* __restrict__ f, int n,
                                    const b2r_intersection* __restrict__ hits, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const b2r_intersection h = hits[k];
    V3 res = mk3(0.f, 0.f, 0.f);
    if (h.triangleIndex >= 0 && h.triangleIndex < T) {
        const float4 g3 = geom[(size_t)h.triangleIndex * kGeomQuads + 3], g4 = geom[(size_t)h.triangleIndex * kGeomQuads + 4];
        const V3 nDir = mk3(g3.x, g3.y, g3.z), colr = mk3(g3.w, g4.x, g4.y);
        const V3 hp = mk3(h.position[0], h.position[1], h.position[2]);
        V3 result = mk3(0.f, 0.f, 0.f), result2 = mk3(0.f, 0.f, 0.f);
        for (int l = 0; l < f->nLights; ++l) {
            const V3 P = mk3(f->lightPower[l][0], f->lightPower[l][1], f->lightPower[l][2]);  // :282,296
            for (int s = 0; s < f->samples; ++s) {
                const float* og = f->origin[1 + l * f->samples + s];
                const V3 lpos = mk3(og[0], og[1], og[2]);           // :284-291
                const V3 dv = xsub3(lpos, hp);
                const float r = xsqrt(xdot3(dv, dv));               // :294
                const float A = sphere_area(r);                     // :295
                const V3 rDir = xscale3(dv, xdiv(1.0f, r));         // :298
                const V3 B = xdivs3(P, A);                          // :301
                V3 D = xscale3(B, std_max(xdot3(rDir, nDir), 0.0f));  // :304
                b2r_intersection j;
                j.position[0] = j.position[1] = j.position[2] = 0.f;
                j.distance = FLT_MAX;                               // :308
                j.triangleIndex = -1;
                float unused = 0.f;
                if (closest_intersection_bruteforce(geom, T, lpos, neg3(rDir), j, true, 0.f, unused))  // :310
                    if (j.distance < xmul(r, 0.99f)) D = mk3(0.f, 0.f, 0.f);                            // :313-314
                result = xadd3(result, D);                          // :319
            }
            result2 = xadd3(result2, result);                       // :322
        }
        res = xmul3(result2, colr);                                 // :325-326
    }
    out[3 * k] = res.x;
    out[3 * k + 1] = res.y;
    out[3 * k + 2] = res.z;
}

struct PixelRec {  // == struct Pixel (rasteriser TestModel.h:34-53), 24 bytes
    int x, y;
    float zinv, px, py, pz;
};

// VertexShader (rasteriser.cpp:532-546)
__global__ void vertex_shader_kernel(RasFrame fr, int n, const float* __restrict__ verts, PixelRec* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const RPixel p = vertex_shader(fr, mk3(verts[3 * k], verts[3 * k + 1], verts[3 * k + 2]));
    PixelRec r = {p.x, p.y, p.zinv, p.p.x, p.p.y, p.p.z};
    out[k] = r;
}

// Interpolate (rasteriser.cpp:615-637): a single edge, every sample, all five chains.
__device__ void interpolate_edge(PixelRec a, PixelRec b, int n, PixelRec* out) {
    const float div = (float)max(n - 1, 1);
    const float sx = xdiv((float)(b.x - a.x), div), sy = xdiv((float)(b.y - a.y), div);
    const float sz = xdiv(xsub(b.zinv, a.zinv), div);
    const V3 sp = xdivs3(xsub3(mk3(b.px, b.py, b.pz), mk3(a.px, a.py, a.pz)), div);
    float cx = (float)a.x, cy = (float)a.y, cz = a.zinv;
    V3 cp = mk3(a.px, a.py, a.pz);
    for (int i = 0; i < n; ++i) {
        PixelRec r = {f2i_x86(cx), f2i_x86(cy), cz, cp.x, cp.y, cp.z};
        out[i] = r;
        cx = xadd(cx, sx);
        cy = xadd(cy, sy);
        cz = xadd(cz, sz);
        cp = xadd3(cp, sp);
    }
}

__global__ void interpolate_kernel(PixelRec a, PixelRec b, int n, PixelRec* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) interpolate_edge(a, b, n, out);
}

// ComputePolygonRows (rasteriser.cpp:674-735) for one polygon; edge buffer = scratch of >= rows entries.
__global__ void polygon_rows_kernel(const PixelRec* __restrict__ vp, int maxRows, PixelRec* __restrict__ left,
                                    PixelRec* __restrict__ right, PixelRec* __restrict__ edge, int* __restrict__ rowsOut) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int maxY = max(max(vp[0].y, vp[1].y), vp[2].y), minY = min(min(vp[0].y, vp[1].y), vp[2].y);
    const long long rowsLL = (long long)maxY - (long long)minY + 1;  // :682
    *rowsOut = rowsLL > (long long)maxRows ? -1 : (int)rowsLL;
    if (rowsLL > (long long)maxRows) return;
    const int rows = (int)rowsLL;
    for (int i = 0; i < rows; ++i) {  // :694-698
        PixelRec z = {0, 0, 0.f, 0.f, 0.f, 0.f};
        left[i] = z;
        right[i] = z;
        left[i].x = INT_MAX;
        right[i].x = -INT_MAX;
    }
    for (int i = 0; i < 3; ++i) {  // :705-734
        const int j = (i + 1) % 3;
        PixelRec a = vp[i], b = vp[j];
        a.y -= minY;  // :709-710
        b.y -= minY;
        const int n = abs(vp[i].y - vp[j].y) + 1;  // :712
        interpolate_edge(a, b, n, edge);
        for (int k = 0; k < n; ++k) {
            const PixelRec e = edge[k];
            if (e.x < left[e.y].x) {  // :718
                left[e.y] = e;
                left[e.y].y = e.y + minY;
            }
            if (e.x > right[e.y].x) {  // :726
                right[e.y] = e;
                right[e.y].y = e.y + minY;
            }
        }
    }
}

// PixelShader (rasteriser.cpp:549-589)
__global__ void pixel_shader_kernel(RasFrame fr, int n, const PixelRec* __restrict__ px, const float* __restrict__ colors,
                                    const float* __restrict__ normals, float* __restrict__ outColour,
                                    float* __restrict__ outFocal) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PixelRec p = px[k];
    float focal;
    V3 colour;
    pixel_shader_core(fr, p.zinv, mk3(p.px, p.py, p.pz), mk3(normals[3 * k], normals[3 * k + 1], normals[3 * k + 2]),
                      mk3(colors[3 * k], colors[3 * k + 1], colors[3 * k + 2]), focal, colour);
    outColour[3 * k] = colour.x;
    outColour[3 * k + 1] = colour.y;
    outColour[3 * k + 2] = colour.z;
    outFocal[k] = focal;
}

}  // namespace b2r

using namespace b2r;

namespace {
int sfail(Ctx* c, int code, const char* what) {
    c->err = what;
    return code;
}
int scuda(Ctx* c, cudaError_t e, const char* what) {
    c->err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return B2R_E_CUDA;
}
#define SCU(call, what)                                     \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return scuda(c, e__, what); \
    } while (0)

// Small staging helper: device scratch carved from one growable buffer.
struct Stage {
    Ctx* c;
    size_t off = 0;
    explicit Stage(Ctx* c_) : c(c_) {}
    size_t take(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    }
};

RasFrame ras_frame_of(const Ctx* c) {
    RasFrame r;
    memset(&r, 0, sizeof r);
    const DevFrame& f = c->hostFrame;
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
    return r;
}
}  // namespace


// ---- self-test of the shared-reciprocal division (exact.cuh: Recip / xdiv_by) -------------------------
// Compares xdiv_by(a, recip_make(b)) with div.rn.f32 bit for bit on pseudo-random operand pairs: uniformly random bit
// patterns (every exponent, zeros, denormals, infinities, NaNs), pairs confined to and straddling the [2^-60, 2^60)
// guard, small integers over small integers (the edge and span steps), values near 1 (pos3d / zinv).
__device__ __forceinline__ unsigned mix32(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (unsigned)((z ^ (z >> 31)) >> 16);
}
__device__ __forceinline__ float pattern_operand(unsigned r, unsigned mode) {
    switch (mode) {
        case 0: return __uint_as_float(r);                                                   // anything
        case 1: return __uint_as_float((r & 0x807FFFFFu) | ((67u + (r >> 23) % 120u) << 23));  // inside the guard
        case 2: return __uint_as_float((r & 0x807FFFFFu) | ((60u + (r >> 23) % 16u) << 23));   // around 2^-60
        case 3: return __uint_as_float((r & 0x807FFFFFu) | ((180u + (r >> 23) % 16u) << 23));  // around 2^60
        case 4: return (float)((int)(r % 8193u) - 4096);                                     // small integers
        case 5: return (float)(1 + r % 4096u);                                               // row / step counts
        default: return __uint_as_float((r & 0x007FFFFFu) | 0x3F000000u | (r & 0x80000000u));  // [0.5, 1)
    }
}
__global__ void division_selftest_kernel(unsigned long long n, unsigned seed, unsigned long long* __restrict__ bad,
                                         float* __restrict__ firstBad) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long mism = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned r0 = mix32(i * 3 + seed), r1 = mix32(i * 3 + 1 + seed), r2 = mix32(i * 3 + 2 + seed);
        const unsigned ma = r2 % 7u, mb = (r2 >> 8) % 7u;
        const float a = pattern_operand(r0, ma), b = pattern_operand(r1, mb);
        const float want = __fdiv_rn(a, b);
        const float got = xdiv_by(a, recip_make(b));
        const bool same = __float_as_uint(want) == __float_as_uint(got) || (want != want && got != got);
        if (!same) {
            if (atomicAdd(bad, 1ull) == 0ull) {
                firstBad[0] = a;
                firstBad[1] = b;
                firstBad[2] = want;
                firstBad[3] = got;
            }
            ++mism;
        }
    }
    (void)mism;
}

extern "C" {

int b2r_rt_closest_intersection_batch(b2r_ctx* ctx, int n, const float* starts, const float* dirs, const int32_t* isLight,
                                      b2r_intersection* io, int32_t* hit, float* focal) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    if (n < 0 || (n && (!starts || !dirs || !io))) return sfail(c, B2R_E_INVALID, "closest_intersection_batch: bad arguments");
    if (!c->haveScene || !c->haveFrame) return sfail(c, B2R_E_NO_SCENE, "needs b2r_set_triangles and b2r_set_frame");
    if (n == 0) return B2R_OK;
    Stage st(c);
    const size_t oS = st.take(12 * (size_t)n), oD = st.take(12 * (size_t)n), oL = st.take(4 * (size_t)n),
                 oI = st.take(20 * (size_t)n), oH = st.take(4 * (size_t)n), oF = st.take(4 * (size_t)n);
    SCU(c->subScratch.reserve(st.off), "scratch alloc");
    char* d = c->subScratch.as<char>();
    cudaStream_t s = c->stream;
    SCU(cudaMemcpyAsync(d + oS, starts, 12 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    SCU(cudaMemcpyAsync(d + oD, dirs, 12 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    if (isLight) SCU(cudaMemcpyAsync(d + oL, isLight, 4 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    SCU(cudaMemcpyAsync(d + oI, io, 20 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    closest_intersection_kernel<<<(n + 127) / 128, 128, 0, s>>>(c->geom.as<float4>(), c->T, c->frame.as<DevFrame>(), n,
                                                              (const float*)(d + oS), (const float*)(d + oD),
                                                              isLight ? (const int32_t*)(d + oL) : nullptr,
                                                              (b2r_intersection*)(d + oI), (int32_t*)(d + oH), (float*)(d + oF));
    c->launches++;
    SCU(cudaGetLastError(), "closest_intersection_kernel");
    SCU(cudaMemcpyAsync(io, d + oI, 20 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    if (hit) SCU(cudaMemcpyAsync(hit, d + oH, 4 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    if (focal) SCU(cudaMemcpyAsync(focal, d + oF, 4 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "closest_intersection_batch");
    return B2R_OK;
}

int b2r_rt_direct_light_batch(b2r_ctx* ctx, int n, const b2r_intersection* hits, float* out3) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    if (n < 0 || (n && (!hits || !out3))) return sfail(c, B2R_E_INVALID, "direct_light_batch: bad arguments");
    if (!c->haveScene || !c->haveFrame) return sfail(c, B2R_E_NO_SCENE, "needs b2r_set_triangles and b2r_set_frame");
    if (n == 0) return B2R_OK;
    Stage st(c);
    const size_t oI = st.take(20 * (size_t)n), oO = st.take(12 * (size_t)n);
    SCU(c->subScratch.reserve(st.off), "scratch alloc");
    char* d = c->subScratch.as<char>();
    cudaStream_t s = c->stream;
    SCU(cudaMemcpyAsync(d + oI, hits, 20 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    direct_light_kernel<<<(n + 127) / 128, 128, 0, s>>>(c->geom.as<float4>(), c->T, c->frame.as<DevFrame>(), n,
                                                      (const b2r_intersection*)(d + oI), (float*)(d + oO));
    c->launches++;
    SCU(cudaGetLastError(), "direct_light_kernel");
    SCU(cudaMemcpyAsync(out3, d + oO, 12 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "direct_light_batch");
    return B2R_OK;
}

int b2r_ras_vertex_shader_batch(b2r_ctx* ctx, int n, const float* verts3, void* pixels24) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    if (n < 0 || (n && (!verts3 || !pixels24))) return sfail(c, B2R_E_INVALID, "vertex_shader_batch: bad arguments");
    if (!c->haveFrame) return sfail(c, B2R_E_NO_SCENE, "needs b2r_set_frame");
    if (n == 0) return B2R_OK;
    Stage st(c);
    const size_t oV = st.take(12 * (size_t)n), oP = st.take(24 * (size_t)n);
    SCU(c->subScratch.reserve(st.off), "scratch alloc");
    char* d = c->subScratch.as<char>();
    cudaStream_t s = c->stream;
    SCU(cudaMemcpyAsync(d + oV, verts3, 12 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    vertex_shader_kernel<<<(n + 127) / 128, 128, 0, s>>>(ras_frame_of(c), n, (const float*)(d + oV), (PixelRec*)(d + oP));
    c->launches++;
    SCU(cudaGetLastError(), "vertex_shader_kernel");
    SCU(cudaMemcpyAsync(pixels24, d + oP, 24 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "vertex_shader_batch");
    return B2R_OK;
}

int b2r_ras_interpolate(b2r_ctx* ctx, const void* a24, const void* b24, int n, void* out24) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    if (n < 0 || n > (1 << 22) || !a24 || !b24 || (n && !out24)) return sfail(c, B2R_E_INVALID, "interpolate: bad arguments");
    if (n == 0) return B2R_OK;
    SCU(c->subScratch.reserve(24 * (size_t)n), "scratch alloc");
    PixelRec a, b;
    memcpy(&a, a24, 24);
    memcpy(&b, b24, 24);
    cudaStream_t s = c->stream;
    interpolate_kernel<<<1, 32, 0, s>>>(a, b, n, c->subScratch.as<PixelRec>());
    c->launches++;
    SCU(cudaGetLastError(), "interpolate_kernel");
    SCU(cudaMemcpyAsync(out24, c->subScratch.p, 24 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "interpolate");
    return B2R_OK;
}

int b2r_ras_compute_polygon_rows(b2r_ctx* ctx, const void* vertexPixels3x24, void* left24, void* right24, int maxRows,
                                 int* rows) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    if (!vertexPixels3x24 || !left24 || !right24 || !rows || maxRows < 1 || maxRows > (1 << 22))
        return sfail(c, B2R_E_INVALID, "compute_polygon_rows: bad arguments");
    Stage st(c);
    const size_t oV = st.take(72), oL = st.take(24 * (size_t)maxRows), oR = st.take(24 * (size_t)maxRows),
                 oE = st.take(24 * (size_t)maxRows), oN = st.take(4);
    SCU(c->subScratch.reserve(st.off), "scratch alloc");
    char* d = c->subScratch.as<char>();
    cudaStream_t s = c->stream;
    SCU(cudaMemcpyAsync(d + oV, vertexPixels3x24, 72, cudaMemcpyHostToDevice, s), "H2D");
    polygon_rows_kernel<<<1, 32, 0, s>>>((const PixelRec*)(d + oV), maxRows, (PixelRec*)(d + oL), (PixelRec*)(d + oR),
                                        (PixelRec*)(d + oE), (int*)(d + oN));
    c->launches++;
    SCU(cudaGetLastError(), "polygon_rows_kernel");
    int n = 0;
    SCU(cudaMemcpyAsync(&n, d + oN, 4, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "compute_polygon_rows");
    if (n < 0) return sfail(c, B2R_E_CAPACITY, "compute_polygon_rows: polygon has more rows than maxRows");
    SCU(cudaMemcpy(left24, d + oL, 24 * (size_t)n, cudaMemcpyDeviceToHost), "D2H");
    SCU(cudaMemcpy(right24, d + oR, 24 * (size_t)n, cudaMemcpyDeviceToHost), "D2H");
    *rows = n;
    return B2R_OK;
}

int b2r_ras_pixel_shader_batch(b2r_ctx* ctx, int n, const void* pixels24, const float* colors3, const float* normals3,
                               float* outColours3, float* outFocal) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    if (n < 0 || (n && (!pixels24 || !colors3 || !normals3 || !outColours3 || !outFocal)))
        return sfail(c, B2R_E_INVALID, "pixel_shader_batch: bad arguments");
    if (!c->haveFrame) return sfail(c, B2R_E_NO_SCENE, "needs b2r_set_frame");
    if (n == 0) return B2R_OK;
    Stage st(c);
    const size_t oP = st.take(24 * (size_t)n), oC = st.take(12 * (size_t)n), oN = st.take(12 * (size_t)n),
                 oO = st.take(12 * (size_t)n), oF = st.take(4 * (size_t)n);
    SCU(c->subScratch.reserve(st.off), "scratch alloc");
    char* d = c->subScratch.as<char>();
    cudaStream_t s = c->stream;
    SCU(cudaMemcpyAsync(d + oP, pixels24, 24 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    SCU(cudaMemcpyAsync(d + oC, colors3, 12 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    SCU(cudaMemcpyAsync(d + oN, normals3, 12 * (size_t)n, cudaMemcpyHostToDevice, s), "H2D");
    pixel_shader_kernel<<<(n + 127) / 128, 128, 0, s>>>(ras_frame_of(c), n, (const PixelRec*)(d + oP), (const float*)(d + oC),
                                                      (const float*)(d + oN), (float*)(d + oO), (float*)(d + oF));
    c->launches++;
    SCU(cudaGetLastError(), "pixel_shader_kernel");
    SCU(cudaMemcpyAsync(outColours3, d + oO, 12 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaMemcpyAsync(outFocal, d + oF, 4 * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "pixel_shader_batch");
    return B2R_OK;
}

int b2r_selftest_division(b2r_ctx* ctx, unsigned long long n, unsigned seed, unsigned long long* mismatches, float* firstBad4) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !mismatches) return B2R_E_INVALID;
    SCU(cudaSetDevice(c->device), "cudaSetDevice");
    SCU(c->subScratch.reserve(256), "scratch alloc");
    char* d = c->subScratch.as<char>();
    cudaStream_t s = c->stream;
    SCU(cudaMemsetAsync(d, 0, 64, s), "clear");
    division_selftest_kernel<<<c->smCount * 8, 256, 0, s>>>(n, seed, (unsigned long long*)d, (float*)(d + 16));
    c->launches++;
    SCU(cudaGetLastError(), "division_selftest_kernel");
    unsigned long long hostBad = 0;
    float fb[4] = {0, 0, 0, 0};
    SCU(cudaMemcpyAsync(&hostBad, d, 8, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaMemcpyAsync(fb, d + 16, 16, cudaMemcpyDeviceToHost, s), "D2H");
    SCU(cudaStreamSynchronize(s), "division selftest");
    *mismatches = hostBad;
    if (firstBad4) memcpy(firstBad4, fb, sizeof fb);
    return B2R_OK;
}

}  // extern "C"
