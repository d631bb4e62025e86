// rt_kernels.cu -- the raytracer hot path on sm_100a.
//
// Replaces Draw() -> ClosestIntersection() -> DirectLight() of
// raytracer/Source/raytracer.cpp:547-606, :202-257, :265-327 (one thread per
// pixel, the N x N sub-samples looped in-thread because the reference carries
// state from one sub-sample to the next, SURVEY.md 8a-1 quirks ii/iii).
//
// Structure of one CTA (256 threads = 8 warps, each warp an 8x4 pixel tile, persistent over tiles):
//   1. TMA bulk copy (cp.async.bulk + mbarrier) of the scene-static triangle
//      records (v0, e1, e2, e1 x e2, normalised normal, colour; float4 rows)
//      from HBM into shared memory.
//   2. Per (ray origin, triangle) constants computed once per CTA into shared
//      memory: everything in ClosestIntersection that depends only on `start`
//      (b, b x e2, e1 x b, (e1 x e2).b -- raytracer.cpp:218,226-227,231) in
//      reference operation order, plus three conservative filter forms.
//      Origin 0 is the camera, origins 1.. are the light sample positions the
//      shadow rays start from (raytracer.cpp:284-291,310).
//   3. Culling hierarchy, each level only dropping (ray, triangle) pairs the
//      reference is CERTAIN to reject:
//        a. per warp tile: triangles whose filter forms are negative on the
//           whole tile rectangle (all N*N sub-samples at once);
//        b. per warp and light sample: triangles that cannot shadow any hit
//           point of the warp (forms bounded over the box of light->hit vectors);
//        c. per ray: the three forms evaluated with FMAs.
//   4. Every surviving pair runs the reference's arithmetic literally (non-fused
//      mul/add, IEEE div/sqrt), so accepted hits, the closest-hit index, positions
//      and distances carry the reference's bits.
//
// The filter (see filter_forms): with s = sign((e1 x e2).b), the reference accepts
// only if s*d1 >= 0, s*d2 >= 0 and s*(d0-d1-d2) >= 0 up to rounding, where d0,d1,d2
// are its three dot products with -dir (raytracer.cpp:232-234).  Each is linear in
// dir, so it is evaluated with FMAs from pre-scaled coefficients and compared
// against a margin that is >= 4x the worst-case rounding error of BOTH
// evaluations.  A pair is skipped only when one form is below -margin; ties,
// grazing rays, degenerate triangles and non-finite values all fall through to the
// exact path.  B2R_OPT_RT_FILTER=0 switches every level off: results are identical
// (tests/test_rt_parity.py checks this).
#include <float.h>

#include <stdlib.h>

#include "b2r_internal.h"
#include "exact.cuh"
#include "pixel_pack.cuh"
#include "rt_device.cuh"

namespace b2r {

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA), sm_90+/sm_100a.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------
// Scene-static triangle records (once per b2r_set_triangles).
// ---------------------------------------------------------------------------
__global__ void tri_prep_kernel(const unsigned char* __restrict__ raw, int stride, int n,
                                float4* __restrict__ geom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    V3 v0 = mk3(t[0], t[1], t[2]), v1 = mk3(t[3], t[4], t[5]), v2 = mk3(t[6], t[7], t[8]);
    V3 nrm = mk3(t[9], t[10], t[11]), col = mk3(t[12], t[13], t[14]);
    V3 e1 = xsub3(v1, v0);      // raytracer.cpp:216
    V3 e2 = xsub3(v2, v0);      // :217
    V3 nn = xcross3(e1, e2);    // :225
    V3 nh = xnormalize3(nrm);   // :300 (re-normalised on every DirectLight call in the reference)
    float4* g = geom + (size_t)i * kGeomQuads;
    g[0] = make_float4(v0.x, v0.y, v0.z, e1.x);
    g[1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    g[2] = make_float4(e2.z, nn.x, nn.y, nn.z);
    g[3] = make_float4(nh.x, nh.y, nh.z, col.x);
    g[4] = make_float4(col.y, col.z, 0.f, 0.f);
}

cudaError_t launch_tri_prep(Ctx* c, cudaStream_t s) {
    if (c->T == 0) return cudaSuccess;
    tri_prep_kernel<<<(c->T + 255) / 256, 256, 0, s>>>(c->raw.as<unsigned char>(), c->stride, c->T,
                                                       c->geom.as<float4>());
    c->launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Per (origin, triangle) constants.
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Conservative filter forms for one (ray origin, triangle) pair: 9 floats.
//
// Margin scale: coefficients are divided by M = 2^-18 * L1 where
// L1 = |n|_1 + |b x e2|_1 + |e1 x b|_1, so that the margin is exactly
// 1 * |dir|_inf.  Error budget in those units (u = 2^-24):
//   reference's own dot products (5 roundings each)      3u*L1/M          = 0.047
//   its u+v<=1 test on rounded quotients                 4u*L1/M          = 0.063
//   primary rays: dir = fl(cameraRot*d) vs exact R*d     3u*L1/M          = 0.047
//   our coefficients (rounded once) + FMA chain          4u*2^18          = 0.063
// total < 0.25, i.e. the margin of 1 has >= 4x slack.
//
// Primary rays (origin = camera): form k is  E_k(dx,dy) = B*dx + C*dy + A  with the
// camera rotation, the focal length and the margin (A += primaryDmax) folded in;
// out = {B1,C1,A1, B2,C2,A2, B3,C3,A3}.
// Shadow rays (origin = light sample, -dir = rDir, |rDir| <= 1):
// G_k(rDir) = g.rDir + 1.0001; out = {g1.xyz, g2.xyz, g3.xyz}.
// The pair is rejected iff some form is negative (sign bit), never otherwise.
// ---------------------------------------------------------------------------
constexpr float kShadowMargin = 1.0001f;

__host__ __device__ inline void filter_forms(V3 n, V3 be2, V3 e1b, float nb, bool primary, const float* R,
                                             float focal, float primaryDmax, float* out, double* invMout = nullptr) {
    if (invMout) *invMout = 0.0;
    double s = (nb > 0.f) ? 1.0 : ((nb < 0.f) ? -1.0 : 0.0);
    if (fabsf(nb) < 7.9e-31f) s = 0.0;  // 2^-100: t = nb/d0 could flush to +-0, which `t >= 0` accepts
    const double L1 = fabs((double)n.x) + fabs((double)n.y) + fabs((double)n.z) + fabs((double)be2.x) +
                      fabs((double)be2.y) + fabs((double)be2.z) + fabs((double)e1b.x) + fabs((double)e1b.y) +
                      fabs((double)e1b.z);
    bool ok = (s != 0.0) && (L1 > 8.7e-19) && (L1 < 1.1e18);  // 2^-60 .. 2^60; false for NaN/inf
    if (ok) {
        const double invM = 262144.0 / L1;  // 1/M
        if (invMout) *invMout = invM;
        // forms in terms of dir:  c_k . dir  with c_k = -s * vec_k / M  (d_k = vec_k . (-dir), raytracer.cpp:232-234)
        const double c[3][3] = {
            {-s * be2.x * invM, -s * be2.y * invM, -s * be2.z * invM},
            {-s * e1b.x * invM, -s * e1b.y * invM, -s * e1b.z * invM},
            {-s * ((double)n.x - be2.x - e1b.x) * invM, -s * ((double)n.y - be2.y - e1b.y) * invM,
             -s * ((double)n.z - be2.z - e1b.z) * invM}};
        for (int k = 0; k < 3; ++k) {
            if (primary) {
                // dir = col0*dx + col1*dy + col2*focal  (cameraRot*d, raytracer.cpp:579-580)
                out[3 * k] = (float)(c[k][0] * R[0] + c[k][1] * R[1] + c[k][2] * R[2]);
                out[3 * k + 1] = (float)(c[k][0] * R[3] + c[k][1] * R[4] + c[k][2] * R[5]);
                out[3 * k + 2] = (float)((c[k][0] * R[6] + c[k][1] * R[7] + c[k][2] * R[8]) * (double)focal +
                                         (double)primaryDmax);
            } else {
                // in terms of rDir = -dir
                out[3 * k] = (float)(-c[k][0]);
                out[3 * k + 1] = (float)(-c[k][1]);
                out[3 * k + 2] = (float)(-c[k][2]);
            }
            ok = ok && isfinite(out[3 * k]) && isfinite(out[3 * k + 1]) && isfinite(out[3 * k + 2]);
        }
    }
    if (!ok && invMout) *invMout = 0.0;
    if (!ok)  // always a candidate: the exact path decides
        for (int k = 0; k < 3; ++k) {
            out[3 * k] = out[3 * k + 1] = 0.f;
            out[3 * k + 2] = primary ? 1.f : 0.f;
        }
}

struct PixelState {  // == struct Intersection (+ the focalDistances slot), raytracer.cpp:91-96,249
    V3 pos;
    float dist;
    int idx;
    float focal;
};

struct Counters {
    unsigned long long primary = 0, shadow = 0, exact = 0, shadowEval = 0;
};

constexpr int kTileW = 32, kTileH = 8, kThreads = 256;
constexpr unsigned kFull = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------
// Hierarchical culling.  All of it only ever removes (ray, triangle) pairs
// that the per-ray filter would also remove, i.e. pairs the reference is
// certain to reject; every surviving pair still goes through the per-ray
// filter and then the exact reference-order test.
//
// Primary rays: one warp owns an 8x4 pixel tile.  Every sub-sample of every
// pixel of the tile has (dx,dy) inside a rectangle [cx-hx, cx+hx]x[cy-hy, cy+hy]
// (the AA offsets stay within [-0.5, +0.5], raytracer.cpp:564-574,593,596), and
// each filter form E = B*dx + C*dy + A is linear, so
//      max over the tile of E  =  E(cx,cy) + |B|*hx + |C|*hy.
// Lane j evaluates triangle base+j; a triangle whose form maximum is negative
// is rejected for the whole tile and for all N*N sub-samples at once.
// ---------------------------------------------------------------------------
// Form k (0..2) of a pair from the interleaved table written by write_pair_constants.
__device__ __forceinline__ V3 form_of(const float4& f0, const float4& f1, const float4& f2, int k) {
    return k == 0 ? mk3(f0.x, f0.z, f1.x) : (k == 1 ? mk3(f0.y, f0.w, f1.y) : mk3(f1.z, f1.w, f2.x));
}

template <bool FILTER>
__device__ __forceinline__ unsigned primary_tile_mask(const float4* __restrict__ F0, int base, int T, int lane,
                                                      float cx, float cy, float hx, float hy) {
    const int i = base + lane;
    const bool valid = i < T;
    bool keep = valid;
    if (FILTER && valid) {
        const float4* F = F0 + 3 * i;
        const float4 f0 = F[0], f1 = F[1], f2 = F[2];
        // max over the tile of E_k = |B_k|*hx + |C_k|*hy + (B_k*cx + C_k*cy + A_k); forms 1 and 2 packed, form 3 scalar
        const float2 B12 = make_float2(f0.x, f0.y), C12 = make_float2(f0.z, f0.w);
        const float2 e12 = __ffma2_rn(
            make_float2(fabsf(B12.x), fabsf(B12.y)), make_float2(hx, hx),
            __ffma2_rn(make_float2(fabsf(C12.x), fabsf(C12.y)), make_float2(hy, hy),
                       __ffma2_rn(B12, make_float2(cx, cx), __ffma2_rn(C12, make_float2(cy, cy), make_float2(f1.x, f1.y)))));
        const float e3 = fmaf(fabsf(f1.z), hx, fmaf(fabsf(f1.w), hy, fmaf(f1.z, cx, fmaf(f1.w, cy, f2.x))));
        keep = !(e12.x < 0.f) && !(e12.y < 0.f) && !(e3 < 0.f);
    }
    return __ballot_sync(kFull, keep);
}

// Per-ray filter over the triangles of a tile mask (bit j <-> triangle base+j).
template <bool FILTER>
__device__ __forceinline__ unsigned primary_ray_mask(const float4* __restrict__ F0, int base, unsigned tileMask,
                                                     float dx, float dy) {
    if (!FILTER) return tileMask;
    unsigned m = 0u;
    for (unsigned tm = tileMask; tm; tm &= tm - 1) {  // warp-uniform loop
        const int j = __ffs(tm) - 1;
        const float4* F = F0 + 3 * (base + j);
        const float4 f0 = F[0], f1 = F[1], f2 = F[2];
        // E_k = B_k*dx + C_k*dy + A_k; forms 1 and 2 as one packed chain (two FFMA2), form 3 scalar
        const float2 E12 = __ffma2_rn(make_float2(f0.x, f0.y), make_float2(dx, dx),
                                      __ffma2_rn(make_float2(f0.z, f0.w), make_float2(dy, dy), make_float2(f1.x, f1.y)));
        const float E3 = fmaf(f1.z, dx, fmaf(f1.w, dy, f2.x));
        // any sign bit set = some form negative = certain reject
        if ((int)(__float_as_uint(E12.x) | __float_as_uint(E12.y) | __float_as_uint(E3)) >= 0) m |= 1u << j;
    }
    return m;
}

// Shadow rays of one light sample: q = lightPos - hitPos per lane, rDir = q/|q|.  The per-ray forms are
// G_k(rDir) = g_k.rDir + 1.0001 >= 0, equivalently g_k.q + 1.0001*|q| >= 0.  Over the warp's bounding box
// [qlo,qhi] of q and with rb >= |q| for every lane, the form is at most
//      sum_j max(g_kj*qlo_j, g_kj*qhi_j) + 1.0001*rb ;
// a triangle for which some form's bound is negative cannot pass the per-ray filter on any lane.
template <bool FILTER>
__device__ __forceinline__ unsigned shadow_warp_mask(const float4* __restrict__ Fo, int base, int T, int lane,
                                                     V3 qlo, V3 qhi, float rb) {
    const int i = base + lane;
    const bool valid = i < T;
    bool keep = valid;
    if (FILTER && valid) {
        const float4* F = Fo + 3 * i;
        const float4 f0 = F[0], f1 = F[1], f2 = F[2];
        const float slack = kShadowMargin * rb;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const V3 g = form_of(f0, f1, f2, k);
            // g_j * qlo_j and g_j * qhi_j side by side (FMUL2), the larger one bounds the term over the box
            const float2 px = __fmul2_rn(make_float2(g.x, g.x), make_float2(qlo.x, qhi.x));
            const float2 py = __fmul2_rn(make_float2(g.y, g.y), make_float2(qlo.y, qhi.y));
            const float2 pz = __fmul2_rn(make_float2(g.z, g.z), make_float2(qlo.z, qhi.z));
            const float ub = fmaxf(px.x, px.y) + fmaxf(py.x, py.y) + fmaxf(pz.x, pz.y) + slack;
            keep = keep && !(ub < 0.f);
        }
    }
    return __ballot_sync(kFull, keep);
}

// Besides the sign test, a candidate is dropped when its plane is hit CERTAINLY beyond the occlusion threshold
// thr = r*0.99f (raytracer.cpp:313) -- typically the triangle the shaded point itself lies on.  With the scaled,
// sign-normalised denominator Ds = G1+G2+G3 - 3*1.0001 = s*d0/M, the ray parameter of the plane hit is tnum/Ds.
// Only well-conditioned pairs qualify (Ds >= 2^12, i.e. |d0| >= 2^-6*L1): there the reference's u, v, hit position
// and distance are within posErr + 2^-13*t of the exact values (derivation in DESIGN.md 3.1), so its computed
// distance cannot be below thr and the exact test could only say "not an occluder".
template <bool FILTER>
__device__ __forceinline__ unsigned shadow_ray_mask(const float4* __restrict__ Fo, int base, unsigned warpMask, V3 r,
                                                    float thr) {
    if (!FILTER) return warpMask;
    unsigned m = 0u;
    for (unsigned tm = warpMask; tm; tm &= tm - 1) {
        const int j = __ffs(tm) - 1;
        const float4* F = Fo + 3 * (base + j);
        const float4 f0 = F[0], f1 = F[1], f2 = F[2];
        // G_k = g_k . r + margin; forms 1 and 2 as one packed chain (three FFMA2), form 3 scalar
        const float2 G12 = __ffma2_rn(
            make_float2(f0.x, f0.y), make_float2(r.x, r.x),
            __ffma2_rn(make_float2(f0.z, f0.w), make_float2(r.y, r.y),
                       __ffma2_rn(make_float2(f1.x, f1.y), make_float2(r.z, r.z), make_float2(kShadowMargin, kShadowMargin))));
        const float G1 = G12.x, G2 = G12.y;
        const float G3 = fmaf(f1.z, r.x, fmaf(f1.w, r.y, fmaf(f2.x, r.z, kShadowMargin)));
        if ((int)(__float_as_uint(G1) | __float_as_uint(G2) | __float_as_uint(G3)) >= 0) {
            const float Ds = (G1 + G2) + (G3 - 3.0f * kShadowMargin);
            const bool beyond = Ds >= 4096.0f && fmaf(__fdividef(f2.y, Ds), 0.999755859375f /* 1 - 2^-12 */, -f2.z) > thr;
            if (!beyond) m |= 1u << j;
        }
    }
    return m;
}

// Order-preserving float <-> int map so warp min/max can use the integer REDUX unit.
__device__ __forceinline__ int f2ord(float f) {
    const int k = __float_as_int(f);
    return k ^ ((k >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float ord2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7FFFFFFF)); }
__device__ __forceinline__ float warp_min(float v) { return ord2f(__reduce_min_sync(kFull, f2ord(v))); }
__device__ __forceinline__ float warp_max(float v) { return ord2f(__reduce_max_sync(kFull, f2ord(v))); }

// ---------------------------------------------------------------------------
// Per (origin, triangle) constants for one pair: exact part (2 quads) and filter forms (3 quads).
__device__ __forceinline__ void write_pair_constants(const float4* __restrict__ g, float4 og, bool primary,
                                                     const DevFrame* __restrict__ f, bool withForms,
                                                     float4* __restrict__ X, float4* __restrict__ F) {
    const TriG t = load_geom(g);
    const OriginTri c = origin_constants(t.v0, t.e1, t.e2, t.n, mk3(og.x, og.y, og.z));
    X[0] = make_float4(c.be2.x, c.be2.y, c.be2.z, c.nb);
    X[1] = make_float4(c.e1b.x, c.e1b.y, c.e1b.z, 0.f);
    if (withForms) {
        float q[9];
        double invM;
        filter_forms(t.n, c.be2, c.e1b, c.nb, primary, f->R, f->focal, f->primaryDmax, q, &invM);
        // Shadow rays only (see shadow_ray_mask): tnum = |nb|/M so that the ray parameter of the plane hit is
        // tnum / (scaled s*d0); posErr bounds how far the reference's computed hit distance can fall below that.
        float tnum = 0.f, posErr = 0.f;
        if (!primary && invM > 0.0) {
            const double e1n = fabs((double)t.e1.x) + fabs((double)t.e1.y) + fabs((double)t.e1.z) + fabs((double)t.e2.x) +
                               fabs((double)t.e2.y) + fabs((double)t.e2.z);
            const double e0n = fabs((double)t.v0.x) + fabs((double)t.v0.y) + fabs((double)t.v0.z) + fabs((double)og.x) +
                               fabs((double)og.y) + fabs((double)og.z) + e1n;
            tnum = (float)(fabs((double)c.nb) * invM);
            posErr = (float)(e1n * 6.103515625e-5 /* 2^-14 */ + e0n * 9.5367431640625e-7 /* 2^-20 */);
            if (!isfinite(tnum) || !isfinite(posErr)) tnum = posErr = 0.f;
        }
        // forms 1 and 2 interleaved, so that the per-ray filters evaluate them as one packed FMA chain (FFMA2):
        // (f1[0], f2[0], f1[1], f2[1]) (f1[2], f2[2], f3[0], f3[1]) (f3[2], tnum, posErr, -)
        F[0] = make_float4(q[0], q[3], q[1], q[4]);
        F[1] = make_float4(q[2], q[5], q[6], q[7]);
        F[2] = make_float4(q[8], tnum, posErr, 0.f);
    }
}

// Scenes too large for shared memory: the same constants, once per frame, into HBM (read back through L1/L2).
__global__ void __launch_bounds__(256) rt_origin_setup_kernel(const float4* __restrict__ geom, const DevFrame* __restrict__ f,
                                                              int T, int withForms, float4* __restrict__ X,
                                                              float4* __restrict__ F) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= f->nOrigins * T) return;
    const int o = it / T, i = it - o * T;
    const float4 og = make_float4(f->origin[o][0], f->origin[o][1], f->origin[o][2], 0.f);
    write_pair_constants(geom + (size_t)i * kGeomQuads, og, o == 0, f, withForms != 0, X + 2 * (size_t)it, F + 3 * (size_t)it);
}

// ---------------------------------------------------------------------------
// The kernel.  RESIDENT: triangles and constants live in shared memory (staged/computed per CTA); otherwise they
// are read from HBM.  TILECULL = false keeps only the per-ray filter (every ray looks at every triangle).
// ---------------------------------------------------------------------------
// SINGLE (implies RESIDENT): at most 32 triangles, i.e. one chunk.  The tables then sit at compile-time offsets
// (32 triangle slots; per origin one block of 32*2 exact quads followed by 32*3 form quads), so the per-sample code
// carries no address arithmetic and no chunk loops.
constexpr int kSingleBlockQuads = 32 * 5;
// ONE (implies SINGLE): one light, one shadow sample (the reference's defaults); the light loop disappears and the
// light's position and power come from the kernel parameters.
template <bool RESIDENT, bool TILECULL, bool FILTER, bool STATS, bool SINGLE, bool ONE>
__global__ void __launch_bounds__(kThreads, SINGLE ? 4 : 3) rt_trace_shade_kernel(const __grid_constant__ RtLaunch a) {
    static_assert(RESIDENT || !SINGLE, "SINGLE needs the shared-memory tables");
    static_assert(SINGLE || !ONE, "ONE is a specialisation of SINGLE");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int T = a.T;
    const DevFrame* __restrict__ f = a.frame;
    const int nO = ONE ? 2 : a.fr.nOrigins;
    const int nChunks = SINGLE ? 1 : (T + 31) >> 5;
    // quads from one origin's table to the next: exact constants / filter forms
    const int oStrideX = SINGLE ? kSingleBlockQuads : 2 * T, oStrideF = SINGLE ? kSingleBlockQuads : 3 * T;
    const float4 *sG, *sX, *sF;  // triangle records, exact (origin,triangle) constants, filter forms
    float4* sOrg;                // nO ray origins, then nLights light powers, then the per-warp tile lists
    if constexpr (RESIDENT) {
        float4* g = reinterpret_cast<float4*>(smem_raw);                       // T * kGeomQuads
        float4* x = g + (SINGLE ? 32 * kGeomQuads : (size_t)T * kGeomQuads);   // nO * T * 2
        float4* ff = SINGLE ? x + 32 * 2 : x + (size_t)nO * T * 2;             // nO * T * 3
        sOrg = SINGLE ? x + (size_t)nO * kSingleBlockQuads : ff + (size_t)nO * T * 3;
        // 1. triangles: HBM -> shared memory by one bulk async copy (TMA), completion on an mbarrier
        const uint32_t geomBytes = (uint32_t)T * kGeomQuads * 16u;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && geomBytes) {
            mbar_expect_tx(&bar, geomBytes);
            bulk_g2s(g, a.geom, geomBytes, &bar);
        }
        for (int i = threadIdx.x; i < nO; i += kThreads)
            sOrg[i] = make_float4(f->origin[i][0], f->origin[i][1], f->origin[i][2], 0.f);
        if (geomBytes) mbar_wait(&bar, 0);
        __syncthreads();
        // 2. per (origin, triangle) constants, once per CTA
        for (int it = threadIdx.x; it < nO * T; it += kThreads) {
            const int o = it / T, i = it - o * T;
            write_pair_constants(g + (size_t)i * kGeomQuads, sOrg[o], o == 0, f, FILTER, x + (size_t)o * oStrideX + 2 * i,
                                 ff + (size_t)o * oStrideF + 3 * i);
        }
        sG = g;
        sX = x;
        sF = ff;
    } else {
        sG = a.geom;
        sX = a.xconst;
        sF = a.fconst;
        sOrg = reinterpret_cast<float4*>(smem_raw);
        for (int i = threadIdx.x; i < nO; i += kThreads)
            sOrg[i] = make_float4(f->origin[i][0], f->origin[i][1], f->origin[i][2], 0.f);
    }
    float4* sPow = sOrg + nO;
    for (int i = threadIdx.x; i < a.fr.nLights; i += kThreads)
        sPow[i] = make_float4(f->lightPower[i][0], f->lightPower[i][1], f->lightPower[i][2], 0.f);
    uint2* sTileList = reinterpret_cast<uint2*>(sPow + a.fr.nLights);  // 8 warps * nChunks entries (chunk, mask)
    // Shadow-candidate cache, per warp and light sample: 2 quads (box lo, box hi + mask/count) followed, for
    // scenes of more than one 32-triangle chunk, by the (chunk, mask) list.  See "cached shadow candidates" below.
    const int cacheQuads = a.shadowCache ? 2 + (nChunks > 1 ? (nChunks + 1) / 2 : 0) : 0;
    float4* sShadowCache = reinterpret_cast<float4*>(sTileList + (size_t)(kThreads / 32) * (((nChunks + 1) >> 1) << 1));
    __syncthreads();

    const V3 cam = mk3(a.fr.cam[0], a.fr.cam[1], a.fr.cam[2]);
    const float* R = a.fr.R;
    const float dofFocal = a.fr.dofFocal;
    const V3 indirect = mk3(a.fr.indirect[0], a.fr.indirect[1], a.fr.indirect[2]);
    const int N = a.fr.aaN, nLights = ONE ? 1 : a.fr.nLights, samples = ONE ? 1 : a.fr.samples;
    const float halfW = xdiv((float)a.W, 2.0f), halfH = xdiv((float)a.H, 2.0f);  // (float)SCREEN_WIDTH/2.0f :579
    const float stepAA = xdiv(1.0f, (float)(N - 1));                            // :593,596 (+inf when N == 1)
    const float invNN = (float)(N * N);
    // :599 divides by N*N.  For N a power of two x * (1/(N*N)) is the same correctly rounded number as x / (N*N); N = 1
    // needs nothing; otherwise a zero numerator (every pixel without a hit) skips div.rn's slow path: 0/n is 0 itself.
    const bool nnPow2 = (N & (N - 1)) == 0;


    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint2* myTileList = sTileList + warp * nChunks;  // non-empty chunks of this warp's tile
    float4* myCache = sShadowCache + (size_t)warp * (nO - 1) * cacheQuads;
    int nList = 0;
    Counters cnt;
    unsigned tile0 = 0u;  // tile mask of the only chunk when T <= 32

    // Warp tiles (8x4 pixels; 8 of them form a 32x8 block) are handed out dynamically in batches of a.batch, so
    // warps that drew empty or cheap tiles keep going instead of idling until the slowest warp is done.  Every
    // warp's first batch is static (its global index); the following ones come from one atomic each, issued a
    // batch ahead to hide its latency.
    const int numWarpTiles = a.numTiles * (kThreads / 32);
    // a.batch == 0 (small frames): plain static striding, no atomics.
    const bool dynamic = a.batch > 0;
    const int batch = dynamic ? a.batch : 1, firstDynamic = (int)gridDim.x * (kThreads / 32) * batch;
    int base = ((int)blockIdx.x * (kThreads / 32) + warp) * batch;
    auto fetch_batch = [&]() {
        if (!dynamic) return base + firstDynamic;
        int v = 0;
        if (lane == 0) v = firstDynamic + (int)atomicAdd(a.sched, (unsigned)batch);
        return __shfl_sync(kFull, v, 0);
    };
    int nextBase = base < numWarpTiles ? fetch_batch() : numWarpTiles;
    for (; base < numWarpTiles; base = nextBase, nextBase = nextBase < numWarpTiles ? fetch_batch() : numWarpTiles)
    for (int wt = base; wt < min(base + batch, numWarpTiles); ++wt) {
        const int tile = a.tileOrder ? a.numTiles - 1 - (wt >> 3) : (wt >> 3), sub = wt & 7;
        // tile / tilesX without the integer-division sequence: tile < 2^22, so the float quotient is off by one at most
        int ty = __float2int_rz(__int2float_rn(tile) * a.rcpTilesX), tx = tile - ty * a.tilesX;
        if (tx < 0) {
            --ty;
            tx += a.tilesX;
        } else if (tx >= a.tilesX) {
            ++ty;
            tx -= a.tilesX;
        }
        B2R_BOUND(tx, a.tilesX);
        B2R_BOUND(ty * a.tilesX + tx, a.numTiles);
        const int wx0 = tx * kTileW + (sub & 3) * 8,
                  wy0 = a.y0 + (ty * a.tileRowStride + a.tileRowOffset) * kTileH + (sub >> 2) * 4;
        const int x = wx0 + (lane & 7), y = wy0 + (lane >> 3);
        const bool inside = x < a.W && y < a.y1;  // lanes outside stay for the warp collectives
        // Host-buffer draws copy finished row bands out while later ones are still being traced: count this warp
        // tile as finished once its stores are visible device-wide.
        auto signal_done = [&]() {
            if (!a.bandDone) return;
            __threadfence();
            __syncwarp();
            B2R_BOUND(ty / a.bandTileRows, 32);  // kMaxCopyBands
            if (lane == 0) atomicAdd(a.bandDone + ty / a.bandTileRows, 1u);
        };
        if (!__any_sync(kFull, inside)) {
            signal_done();
            continue;
        }

        // tile-level candidate masks for the primary rays of all sub-samples
        {
            // sub-sample coordinates lie in [x-0.5, x+0.5] (AA) or are exactly x; pad for rounding
            const float pad = (N > 1) ? 0.5f + 0.01f : 0.01f;
            const float cx = ((float)wx0 + 3.5f) - halfW, cy = ((float)wy0 + 1.5f) - halfH;
            const float hx = 3.5f + pad, hy = 1.5f + pad;
            nList = 0;
            for (int c = 0; c < nChunks; ++c) {
                unsigned m;
                if (TILECULL) m = primary_tile_mask<FILTER>(sF, c * 32, T, lane, cx, cy, hx, hy);
                else m = (T - c * 32 >= 32) ? kFull : ((1u << (T - c * 32)) - 1u);
                if (nChunks == 1) {
                    tile0 = m;
                    nList = 1;
                } else if (m) {  // m is warp-uniform (a ballot)
                    B2R_BOUND(nList, nChunks);
                    if (lane == 0) myTileList[nList] = make_uint2((unsigned)c, m);
                    ++nList;
                }
            }
            if (nChunks > 1) __syncwarp();
        }
        // No candidate triangle for any sub-sample of any pixel of the tile: every ClosestIntersection call returns
        // false, nothing is shaded and the pixel keeps Update()'s reset values; skip the N*N loop.
        const bool tileEmpty = (nChunks == 1) ? tile0 == 0u : nList == 0;
        if constexpr (STATS) {
            if (tileEmpty && inside) cnt.primary += (unsigned long long)(N * N);  // rays the reference casts
        }
        if (cacheQuads && !tileEmpty) {  // new tile: empty boxes contain nothing, so the first sample rebuilds
            for (int o = lane; o < nO - 1; o += 32) {
                B2R_BOUND(o, nO - 1);
                myCache[(size_t)o * cacheQuads] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f);
                myCache[(size_t)o * cacheQuads + 1] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f);
            }
            __syncwarp();
        }

        PixelState ps;  // Update()'s per-frame reset, raytracer.cpp:335-339 (+P5)
        ps.pos = mk3(0.f, 0.f, 0.f);
        ps.dist = FLT_MAX;
        ps.idx = -1;
        ps.focal = 0.f;
        V3 avg = mk3(0.f, 0.f, 0.f);
        // DirectLight is a pure function of the pixel's carried Intersection (and the frame): a sub-sample whose hit did
        // not replace it (:243) shades the same point of the same triangle as the sub-sample before, to the same bits.
        // lastTerm keeps color * (DirectLight + indirectLight) of the last evaluation; the shading block (with its
        // shadow rays) runs only for lanes whose Intersection changed, and not at all when no lane's did.
        V3 lastTerm = mk3(0.f, 0.f, 0.f);
        float y1 = (N > 1) ? xsub((float)y, 0.5f) : (float)y;  // :564-567
        for (int z = 0; z < (tileEmpty ? 0 : N); ++z) {
            float x1 = (N > 1) ? xsub((float)x, 0.5f) : (float)x;  // :571-574
            for (int z2 = 0; z2 < N; ++z2) {
                // ---- primary ray :579-580
                const float dx = xsub(x1, halfW), dy = xsub(y1, halfH);
                V3 nd = mk3(0.f, 0.f, 0.f);
                bool haveDir = false, any = false, changed = false;
                if (inside) {
                    if constexpr (STATS) cnt.primary++;
                    for (int e = 0; e < nList; ++e) {
                        const uint2 cm = (nChunks > 1) ? myTileList[e] : make_uint2(0u, tile0);
                        const int base = (int)cm.x * 32;
                        unsigned m = primary_ray_mask<FILTER>(sF, base, cm.y, dx, dy);
                        if (m && !haveDir) {  // the direction is only needed by the exact test
                            // cameraRot * vec3(dx, dy, focalLength), the third products (column 2 * focalLength)
                            // taken from the frame constants                                        :580, :229
                            nd = neg3(xadd3(xadd3(xscale3(mk3(R[0], R[1], R[2]), dx), xscale3(mk3(R[3], R[4], R[5]), dy)),
                                            mk3(a.fr.Rf[0], a.fr.Rf[1], a.fr.Rf[2])));  // (col0*dx + col1*dy) + col2*f
                            haveDir = true;
                        }
                        for (; m; m &= m - 1) {  // ascending triangle index
                            const int i = base + __ffs(m) - 1;
                            if constexpr (STATS) cnt.exact++;
                            V3 pos;
                            float dist;
                            if (exact_hit(sG + i * kGeomQuads, sX + 2 * i, cam, nd, pos, dist)) {
                                if (ps.dist >= dist) {  // :243 ties -> later index
                                    ps.pos = pos;
                                    ps.dist = dist;
                                    ps.idx = i;
                                    ps.focal = xsub(dist, dofFocal);  // :249
                                    changed = true;
                                }
                                any = true;  // :251
                            }
                        }
                    }
                }
                // ---- DirectLight :265-327.  The warp stays converged so the shadow rays of all its
                // hit pixels can be culled together; lanes without a hit only take part in the collectives.
                const bool lit = any && (changed || !a.reuseLight);  // this lane's Intersection was replaced: DirectLight has a new argument
                if constexpr (STATS)
                    if (any && !lit) cnt.shadow += (unsigned long long)(nO - 1);  // the reference casts them; same answer as before
                if (__any_sync(kFull, lit)) {
                    V3 nDir = mk3(0.f, 0.f, 0.f), colr = mk3(0.f, 0.f, 0.f);
                    if (lit) {
                        const float4 g3 = sG[ps.idx * kGeomQuads + 3], g4 = sG[ps.idx * kGeomQuads + 4];
                        nDir = mk3(g3.x, g3.y, g3.z);
                        colr = mk3(g3.w, g4.x, g4.y);
                    }
                    V3 result = mk3(0.f, 0.f, 0.f), result2 = mk3(0.f, 0.f, 0.f);
                    // The loops over lights (:279) and their samples (:283) run as one loop over the shadow-ray
                    // origins o = 1 + k*samples + s with running pointers into the per-origin tables.
                    const float4* xs = sX + oStrideX;  // exact constants of origin 1
                    const float4* Fo = sF + oStrideF;  // filter forms of origin 1
                    float4* hdr = myCache;
                    int kLight = 0, sLeft = samples;
                    V3 P = mk3(0.f, 0.f, 0.f);  // (color*intensity)/samples :282,296
                    if (ONE) {
                        P = mk3(a.fr.power0[0], a.fr.power0[1], a.fr.power0[2]);
                    } else if (nLights > 0) {
                        const float4 pw = sPow[0];
                        P = mk3(pw.x, pw.y, pw.z);
                    }
                    bool haveBox = false;  // posLo/posHi: box of the lit lanes' hit points, reduced on first use
                    V3 posLo = mk3(0.f, 0.f, 0.f), posHi = mk3(0.f, 0.f, 0.f);
                    for (int o = 1; o < nO; ++o, xs += oStrideX, Fo += oStrideF, hdr += cacheQuads) {
                        {
                            const float4 og = ONE ? make_float4(a.fr.light0[0], a.fr.light0[1], a.fr.light0[2], 0.f) : sOrg[o];
                            const V3 lpos = mk3(og.x, og.y, og.z);   // :284-291
                            const V3 dv = xsub3(lpos, ps.pos);       // position - i.position
                            const float rr = xdot3(dv, dv);
                            const float r = xsqrt(rr);               // :294
                            const V3 rDir = xscale3(dv, xdiv(1.0f, r));  // :298 normalize
                            V3 D = mk3(0.f, 0.f, 0.f);
                            if (lit) {
                                const float A = sphere_area(r);      // :295
                                const V3 B = xdivs3_shared(P, A);    // :301
                                D = xscale3(B, std_max(xdot3(rDir, nDir), 0.0f));  // :304
                                if constexpr (STATS) cnt.shadow++, cnt.shadowEval++;
                            }
                            // ---- shadow ray from the light towards the surface :307-315
                            // direction -rDir, so -dir == rDir; occluded iff any accepted hit is
                            // closer than r*0.99f (== j.distance < r*0.99f, j the closest hit)
                            const float thr = xmul(r, 0.99f);
                            bool occluded = false;
                            // One candidate chunk: per-ray filter, then the exact test in ascending index.
                            auto shadow_chunk = [&](int base, unsigned wm) {
                                unsigned m = shadow_ray_mask<FILTER>(Fo, base, wm, rDir, thr);
                                for (; m; m &= m - 1) {
                                    const int i = base + __ffs(m) - 1;
                                    if constexpr (STATS) cnt.exact++;
                                    V3 pos;
                                    float dist;
                                    if (exact_hit(sG + i * kGeomQuads, xs + 2 * i, lpos, rDir, pos, dist) && dist < thr) {
                                        occluded = true;
                                        break;
                                    }
                                }
                            };
                            if (!(TILECULL && FILTER)) {
                                for (int c = 0; c < nChunks; ++c) {
                                    const int base = c * 32;
                                    const unsigned wm = (T - base >= 32) ? kFull : ((1u << (T - base)) - 1u);
                                    if (lit && !occluded) shadow_chunk(base, wm);
                                }
                            } else {
                                // Cached shadow candidates.  The warp mask of shadow_warp_mask() is valid for ANY
                                // box that contains the light->hit vectors of the warp's lit lanes.  The N*N
                                // sub-samples of a tile hit almost the same surface points, so the box of one
                                // sample, padded, usually contains those of the following samples: the masks are
                                // kept (per warp and light sample) together with their box and reused while every
                                // lane's vector stays inside -- 6 compares and a vote instead of 6 warp reductions
                                // and the 3-form bound.  A miss rebuilds from the union of the old and the new box.
                                uint2* shList = reinterpret_cast<uint2*>(hdr + 2);
                                unsigned wm0 = 0u;
                                int nSh = 0;
                                bool cached = false;
                                if (cacheQuads) {
                                    const float4 c0 = hdr[0], c1 = hdr[1];
                                    const bool in = !lit || (dv.x >= c0.x && dv.x <= c1.x && dv.y >= c0.y && dv.y <= c1.y &&
                                                             dv.z >= c0.z && dv.z <= c1.z);
                                    cached = __all_sync(kFull, in);
                                    if (cached) {
                                        wm0 = __float_as_uint(c1.w);
                                        nSh = (int)wm0;
                                    }
                                }
                                if (!cached) {
                                    // Box of the light->hit vectors dv = lpos - pos of the lit lanes.  The hit points
                                    // are the same for every light sample of this sub-sample, so their box is
                                    // reduced once (6 REDUX) and each origin derives its own: subtraction is
                                    // monotonic, also after rounding, so lpos - posHi <= dv <= lpos - posLo holds for
                                    // the rounded values the lanes use.
                                    if (!haveBox) {
                                        const float big = 3.0e38f;
                                        posLo = mk3(warp_min(lit ? ps.pos.x : big), warp_min(lit ? ps.pos.y : big), warp_min(lit ? ps.pos.z : big));
                                        posHi = mk3(warp_max(lit ? ps.pos.x : -big), warp_max(lit ? ps.pos.y : -big), warp_max(lit ? ps.pos.z : -big));
                                        haveBox = true;
                                    }
                                    V3 qlo = xsub3(lpos, posHi), qhi = xsub3(lpos, posLo);
                                    if (cacheQuads) {
                                        const float4 c0 = hdr[0], c1 = hdr[1];
                                        const float px = fmaf(0.25f, qhi.x - qlo.x, 9.765625e-4f * fmaxf(fabsf(qlo.x), fabsf(qhi.x)));
                                        const float py = fmaf(0.25f, qhi.y - qlo.y, 9.765625e-4f * fmaxf(fabsf(qlo.y), fabsf(qhi.y)));
                                        const float pz = fmaf(0.25f, qhi.z - qlo.z, 9.765625e-4f * fmaxf(fabsf(qlo.z), fabsf(qhi.z)));
                                        qlo = mk3(fminf(c0.x, qlo.x - px), fminf(c0.y, qlo.y - py), fminf(c0.z, qlo.z - pz));
                                        qhi = mk3(fmaxf(c1.x, qhi.x + px), fmaxf(c1.y, qhi.y + py), fmaxf(c1.z, qhi.z + pz));
                                    }
                                    const float mx = fmaxf(fabsf(qlo.x), fabsf(qhi.x)), my = fmaxf(fabsf(qlo.y), fabsf(qhi.y)),
                                                mz = fmaxf(fabsf(qlo.z), fabsf(qhi.z));
                                    const float rb = sqrtf(fmaf(mx, mx, fmaf(my, my, mz * mz))) * 1.00001f;
                                    if (cacheQuads) __syncwarp();  // every lane has read the old list
                                    for (int c = 0; c < nChunks; ++c) {
                                        const unsigned wm = shadow_warp_mask<FILTER>(Fo, c * 32, T, lane, qlo, qhi, rb);
                                        if (nChunks == 1) {
                                            wm0 = wm;
                                        } else if (!cacheQuads) {
                                            if (lit && !occluded) shadow_chunk(c * 32, wm);
                                        } else if (wm) {
                                            B2R_BOUND(nSh, nChunks);
                                            if (lane == 0) shList[nSh] = make_uint2((unsigned)c, wm);
                                            ++nSh;
                                        }
                                    }
                                    if (cacheQuads) {
                                        if (lane == 0) {
                                            hdr[0] = make_float4(qlo.x, qlo.y, qlo.z, 0.f);
                                            hdr[1] = make_float4(qhi.x, qhi.y, qhi.z, __uint_as_float(nChunks == 1 ? wm0 : (unsigned)nSh));
                                        }
                                        __syncwarp();
                                    }
                                }
                                if (lit) {
                                    if (nChunks == 1) {
                                        shadow_chunk(0, wm0);
                                    } else if (cacheQuads) {
                                        for (int e = 0; e < nSh && !occluded; ++e) {
                                            const uint2 cm = shList[e];
                                            shadow_chunk((int)cm.x * 32, cm.y);
                                        }
                                    }
                                }
                            }
                            if (occluded) D = mk3(0.f, 0.f, 0.f);
                            result = xadd3(result, D);      // :319
                        }
                        if (--sLeft == 0) {                     // last sample of this light
                            result2 = xadd3(result2, result);   // :322 (result is not reset per light)
                            sLeft = samples;
                            if (++kLight < nLights) {
                                const float4 pw = sPow[kLight];
                                P = mk3(pw.x, pw.y, pw.z);
                            }
                        }
                    }
                    if (lit) {
                        const V3 color = xmul3(result2, colr);  // :325-326
                        const V3 Tsum = xadd3(color, indirect); // :586
                        lastTerm = xmul3(colr, Tsum);           // :587-591
                    }
                }
                if (any) {
                    avg = xadd3(avg, lastTerm);  // :587-591
                    x1 = xadd(x1, stepAA);       // :593 -- only after a hit
                }
            }
            y1 = xadd(y1, stepAA);  // :596
        }
        if (inside) {
            if (nnPow2) {  // :599
                if (N > 1) avg = xscale3(avg, a.rcpNN);
            } else {
                avg = mk3(avg.x == 0.0f ? avg.x : xdiv(avg.x, invNN), avg.y == 0.0f ? avg.y : xdiv(avg.y, invNN),
                          avg.z == 0.0f ? avg.z : xdiv(avg.z, invNN));
            }
            const size_t idx = (size_t)y * (size_t)a.W + (size_t)x;  // P2: row stride = width
            if (a.colours) {
                a.colours[3 * idx] = avg.x;
                a.colours[3 * idx + 1] = avg.y;
                a.colours[3 * idx + 2] = avg.z;
            }
            if (a.closest) {
                b2r_intersection* c = a.closest + idx;
                c->position[0] = ps.pos.x;
                c->position[1] = ps.pos.y;
                c->position[2] = ps.pos.z;
                c->distance = ps.dist;
                c->triangleIndex = ps.idx;
            }
            if (a.focal) a.focal[idx] = ps.focal;
            // CalculateDOF without depth of field + PutPixelSDL (:643-651), fused: no second pass over the colours
            if (a.surface) {
                const uint32_t px = inside_border(x, y, a.W, a.H) ? pack_xrgb(avg.x, avg.y, avg.z) : 0u;
                a.surface[idx] = px;
                for (int d = 0; d < a.nPeerSurfaces; ++d) a.peerSurface[d][idx] = px;  // other GPUs' copies, over NVLink
            }
        }
        signal_done();
    }

    // The last CTA of the grid to run out of tiles re-arms the scheduler for the next launch (every other CTA
    // has made its final fetch before it counts itself as finished).
    if (dynamic) __syncthreads();
    if (dynamic && threadIdx.x == 0) {
        const unsigned done = atomicAdd(a.sched + 1, 1u);
        if (done == gridDim.x - 1u) {
            a.sched[0] = 0u;
            a.sched[1] = 0u;
        }
    }

    // Gather to a root: the last CTA of this launch tells the root (one increment, over NVLink when the word lives
    // in a peer's memory) that every pixel of this launch has landed.  Each CTA publishes its own stores first.
    if (a.arrive) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(a.arriveCtr, 1u) == gridDim.x - 1u) {
                *a.arriveCtr = 0u;
                __threadfence_system();
                atomicAdd_system(a.arrive, 1u);
            }
        }
    }

    if constexpr (STATS) {
        const unsigned long long v[4] = {cnt.primary, cnt.shadow, cnt.exact, cnt.shadowEval};
        const int slot[4] = {B2R_STAT_PRIMARY_RAYS, B2R_STAT_SHADOW_RAYS, B2R_STAT_EXACT_TESTS, B2R_STAT_SHADOW_RAYS_EVALUATED};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned long long s = v[k];
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(kFull, s, off);
            if (lane == 0 && s) atomicAdd(a.stats + slot[k], s);
        }
    }
}

// Bytes of the per-warp shadow-candidate cache (0 = not used: too many light samples for the space it would take).
static size_t rt_cache_bytes(int T, int nO) {
    const size_t nChunks = (size_t)(T + 31) / 32;
    const size_t quads = 2 + (nChunks > 1 ? (nChunks + 1) / 2 : 0);
    const size_t bytes = (size_t)(kThreads / 32) * (size_t)(nO > 1 ? nO - 1 : 0) * quads * 16;
    return bytes <= 32 * 1024 ? bytes : 0;
}

static size_t rt_smem_bytes(int T, int nO, int nLights, bool resident, bool cache) {
    const size_t slots = T <= 32 ? 32 : (size_t)T;  // the single-chunk layout has 32 fixed slots
    const size_t quads = (resident ? slots * kGeomQuads + (size_t)nO * slots * 5 : 0) + nO + nLights;
    const size_t nChunks = (size_t)(T + 31) / 32;
    return quads * 16 + (size_t)(kThreads / 32) * ((nChunks + 1) / 2 * 2) * 8 + (cache ? rt_cache_bytes(T, nO) : 0) + 16;
}

constexpr int kMaxDynamicSmem = 200 * 1024;  // launch_rt_trace_shade refuses scenes that need more

__global__ void rt_arrive_kernel(unsigned* arrive) {
    __threadfence_system();
    atomicAdd_system(arrive, 1u);
}

template <bool RESIDENT, bool TILECULL, bool SINGLE = false, bool ONE = false>
static cudaError_t launch_variant(Ctx* c, const RtLaunch& a, size_t smem, cudaStream_t s) {
    auto kern = a.useFilter ? (a.stats ? rt_trace_shade_kernel<RESIDENT, TILECULL, true, true, SINGLE, ONE>
                                       : rt_trace_shade_kernel<RESIDENT, TILECULL, true, false, SINGLE, ONE>)
                            : (a.stats ? rt_trace_shade_kernel<RESIDENT, TILECULL, false, true, SINGLE, ONE>
                                       : rt_trace_shade_kernel<RESIDENT, TILECULL, false, false, SINGLE, ONE>);
    // shared-memory opt-in and occupancy: once per (variant, shared-memory size) and context
    int perSM = 0;
    for (const auto& k : c->rtKernelCache)
        if (k.fn == (const void*)kern && k.smem == smem) perSM = k.perSM;
    cudaError_t e;
    if (perSM == 0) {
        // the opt-in is a per-function, per-device setting shared by every context: always the same upper bound
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynamicSmem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kThreads, smem);
        if (e != cudaSuccess) return e;
        if (perSM < 1) perSM = 1;
        c->rtKernelCache.push_back({(const void*)kern, smem, perSM});
    }
    int grid = c->smCount * perSM;
    if (const char* env = getenv("B2R_RT_CTAS_PER_SM")) {  // diagnostic override
        const int v = atoi(env);
        if (v >= 1 && v <= perSM) grid = c->smCount * v;
    }
    if (grid > a.numTiles) grid = a.numTiles;
    if (grid < 1) {  // nothing to draw: the root still hears from this part
        if (a.arrive) {
            rt_arrive_kernel<<<1, 1, 0, s>>>(a.arrive);
            c->launches++;
        }
        return cudaGetLastError();
    }
    // Warp tiles per scheduler fetch, by the work a tile carries (sub-samples x rays per sub-sample); frames with
    // fewer than 8 tiles per resident warp are strided statically (batch 0).
    RtLaunch b = a;
    const int work = a.fr.aaN * a.fr.aaN * a.fr.nOrigins;
    int batch = work >= 8 ? 1 : (work >= 4 ? 2 : 4);
    // ... unless a tile is heavy (16+ rays per pixel): one part of a frame split over 8 GPUs has ~7 tiles per warp, and
    // strided static assignment left some SMs busy 1.7x as long as others (220 -> 188 us for an eighth of config 3)
    if ((long long)a.numTiles < 8LL * grid && work < 16) batch = 0;
    b.batch = batch;
    kern<<<grid, kThreads, smem, s>>>(b);
    c->launches++;
    return cudaGetLastError();
}

// B2R_OPT_RT_VARIANT: 0 = tile/warp culling + per-ray filter (default); 1 = per-ray filter only;
// 2 = like 0 but constants read from HBM even when they would fit in shared memory (tests the large-scene path);
// 3 = like 0 without the shadow-candidate cache; 5 = like 0, DirectLight evaluated for every hit sub-sample (no reuse).
cudaError_t launch_rt_trace_shade(Ctx* c, const RtLaunch& a0, cudaStream_t s) {
    const DevFrame& f = c->hostFrame;
    RtLaunch a = a0;
    // (one sample per tile: nothing to reuse, the cache would only be rebuilt and stored every time)
    const bool cache = rt_cache_bytes(a.T, f.nOrigins) > 0 && c->optRtVariant != 3 && f.aaN > 1;
    a.shadowCache = cache ? 1 : 0;
    const size_t smemRes = rt_smem_bytes(a.T, f.nOrigins, f.nLights, true, cache);
    const bool resident = smemRes <= 96 * 1024 && c->optRtVariant != 2;
    if (resident) {
        a.xconst = a.fconst = nullptr;
        if (a.T <= 32) {
            if (f.nLights == 1 && f.samples == 1 && f.nOrigins == 2 && (c->optRtVariant == 0 || c->optRtVariant == 4 || c->optRtVariant == 5)) {  // (4 and 5 do not change what this kernel is given)
                for (int i = 0; i < 3; ++i) a.fr.light0[i] = f.origin[1][i], a.fr.power0[i] = f.lightPower[0][i];
                return launch_variant<true, true, true, true>(c, a, smemRes, s);
            }
            return c->optRtVariant == 1 ? launch_variant<true, false, true>(c, a, smemRes, s)
                                        : launch_variant<true, true, true>(c, a, smemRes, s);
        }
        return c->optRtVariant == 1 ? launch_variant<true, false>(c, a, smemRes, s) : launch_variant<true, true>(c, a, smemRes, s);
    }
    const size_t smem = rt_smem_bytes(a.T, f.nOrigins, f.nLights, false, cache);
    if (smem > (size_t)kMaxDynamicSmem) return cudaErrorInvalidConfiguration;  // > ~100k triangles: reported as B2R_E_UNSUPPORTED
    const size_t pairs = (size_t)f.nOrigins * (size_t)a.T;
    cudaError_t e;
    if ((e = c->rtX.reserve(pairs * 32 + 64)) != cudaSuccess) return e;
    if ((e = c->rtF.reserve(pairs * 48 + 64)) != cudaSuccess) return e;
    a.xconst = c->rtX.as<float4>();
    a.fconst = c->rtF.as<float4>();
    rt_origin_setup_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, s>>>(a.geom, a.frame, a.T, a.useFilter, c->rtX.as<float4>(),
                                                                          c->rtF.as<float4>());
    c->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return launch_variant<false, true>(c, a, smem, s);
}

}  // namespace b2r
