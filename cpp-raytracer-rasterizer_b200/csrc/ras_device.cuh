// ras_device.cuh -- device functions of the rasteriser shared by the frame pipelines (ras_sortlast.cu, ras_tiles.cu) and the
// sub-stage entry points (substage_kernels.cu).  Reference-order arithmetic only (see exact.cuh).
#pragma once
#include <limits.h>

#include "b2r_internal.h"
#include "exact.cuh"

namespace b2r {

// (int)float with x86 cvttss2si semantics (out of range / NaN -> INT_MIN), which is
// what the reference's int(...) conversions do on the CPU it was built for.
__device__ __forceinline__ int f2i_x86(float f) {
    return (f >= -2147483648.0f && f < 2147483648.0f) ? __float2int_rz(f) : INT_MIN;
}

// IEEE a/b for the edge-step set-up, where a is very often exactly 0 (axis-aligned edges of a regular mesh).
// __fdiv_rn's fast path bails out to a ~100-instruction routine for a zero numerator, and a warp pays for it
// if any lane does; 0/b is +-0 with the XOR of the signs for every finite non-zero or infinite b.
__device__ __forceinline__ float xdiv_step(float a, float b) {
    if (b == 1.0f) return a;  // two-row edges: x/1 == x
    if (a == 0.0f && b != 0.0f && b == b)
        return __int_as_float((__float_as_int(a) ^ __float_as_int(b)) & 0x80000000);
    return xdiv(a, b);
}

// the same with the reciprocal of b shared between several numerators (see Recip in exact.cuh)
__device__ __forceinline__ float xdiv_step_by(float a, const Recip& d) {
    if (d.b == 1.0f) return a;
    if (a == 0.0f && d.b != 0.0f && d.b == d.b)
        return __int_as_float((__float_as_int(a) ^ __float_as_int(d.b)) & 0x80000000);
    return xdiv_by(a, d);
}

// Branch-free form for a group of step divisions by one small positive integer b (edge steps :622-624, span steps
// :648-649): the fast sequence is evaluated unconditionally -- it is exact for b == 1 (r == 1) -- a zero numerator gets
// its signed zero, and `ok` collects whether every non-zero numerator was inside the fast path's range; the caller
// redoes the group with xdiv_step() when it was not (practically never).
__device__ __forceinline__ float xdiv_step_fast(float a, const Recip& d, bool& ok) {
    const float q0 = __fmul_rn(a, d.r);
    const float q = __fmaf_rn(d.r, __fmaf_rn(-d.b, q0, a), q0);
    const bool zero = a == 0.0f;
    ok = ok && (zero || div_safe(a));
    return zero ? __int_as_float((__float_as_int(a) ^ __float_as_int(d.b)) & 0x80000000) : q;
}

struct RPixel {  // == struct Pixel (rasteriser TestModel.h:34-53)
    int x, y;
    float zinv;
    V3 p;
};

// f: frame constants in the kernel-parameter bank (uniform loads, no global traffic per vertex)
// SHARED: the three divisions by pos.z share one reciprocal (frame pipeline); the plain form is what the sub-stage
// entry points run, so the two are checked against each other through the reference's vectors.
template <bool SHARED = false>
__device__ __forceinline__ RPixel vertex_shader(const RasFrame& f, V3 v) {
    RPixel p;
    V3 pos = xvec_mat(xsub3(v, mk3(f.cam[0], f.cam[1], f.cam[2])), f.R);       // :535
    // :538 pos / pos.z; the z component is x/x == 1.0f exactly for every finite non-zero x
    const bool plain = pos.z != 0.0f && fabsf(pos.z) <= 3.402823466e+38f;
    if (SHARED) {
        const Recip d = recip_make(pos.z);
        p.p = mk3(xdiv_by(pos.x, d), xdiv_by(pos.y, d), plain ? 1.0f : xdiv(pos.z, pos.z));
        p.zinv = xdiv_by(1.0f, d);                                              // :541
    } else {
        p.p = mk3(xdiv(pos.x, pos.z), xdiv(pos.y, pos.z), plain ? 1.0f : xdiv(pos.z, pos.z));
        p.zinv = xdiv(1.0f, pos.z);                                             // :541
    }
    float fx = xmul(f.focal, xmul(pos.x, p.zinv));
    float fy = xmul(f.focal, xmul(pos.y, p.zinv));
    p.x = f2i_x86(xadd(__int2float_rn(f2i_x86(fx)), f.halfW));                 // :544  + (SCREEN_WIDTH / 2.0f)
    p.y = f2i_x86(xadd(__int2float_rn(f2i_x86(fy)), f.halfH));                 // :545
    return p;
}


// PixelShader (rasteriser.cpp:549-589) for one fragment with interpolated zinv / pos3d.
template <bool SHARED = false>
__device__ __forceinline__ void pixel_shader_core(const RasFrame& fr, float zinv, V3 pos3d, V3 normal, V3 color,
                                                  float& focal, V3& colour) {
    const RasFrame* f = &fr;
    const V3 cam = mk3(f->cam[0], f->cam[1], f->cam[2]);
    V3 P;                              // :557 pos3d / zinv
    if (SHARED) {
        const Recip d = recip_make(zinv);
        P = mk3(xdiv_by(pos3d.x, d), xdiv_by(pos3d.y, d), xdiv_by(pos3d.z, d));
    } else {
        P = xdivs3(pos3d, zinv);
    }
    P = xvec_mat(P, f->Rinv);          // :559
    P = xadd3(P, cam);                 // :560
    const V3 dc = xsub3(cam, P);       // glm::distance(pPos3d, cameraPos) = length(cameraPos - pPos3d)
    focal = xsub(xsqrt(xdot3(dc, dc)), f->dofFocal);  // :564-565
    V3 result = mk3(0.f, 0.f, 0.f);
    for (int k = 0; k < f->nLights; ++k) {  // :567-584
        const V3 L = mk3(f->lightPos[k][0], f->lightPos[k][1], f->lightPos[k][2]);
        const V3 dl = xsub3(L, P);
        const float r2 = xdot3(dl, dl);
        const float rr = xsqrt(r2);                        // :575
        const float A = sphere_area(rr);                   // :576
        const V3 lc = mk3(f->lightColor[k][0], f->lightColor[k][1], f->lightColor[k][2]);  // :577
        const V3 rDir = xscale3(dl, xdiv(1.0f, rr));       // :578
        const V3 B = xdivs3_shared(lc, A);                 // :580
        const V3 D = xscale3(B, std_max(xdot3(rDir, normal), 0.0f));  // :582 (normal not re-normalised)
        result = xadd3(result, D);
    }
    const V3 refl = mk3(f->reflectance[0], f->reflectance[1], f->reflectance[2]);
    const V3 ind = mk3(f->indirect[0], f->indirect[1], f->indirect[2]);
    colour = xmul3(xmul3(refl, xadd3(result, ind)), color);  // :587
}

}  // namespace b2r
