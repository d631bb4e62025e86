// rt_device.cuh -- device functions of the raytracer shared by the frame kernel (rt_trace.cu) and the
// sub-stage entry points (substage_kernels.cu).  Reference-order arithmetic only (see exact.cuh).
#pragma once
#include "b2r_internal.h"
#include "exact.cuh"

namespace b2r {

struct TriG {
    V3 v0, e1, e2, n;
};
__device__ __forceinline__ TriG load_geom(const float4* g) {
    float4 a = g[0], b = g[1], c = g[2];
    TriG t;
    t.v0 = mk3(a.x, a.y, a.z);
    t.e1 = mk3(a.w, b.x, b.y);
    t.e2 = mk3(b.z, b.w, c.x);
    t.n = mk3(c.y, c.z, c.w);
    return t;
}
// Everything in ClosestIntersection that depends only on `start` and the triangle
// (raytracer.cpp:218,226-227,231), in reference operation order.
struct OriginTri {
    V3 be2, e1b;
    float nb;
};
__host__ __device__ inline OriginTri origin_constants(V3 v0, V3 e1, V3 e2, V3 n, V3 org) {
    OriginTri o;
    const V3 b = xsub3(org, v0);   // :218
    o.be2 = xcross3(b, e2);        // :226
    o.e1b = xcross3(e1, b);        // :227
    o.nb = xadd(xadd(xmul(n.x, b.x), xmul(n.y, b.y)), xmul(n.z, b.z));  // :231 hand-written dot, left to right
    return o;
}

// ---------------------------------------------------------------------------
// The exact test: ClosestIntersection's loop body, raytracer.cpp:229-252.
// nd = -dir.  Returns true when the reference accepts the hit; pos/dist valid then.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool exact_hit_core(const TriG& t, V3 be2, V3 e1b, float nb, V3 start, V3 nd, V3& pos,
                                               float& dist) {
    float d0 = xdot3(t.n, nd);    // e1e2d :232  ((x*x + y*y) + z*z, left to right)
    float d1 = xdot3(be2, nd);    // be2d  :233
    float d2 = xdot3(e1b, nd);    // e1bd  :234
    float tt = xdiv(nb, d0), u = xdiv(d1, d0), v = xdiv(d2, d0);                       // :237
    if (!(xadd(u, v) <= 1.0f && u >= 0.0f && v >= 0.0f && tt >= 0.0f)) return false;   // :239
    pos = xadd3(xadd3(t.v0, xscale3(t.e1, u)), xscale3(t.e2, v));                      // :241
    V3 dv = xsub3(pos, start);                                                         // glm::distance(start,pos)
    dist = xsqrt(xdot3(dv, dv));                                                       // :242
    return true;
}
// with the (origin, triangle) constants precomputed: g = triangle record, xo = (be2, nb), (e1b, -)
__device__ __forceinline__ bool exact_hit(const float4* g, const float4* xo, V3 start, V3 nd, V3& pos,
                                          float& dist) {
    const TriG t = load_geom(g);
    const float4 q0 = xo[0], q1 = xo[1];
    return exact_hit_core(t, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), q0.w, start, nd, pos, dist);
}


}  // namespace b2r
