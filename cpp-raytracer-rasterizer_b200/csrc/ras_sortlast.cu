// ras_sortlast.cu -- the rasteriser hot path on sm_100a, sort-last form (the default pipeline).
//
// Replaces Draw() -> DrawPolygon() -> VertexShader / ComputePolygonRows /
// Interpolate / DrawRows / DrawLineSDL / Bresenham / PixelShader of
// rasteriser/Source/rasteriser.cpp:461-482, :532-546, :549-589, :592-672, :674-768.
//
// The reference draws triangles serially in index order with a strict
// `zinv > depthBuffer` test against a buffer cleared to 0 (:606, :188).  The
// final image therefore only depends on, per pixel, the fragment with the
// largest zinv, the lowest triangle index among exact ties, and only
// fragments with zinv > 0.  A 64-bit key (zinv bits << 32 | ~index) and an
// atomicMax reproduce that for any execution order; shading is deferred to
// the winner (PixelShader's writes are simply overwritten by later winners in
// the reference, so shading only the final one gives identical arrays).
//
// Every triangle is set up and walked exactly once (one thread per triangle), fragments go to a full-frame key
// buffer in HBM with RED.MAX.64 -- a native, fire-and-forget L2 operation -- and one streaming pass shades the
// winners.  The screen-tile form of the same rule (ras_tiles.cu, B2R_OPT_RAS_VARIANT = 2) keeps keys and spans in
// shared memory and moves a third of the bytes, but re-derives a triangle once per tile it touches and resolves the
// 64-bit order with compare-and-swap loops; on this instruction-bound workload that costs more than it saves
// (DESIGN.md section 3.2 has both sets of measurements).
//
// All arithmetic that decides coverage, depth or colour is in reference order,
// non-fused.  Interpolate's serial float accumulation along each edge (:632-635)
// decides both coverage (via int(current.x)) and the zinv values compared in the
// depth test, so it is replayed step by step, never re-associated.
//
// Pipeline:
//   ras_small     1 thread / triangle.  VertexShader x3.  Triangles of <= 20 rows
//                 (the 1M-triangle regime) are finished here: the three edge walks
//                 update per-row left/right ends kept in shared memory, then every
//                 on-screen fragment goes to the key buffer with atomicMax.  The only
//                 intermediate that reaches HBM is one 32-byte row record per polygon
//                 row that has fragments (fixed slot, no allocation pass): Bresenham's start and
//                 step of pos3d.xy, so that the shade pass interpolates without a division.
//                 Larger triangles are appended to a compact list (setup record + row/edge-sample counts).
//   big path      only for the listed large triangles: exclusive scan of the counts,
//                 ras_edges (1 thread / triangle-edge-chain, stores every edge sample),
//                 ras_rows (1 lane / polygon row; short rows per lane, long rows
//                 cooperatively across the warp).
//   ras_shade     1 thread / 2 pixels, streaming: key -> winner -> its row record ->
//                 PixelShader in reference order, coalesced writes; clears the key.
#include <limits.h>

#include "b2r_internal.h"
#include "exact.cuh"
#include "pixel_pack.cuh"
#include "ras_common.cuh"
#include "ras_device.cuh"

namespace b2r {
namespace sl {

struct TriSetup {  // 24 words
    int vx[3], vy[3];
    float vz[3];
    float vp[9];
    int minY, rows;
    unsigned rowBase, sampleBase;
    int drawn, tri;  // tri = index in the caller's triangle array (draw order)
};

struct EdgeSample {  // 5 words
    int x;
    float zinv;
    float p[3];
};

struct RowRec {  // 12 words: left/right ends of one polygon row (y implied) and Bresenham's step of pos3d.xy (:649)
    int lx, rx;
    float lz, rz;
    float lp[3], rp[3];
    float psx, psy;  // only rows with fragments have them (pixels > 0)
};

constexpr int kMaxRowsPerTriangle = 1 << 22;
constexpr int kCoordLimit = 1 << 24;

// ---- stage 1: setup, classification, and the complete small-triangle path ----------------------
constexpr int kSmallRows = 20;      // triangles up to this many polygon rows are finished by ras_small
constexpr int kWindowRows = 7;      // rows whose ends are held in shared memory at a time (one walk per window)
constexpr int kSmallThreads = 128;


// What PixelShader needs from one polygon row of a small triangle, written by ras_small for the shade pass: 32 bytes at
// the fixed slot (triangle * kSmallRows + row), so no allocation pass is needed.  Bresenham's start and step of
// pos3d.xy (:649, :668) -- the divisions are done once per row here, not once per pixel there; zinv is the key's high
// word.  pos3d.z is omitted: it is pos.z/pos.z == 1.0f exactly for every triangle that passes the coordinate limits,
// and Interpolate's z step is then (1-1)/n == 0, so that chain stays 1.0f.
struct SmallRow {
    int lx, pad0;
    float lpx, lpy, psx, psy;
    int pad1[2];
};

// Depth key: larger zinv wins, then the lower triangle index (the reference's strict `>` in draw order, :606).
// Bit 0 tells the shade pass which kind of row record the winner has; it cannot affect the order because it is
// a function of the triangle index in the bits above it.  Triangle indices are below 2^31.
__device__ __forceinline__ unsigned long long pack_key(float zinv, unsigned tri, unsigned big) {
    return ((unsigned long long)__float_as_uint(zinv) << 32) | (unsigned long long)(((0x7FFFFFFFu - tri) << 1) | big);
}
__device__ __forceinline__ unsigned key_triangle(unsigned long long key) { return 0x7FFFFFFFu - ((unsigned)key >> 1); }
// slot of polygon row y of small triangle tri: a small triangle spans at most kSmallRows consecutive rows,
// so y modulo kSmallRows is unique within it
__device__ __forceinline__ size_t small_row_slot(unsigned tri, int y) {
    int m = y % kSmallRows;
    if (m < 0) m += kSmallRows;
    return (size_t)tri * kSmallRows + (size_t)m;
}

// One edge of Interpolate (:615-637): the x and zinv chains decide coverage and depth, the pos3d.xy chains
// feed PixelShader (pos3d.z stays 1.0f, see SmallRow).
struct EdgeStep {
    int n, sgn;
    float cx, cz, cpx, cpy, sx, sz, spx, spy;
};
__device__ __forceinline__ EdgeStep edge_begin(const RPixel& a, const RPixel& b) {
    EdgeStep e;
    e.n = abs(a.y - b.y) + 1;                            // :712
    e.sgn = (b.y > a.y) - (b.y < a.y);
    const Recip div = recip_make((float)max(e.n - 1, 1));  // :622
    const float dx = (float)(b.x - a.x), dz = xsub(b.zinv, a.zinv);  // Pixel operator-
    const float dpx = xsub(b.p.x, a.p.x), dpy = xsub(b.p.y, a.p.y);
    bool ok = div.ok;
    e.sx = xdiv_step_fast(dx, div, ok);                  // fPixel operator/
    e.sz = xdiv_step_fast(dz, div, ok);
    e.spx = xdiv_step_fast(dpx, div, ok);
    e.spy = xdiv_step_fast(dpy, div, ok);
    if (!ok) {
        e.sx = xdiv_step(dx, div.b);
        e.sz = xdiv_step(dz, div.b);
        e.spx = xdiv_step(dpx, div.b);
        e.spy = xdiv_step(dpy, div.b);
    }
    e.cx = (float)a.x;                                   // fPixel(Pixel&)
    e.cz = a.zinv;
    e.cpx = a.p.x;
    e.cpy = a.p.y;
    return e;
}

__global__ void __launch_bounds__(kSmallThreads, 8) ras_small_kernel(RasLaunch a, unsigned long long* __restrict__ keys,
                                                                   TriSetup* __restrict__ bigTs, uint2* __restrict__ bigCounts,
                                                                   int2* __restrict__ triInfo, SmallRow* __restrict__ rowRec,
                                                                   RasCounters* __restrict__ ctr) {
    // per-thread row ends, [row][field][thread] so that a warp's accesses never conflict
    extern __shared__ int srow[];
    int* const mine = srow + threadIdx.x;
    float* const minef = reinterpret_cast<float*>(mine);
    auto LX = [&](int r) -> int& { return mine[(8 * r + 0) * kSmallThreads]; };
    auto RX = [&](int r) -> int& { return mine[(8 * r + 1) * kSmallThreads]; };
    auto LZ = [&](int r) -> float& { return minef[(8 * r + 2) * kSmallThreads]; };
    auto RZ = [&](int r) -> float& { return minef[(8 * r + 3) * kSmallThreads]; };
    auto LPX = [&](int r) -> float& { return minef[(8 * r + 4) * kSmallThreads]; };
    auto LPY = [&](int r) -> float& { return minef[(8 * r + 5) * kSmallThreads]; };
    auto RPX = [&](int r) -> float& { return minef[(8 * r + 6) * kSmallThreads]; };
    auto RPY = [&](int r) -> float& { return minef[(8 * r + 7) * kSmallThreads]; };

    const int i = blockIdx.x * kSmallThreads + threadIdx.x;
    unsigned long long nTests = 0, nRows = 0, nDrawn = 0;
    // the three vertices: 9 floats at the head of the 64-byte record (60-byte scenes are repacked at upload), fetched
    // next to the isCulled flag (the compiler may still sink them below the test of the flag; measured, that costs nothing)
    float t[12];
    unsigned char isCulled = 1;
    if (i < a.T) {
        isCulled = a.culled ? a.culled[i] : (unsigned char)0;
        const float4* q = reinterpret_cast<const float4*>(a.raw + (size_t)i * 64);
        const float4 q0 = q[0], q1 = q[1], q2 = q[2];
        t[0] = q0.x; t[1] = q0.y; t[2] = q0.z; t[3] = q0.w; t[4] = q1.x; t[5] = q1.y; t[6] = q1.z; t[7] = q1.w;
        t[8] = q2.x; t[9] = q2.y; t[10] = q2.z; t[11] = q2.w;
    }
    if (!isCulled) {  // :470
        RPixel v[3];
        int maxY = INT_MIN, minY = INT_MAX;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = vertex_shader<true>(a.fr, mk3(t[3 * k], t[3 * k + 1], t[3 * k + 2]));  // :760-761
            maxY = max(maxY, v[k].y);
            minY = min(minY, v[k].y);
            bad = bad || v[k].x <= -kCoordLimit || v[k].x >= kCoordLimit || v[k].y <= -kCoordLimit || v[k].y >= kCoordLimit;
        }
        const int rows = maxY - minY + 1;  // :682
        if (bad || rows > kMaxRowsPerTriangle) {
            atomicExch(&ctr->err, 1u);  // the reference would try to allocate/walk an absurd row count
            atomicExch(&ctr->sticky, 1u);
        } else if (rows > kSmallRows) {
            const unsigned samples = (unsigned)(abs(v[0].y - v[1].y) + abs(v[1].y - v[2].y) + abs(v[2].y - v[0].y) + 3);
            const unsigned slot = atomicAdd(&ctr->nBig, 1u);
            atomicAdd(&ctr->bigRows, (unsigned)rows);
            atomicAdd(&ctr->bigSamples, samples);
            TriSetup s;
            for (int k = 0; k < 3; ++k) {
                s.vx[k] = v[k].x;
                s.vy[k] = v[k].y;
                s.vz[k] = v[k].zinv;
                s.vp[3 * k] = v[k].p.x;
                s.vp[3 * k + 1] = v[k].p.y;
                s.vp[3 * k + 2] = v[k].p.z;
            }
            s.minY = minY;
            s.rows = rows;
            s.rowBase = s.sampleBase = 0;
            s.drawn = 1;
            s.tri = i;
            if (a.bandSlots) {
                // fixed-capacity slots: only rows of the band are kept, so a triangle needs at most bandH row
                // records and bandH samples per edge; rows are addressed by y - y0 (no scan, no readback)
                const unsigned bandH = (unsigned)(a.y1 - a.y0);
                s.rowBase = slot * bandH;
                s.sampleBase = slot * 3u * bandH;
                bigTs[slot] = s;
                triInfo[i] = make_int2((int)s.rowBase, a.y0);  // row record of y: rowBase + (y - y0)
            } else {
                bigTs[slot] = s;
                bigCounts[slot] = make_uint2((unsigned)rows, samples);
                triInfo[i] = make_int2(0, minY);  // .x = rowBase once scan_apply knows it: row record of y at rowBase + (y - minY)
            }
            nDrawn = 1;
            nRows = (unsigned long long)rows;
        } else {
            nDrawn = 1;
            nRows = (unsigned long long)rows;
            const int r0 = max(0, a.y0 - minY), r1 = min(rows, a.y1 - minY);  // rows of this band (DrawRows :743)
            // Row ends live in shared memory for kWindowRows rows at a time; a triangle with more rows replays its
            // edge walks once per window (86 % of config 4's triangles need one window).  The accumulation itself
            // always runs over every step -- only the stores are windowed -- so the values are unchanged.
            for (int w0 = 0; w0 < rows; w0 += kWindowRows) {
                const int w1 = min(rows, w0 + kWindowRows);
                const int e0 = max(w0, r0), e1 = min(w1, r1);  // rows of this window that are in the band
                if (e0 >= e1) continue;
                for (int r = 0; r < w1 - w0; ++r) {  // :694-698
                    LX(r) = INT_MAX;
                    RX(r) = -INT_MAX;
                }
                // ComputePolygonRows: edges 0->1, 1->2, 2->0 in order, strict </> so the first edge to
                // reach an extreme x keeps its attributes (:705-733)
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const int j = (e + 1) % 3;
                    EdgeStep st = edge_begin(v[e], v[j]);
                    int r = v[e].y - minY - w0;
                    for (int k = 0; k < st.n; ++k) {  // :626-636, serial accumulation
                        if ((unsigned)r < (unsigned)(w1 - w0)) {
                            // int(current.x): the chain stays within one pixel of the vertex range, which passed
                            // the +-2^24 limit, so the plain truncating conversion equals the x86 one
                            const int x = __float2int_rz(st.cx);
                            if (x < LX(r)) {
                                LX(r) = x;
                                LZ(r) = st.cz;
                                LPX(r) = st.cpx;
                                LPY(r) = st.cpy;
                            }
                            if (x > RX(r)) {
                                RX(r) = x;
                                RZ(r) = st.cz;
                                RPX(r) = st.cpx;
                                RPY(r) = st.cpy;
                            }
                        }
                        st.cx = xadd(st.cx, st.sx);
                        st.cz = xadd(st.cz, st.sz);
                        st.cpx = xadd(st.cpx, st.spx);
                        st.cpy = xadd(st.cpy, st.spy);
                        r += st.sgn;
                    }
                }
                // DrawRows / DrawLineSDL / Bresenham with dy == 0 (:738-753, :592-612, :639-672)
                for (int rr = e0; rr < e1; ++rr) {
                    const int r = rr - w0;
                    const int lx = LX(r), rx = RX(r), pixels = rx - lx;  // :598
                    const float lz = LZ(r), rz = RZ(r);
                    const int i0 = max(0, -lx - 1), i1 = min(pixels, a.W - lx - 1);  // :663 keeps 0 <= x < W
                    if (i1 <= i0) continue;  // no fragment (incl. pixels == 0, whose 0/0 step nobody reads)
                    const Recip fdx = recip_make((float)pixels);
                    const float lpx = LPX(r), lpy = LPY(r);
                    const float dz = xsub(rz, lz), dpx = xsub(RPX(r), lpx), dpy = xsub(RPY(r), lpy);
                    bool ok = fdx.ok;
                    float zstep = xdiv_step_fast(dz, fdx, ok);  // :648 (constant-depth rows: 0/n)
                    float psx = xdiv_step_fast(dpx, fdx, ok), psy = xdiv_step_fast(dpy, fdx, ok);  // :649
                    if (!ok) {
                        zstep = xdiv_step(dz, fdx.b);
                        psx = xdiv_step(dpx, fdx.b);
                        psy = xdiv_step(dpy, fdx.b);
                    }
                    {   // one full 32-byte sector per row with fragments, for the shade pass
                        B2R_BOUND(small_row_slot((unsigned)i, minY + rr), (size_t)a.T * kSmallRows);
                        float4* rec = reinterpret_cast<float4*>(rowRec + small_row_slot((unsigned)i, minY + rr));
                        rec[0] = make_float4(__int_as_float(lx), 0.f, lpx, lpy);
                        rec[1] = make_float4(psx, psy, 0.f, 0.f);
                    }
                    unsigned long long* keyRow = keys + (size_t)(minY + rr - a.y0) * (size_t)a.W;
                    B2R_BOUND(minY + rr - a.y0, a.y1 - a.y0);
                    B2R_BOUND(lx + 1 + i0, a.W);
                    B2R_BOUND(lx + i1, a.W);
                    for (int q = i0; q < i1; ++q) {
                        const float zinv = xadd(lz, xmul(zstep, (float)q));  // :667
                        if (zinv > 0.0f)                                    // :606 against a buffer cleared to 0 (:188)
                            atomicMax(keyRow + (lx + 1 + q), pack_key(zinv, (unsigned)i, 0u));
                    }
                    nTests += (unsigned long long)(i1 - i0);
                }
            }
        }
    }
    if (a.stats) {
        for (int off = 16; off > 0; off >>= 1) {
            nTests += __shfl_xor_sync(0xffffffffu, nTests, off);
            nRows += __shfl_xor_sync(0xffffffffu, nRows, off);
            nDrawn += __shfl_xor_sync(0xffffffffu, nDrawn, off);
        }
        if ((threadIdx.x & 31) == 0) {
            if (nTests) atomicAdd(a.stats + B2R_STAT_RAS_DEPTH_TESTS, nTests);
            if (nRows) atomicAdd(a.stats + B2R_STAT_RAS_ROWS, nRows);
            if (nDrawn) atomicAdd(a.stats + B2R_STAT_RAS_TRIANGLES, nDrawn);
        }
    }
}

// ---- exclusive scan of uint2 (three small kernels) ---------------------------
constexpr int kScanBlock = 1024;

__device__ __forceinline__ uint2 add2(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }

__device__ uint2 block_exclusive_scan(uint2 v, uint2* total) {
    __shared__ uint2 warpSums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 inc = v;
    for (int off = 1; off < 32; off <<= 1) {
        unsigned x = __shfl_up_sync(0xffffffffu, inc.x, off), y = __shfl_up_sync(0xffffffffu, inc.y, off);
        if (lane >= off) inc = add2(inc, make_uint2(x, y));
    }
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint2 w = (lane < (int)(blockDim.x >> 5)) ? warpSums[lane] : make_uint2(0u, 0u);
        uint2 winc = w;
        for (int off = 1; off < 32; off <<= 1) {
            unsigned x = __shfl_up_sync(0xffffffffu, winc.x, off), y = __shfl_up_sync(0xffffffffu, winc.y, off);
            if (lane >= off) winc = add2(winc, make_uint2(x, y));
        }
        warpSums[lane] = make_uint2(winc.x - w.x, winc.y - w.y);  // exclusive
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    uint2 base = warpSums[warp];
    __syncthreads();
    return make_uint2(base.x + inc.x - v.x, base.y + inc.y - v.y);
}

__global__ void scan_blocks_kernel(const uint2* __restrict__ in, uint2* __restrict__ out, uint2* __restrict__ blockSums, int n) {
    __shared__ uint2 total;
    int i = blockIdx.x * kScanBlock + threadIdx.x;
    uint2 v = (i < n) ? in[i] : make_uint2(0u, 0u);
    uint2 ex = block_exclusive_scan(v, &total);
    if (i < n) out[i] = ex;
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

// one block: exclusive scan of the block sums in place; totals[0] = grand total
__global__ void scan_sums_kernel(uint2* __restrict__ blockSums, int nBlocks, uint2* __restrict__ totals) {
    __shared__ uint2 total;
    uint2 carry = make_uint2(0u, 0u);
    for (int base = 0; base < nBlocks; base += kScanBlock) {
        int i = base + threadIdx.x;
        uint2 v = (i < nBlocks) ? blockSums[i] : make_uint2(0u, 0u);
        uint2 ex = block_exclusive_scan(v, &total);
        if (i < nBlocks) blockSums[i] = add2(ex, carry);
        carry = add2(carry, total);
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[0] = carry;
}

__global__ void scan_apply_kernel(const uint2* __restrict__ ex, const uint2* __restrict__ blockSums,
                                  TriSetup* __restrict__ ts, int2* __restrict__ triInfo, int n) {
    int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i >= n) return;
    uint2 o = add2(ex[i], blockSums[blockIdx.x]);
    ts[i].rowBase = o.x;
    ts[i].sampleBase = o.y;
    triInfo[ts[i].tri].x = (int)o.x;  // for the shade pass: one load from the winner's index to its row records
}

// ---- stage 2: Interpolate (:615-637), one thread per (triangle, edge) --------
// One thread per (triangle, edge, chain): the five accumulation chains of an edge (x, zinv, pos3d.xyz) are
// independent of each other, so each runs its own serial loop -- 15 threads per triangle instead of 3 -- and writes
// its field of every sample.  A 2160-row edge is latency-bound: one dependent FADD per step and thread.
// BAND: fixed-capacity slots (see ras_small): the number of listed triangles is read from the device counter, the walk
// still runs over every step (the accumulation is serial) but only samples on rows of the band are stored, at y - y0.
template <bool BAND>
__global__ void ras_edges_kernel(const TriSetup* __restrict__ ts, int T /* listed large triangles (upper bound if BAND) */,
                                 const RasCounters* __restrict__ ctr, EdgeSample* __restrict__ samples,
                                 unsigned* __restrict__ rowOwner, int y0, int y1) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid / 15, rem = gid - 15 * i, e = rem / 5, ch = rem - 5 * e;
    if (i >= (BAND ? (int)ctr->nBig : T)) return;
    const TriSetup s = ts[i];
    if (!s.drawn) return;
    if (!BAND && rem < 5)  // five threads share the owner table of this triangle's rows
        for (int r = rem; r < s.rows; r += 5) rowOwner[s.rowBase + r] = (unsigned)i;
    const int j = (e + 1) % 3;  // :707
    const int n = abs(s.vy[e] - s.vy[j]) + 1;  // :712
    unsigned off = s.sampleBase;
    if (BAND) off += (unsigned)e * (unsigned)(y1 - y0);
    else for (int k = 0; k < e; ++k) off += (unsigned)(abs(s.vy[k] - s.vy[(k + 1) % 3]) + 1);
    const float div = (float)max(n - 1, 1);  // :622
    // Pixel operator- (TestModel.h:82-85) then fPixel operator/ (:124-127); fPixel(Pixel&) for the start value
    float cur, step;
    if (ch == 0) {
        cur = (float)s.vx[e];
        step = xdiv_step((float)(s.vx[j] - s.vx[e]), div);
    } else if (ch == 1) {
        cur = s.vz[e];
        step = xdiv_step(xsub(s.vz[j], s.vz[e]), div);
    } else {
        cur = s.vp[3 * e + (ch - 2)];
        step = xdiv_step(xsub(s.vp[3 * j + (ch - 2)], cur), div);
    }
    float* out = reinterpret_cast<float*>(samples + off) + ch;  // field ch of sample 0; samples are 5 words apart
    if (BAND) {
        // step k lies on row vy[e] + sgn*k: the steps before the band only accumulate, the steps inside it are
        // stored, the steps after it are not needed by anyone
        const int ya = s.vy[e], sgn = (s.vy[j] > ya) - (s.vy[j] < ya);
        int kLo, kHi;
        if (sgn > 0) {
            kLo = y0 - ya;
            kHi = y1 - 1 - ya;
        } else if (sgn < 0) {
            kLo = ya - (y1 - 1);
            kHi = ya - y0;
        } else {
            kLo = (ya >= y0 && ya < y1) ? 0 : 1;
            kHi = 0;
        }
        kLo = max(kLo, 0);
        kHi = min(kHi, n - 1);
        if (kLo > kHi) return;
        for (int k = 0; k < kLo; ++k) cur = xadd(cur, step);  // :626-636 -- serial float accumulation, order matters
        out += 5 * (ya + sgn * kLo - y0);
        const int stride = 5 * sgn;
        if (ch == 0) {
            for (int k = kLo; k <= kHi; ++k, out += stride) {
                *reinterpret_cast<int*>(out) = f2i_x86(cur);
                cur = xadd(cur, step);
            }
        } else {
            for (int k = kLo; k <= kHi; ++k, out += stride) {
                *out = cur;
                cur = xadd(cur, step);
            }
        }
    } else if (ch == 0) {
        for (int k = 0; k < n; ++k, out += 5) {  // :626-636 -- serial float accumulation, order matters
            *reinterpret_cast<int*>(out) = f2i_x86(cur);
            cur = xadd(cur, step);
        }
    } else {
        for (int k = 0; k < n; ++k, out += 5) {
            *out = cur;
            cur = xadd(cur, step);
        }
    }
}

// ---- stage 3: ComputePolygonRows' per-row resolve (:716-733) + DrawRows/DrawLineSDL/Bresenham ----
// bandH > 0: band-slot layout, edge e's sample of row y sits at sampleBase + e*bandH + (y - y0)
__device__ __forceinline__ RowRec resolve_row(const TriSetup& s, const EdgeSample* __restrict__ samples, int y,
                                              int bandH = 0, int y0 = 0) {
    RowRec r;
    r.lx = INT_MAX;    // :696
    r.rx = -INT_MAX;   // :697
    r.lz = r.rz = 0.f;
    r.lp[0] = r.lp[1] = r.lp[2] = r.rp[0] = r.rp[1] = r.rp[2] = 0.f;
    r.psx = r.psy = 0.f;
    unsigned off = s.sampleBase;
    for (int e = 0; e < 3; ++e) {  // edge order 0->1, 1->2, 2->0 (:705-707)
        const int j = (e + 1) % 3;
        const int ya = s.vy[e], yb = s.vy[j];
        const int n = abs(ya - yb) + 1;
        if (y >= min(ya, yb) && y <= max(ya, yb)) {
            const EdgeSample q = bandH > 0 ? samples[s.sampleBase + (unsigned)(e * bandH + (y - y0))]
                                           : samples[off + (unsigned)abs(y - ya)];
            if (q.x < r.lx) {  // :718 strict: the first edge to reach an extreme keeps its attributes
                r.lx = q.x;
                r.lz = q.zinv;
                r.lp[0] = q.p[0]; r.lp[1] = q.p[1]; r.lp[2] = q.p[2];
            }
            if (q.x > r.rx) {  // :726
                r.rx = q.x;
                r.rz = q.zinv;
                r.rp[0] = q.p[0]; r.rp[1] = q.p[1]; r.rp[2] = q.p[2];
            }
        }
        off += (unsigned)n;
    }
    return r;
}

// fragments i in [i0,i1) of one row; Bresenham with dy == 0 (:639-672): x = lx+1+i, zinv = lz + zstep*float(i)
__device__ __forceinline__ void raster_span(unsigned long long* __restrict__ keyRow, int lx, float lz, float zstep,
                                            unsigned tri, int i0, int i1, int istride) {
    for (int i = i0; i < i1; i += istride) {
        const float zinv = xadd(lz, xmul(zstep, (float)i));  // :667
        if (zinv > 0.0f)                                     // :606 against a buffer cleared to 0 (:188)
            atomicMax(keyRow + (lx + 1 + i), pack_key(zinv, tri, 1u));
    }
}

constexpr int kShortRow = 8;


template <bool BAND>
__global__ void __launch_bounds__(256) ras_rows_kernel(const TriSetup* __restrict__ ts,
                                                       const RasCounters* __restrict__ ctr,
                                                       const EdgeSample* __restrict__ samples,
                                                       const unsigned* __restrict__ rowOwner, unsigned nRows,
                                                       RowRec* __restrict__ rows,
                                                       unsigned long long* __restrict__ keys, int W, int y0,
                                                       int y1, unsigned long long* __restrict__ stats, int rowsPerWarp) {
    // A warp takes rowsPerWarp (a power of two, 1..32) consecutive polygon rows, one per low lane; the other lanes
    // only help with the long rows.  Few rows per warp when there are few rows in all (30 large triangles): the long
    // rows of a warp are walked one after the other, so the number of warps is what spreads them over the SMs.
    const int lane = threadIdx.x & 31;
    const unsigned warpId = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool rowLane = lane < rowsPerWarp;
    const unsigned rid = warpId * (unsigned)rowsPerWarp + (unsigned)lane;
    int lx = 0, pixels = 0, i0 = 0, i1 = 0, y = 0;
    float lz = 0.f, zstep = 0.f;
    unsigned tri = 0;
    const int bandH = y1 - y0;
    // BAND: thread rid <-> (slot rid / bandH, row y0 + rid % bandH); otherwise the rows of the listed triangles are
    // packed and rowOwner names the triangle
    const unsigned slot = BAND ? rid / (unsigned)bandH : 0u;
    if (rowLane && (BAND ? slot < ctr->nBig : rid < nRows)) {
        const TriSetup s = ts[BAND ? slot : rowOwner[rid]];
        tri = (unsigned)s.tri;
        y = BAND ? y0 + (int)(rid - slot * (unsigned)bandH) : s.minY + (int)(rid - s.rowBase);
        // DrawRows (:743): rows with y outside the screen are skipped; outside the band: another GPU's
        if (y >= y0 && y < y1 && y >= s.minY && y < s.minY + s.rows) {
            RowRec r = BAND ? resolve_row(s, samples, y, bandH, y0) : resolve_row(s, samples, y);
            lx = r.lx;
            lz = r.lz;
            pixels = r.rx - r.lx;                              // :598
            i0 = max(0, -lx - 1);                              // :663 x >= 0
            i1 = min(pixels, W - lx - 1);                      //      x <  W
            if (i1 < i0) i1 = i0;
            if (i1 > i0) {
                const float fdx = (float)pixels;
                zstep = xdiv_step(xsub(r.rz, r.lz), fdx);      // :648
                r.psx = xdiv_step(xsub(r.rp[0], r.lp[0]), fdx);  // :649, once per row instead of once per shaded pixel
                r.psy = xdiv_step(xsub(r.rp[1], r.lp[1]), fdx);
            }
            B2R_BOUND(rid, BAND ? (unsigned long long)ctr->nBig * (unsigned)bandH : (unsigned long long)nRows);
            rows[rid] = r;
        }
    }
    const int count = i1 - i0;
    if (stats) {
        unsigned long long c = (unsigned long long)count;
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (lane == 0 && c) atomicAdd(stats + B2R_STAT_RAS_DEPTH_TESTS, c);
    }
    unsigned long long* keyRow = keys + (size_t)(y - y0) * (size_t)W;
    // short rows: each lane walks its own; long rows: the whole warp walks them one at a time
    const bool isLong = count > kShortRow;
    if (count > 0 && !isLong) raster_span(keyRow, lx, lz, zstep, tri, i0, i1, 1);
    unsigned longMask = __ballot_sync(0xffffffffu, isLong);
    while (longMask) {
        const int src = __ffs(longMask) - 1;
        longMask &= longMask - 1;
        const int blx = __shfl_sync(0xffffffffu, lx, src);
        const float blz = __shfl_sync(0xffffffffu, lz, src);
        const float bzs = __shfl_sync(0xffffffffu, zstep, src);
        const unsigned btri = __shfl_sync(0xffffffffu, tri, src);
        const int bi0 = __shfl_sync(0xffffffffu, i0, src), bi1 = __shfl_sync(0xffffffffu, i1, src);
        const int by = __shfl_sync(0xffffffffu, y, src);
        raster_span(keys + (size_t)(by - y0) * (size_t)W, blx, blz, bzs, btri, bi0 + lane, bi1, 32);
    }
}

// ---- stage 4: PixelShader (:549-589) for the depth winner of every pixel ------------------------
// Loads of one pixel's winner: Bresenham's start and step on its row, and the triangle's normal/colour.
struct ShadeIn {
    int lx;
    float lpx, lpy, psx, psy;
    V3 normal, color;
};
__device__ __forceinline__ ShadeIn shade_fetch(const RasLaunch& a, unsigned long long key, int y,
                                               const TriSetup* __restrict__ bigTs, const int2* __restrict__ triInfo,
                                               const SmallRow* __restrict__ rowRec, const RowRec* __restrict__ rows) {
    ShadeIn in;
    const unsigned tri = key_triangle(key);
    if (key & 1ull) {
        const int2 info = triInfo[tri];  // (index of the triangle's first row record, the y of that record)
        B2R_BOUND(y - info.y, 1 << 22);  // kMaxRowsPerTriangle: the winner's triangle has a row at y
        const RowRec* r = rows + ((unsigned)info.x + (unsigned)(y - info.y));
        const float4 q0 = *reinterpret_cast<const float4*>(r);          // lx, rx, lz, rz
        const float2 q1 = *reinterpret_cast<const float2*>(r->lp);      // lp.x, lp.y
        const float2 q2 = *reinterpret_cast<const float2*>(&r->psx);    // psx, psy
        in.lx = __float_as_int(q0.x);
        in.lpx = q1.x;
        in.lpy = q1.y;
        in.psx = q2.x;
        in.psy = q2.y;
    } else {
        const float4* q = reinterpret_cast<const float4*>(rowRec + small_row_slot(tri, y));
        const float4 q0 = q[0], q1 = q[1];
        in.lx = __float_as_int(q0.x);
        in.lpx = q0.z;
        in.lpy = q0.w;
        in.psx = q1.x;
        in.psy = q1.y;
    }
    const float4* t = reinterpret_cast<const float4*>(a.raw + (size_t)tri * 64);  // 64-byte records (see b2r_set_triangles)
    const float4 n4 = t[2], c4 = t[3];
    in.normal = mk3(n4.y, n4.z, n4.w);
    in.color = mk3(c4.x, c4.y, c4.z);
    return in;
}

// PixelShader (:549-589) for the fragment of pixel x on the winner's row; zinv is the key's high word (:667).
__device__ __forceinline__ void shade_pixel(const RasFrame& fr, const ShadeIn& in, int x, float zinv, float& focal, V3& colour) {
    const float fi = (float)(x - in.lx - 1);
    // pos3d.z: the difference of the row ends is exactly 0 (pos3d.z == 1), so the chain stays 1.0f
    const V3 pos3d = mk3(xadd(in.lpx, xmul(in.psx, fi)), xadd(in.lpy, xmul(in.psy, fi)), 1.0f);  // :649,668
    pixel_shader_core<true>(fr, zinv, pos3d, in.normal, in.color, focal, colour);
}

// kShadePixels pixels per thread (kShadeThreads apart in x, so every access stays coalesced): the key loads of all of them
// are issued first, then all row-record / triangle loads, then the arithmetic -- the kernel is bound by the
// latency of that dependent load chain, not by bandwidth.
constexpr int kShadePixels = 2;
// Small thread blocks: a block's warps all sit in the same phase (keys, gathers, arithmetic, stores), so the more blocks
// an SM holds, the more evenly the phases overlap (config 4: 512 threads 0.196 ms, 256: 0.190, 128: 0.185, 64: 0.184,
// 32: 0.190).
constexpr int kShadeThreads = 64;

__global__ void __launch_bounds__(kShadeThreads, 1024 / kShadeThreads) ras_shade_kernel(RasLaunch a, const TriSetup* __restrict__ bigTs,
                                                        const int2* __restrict__ triInfo,
                                                        const SmallRow* __restrict__ rowRec,
                                                        const RowRec* __restrict__ rows,
                                                        unsigned long long* __restrict__ keys,
                                                        RasCounters* __restrict__ ctr) {
    // last kernel of the frame: leave the per-frame counters clear for the next one (no memset in steady state)
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ctr->nBig = ctr->bigRows = ctr->bigSamples = ctr->err = 0u;
    const int y = a.y0 + blockIdx.y;
    const int xbase = blockIdx.x * (kShadeThreads * kShadePixels) + threadIdx.x;
    unsigned long long key[kShadePixels];
#pragma unroll
    for (int p = 0; p < kShadePixels; ++p) {
        const int x = xbase + kShadeThreads * p;
        key[p] = (x < a.W) ? keys[(size_t)(y - a.y0) * (size_t)a.W + (size_t)x] : 0ull;
    }
    ShadeIn in[kShadePixels];
#pragma unroll
    for (int p = 0; p < kShadePixels; ++p) {
        const int x = xbase + kShadeThreads * p;
        if (key[p] != 0ull) {
            keys[(size_t)(y - a.y0) * (size_t)a.W + (size_t)x] = 0ull;  // depthBuffer = 0 (:188) for the next frame
            in[p] = shade_fetch(a, key[p], y, bigTs, triInfo, rowRec, rows);
        }
    }
#pragma unroll
    for (int p = 0; p < kShadePixels; ++p) {
        const int x = xbase + kShadeThreads * p;
        if (x >= a.W) continue;
        float depth = 0.f, focal = 0.f;
        V3 colour = mk3(0.f, 0.f, 0.f);
        int winner = -1;
        if (key[p] != 0ull) {
            winner = (int)key_triangle(key[p]);
            depth = __uint_as_float((unsigned)(key[p] >> 32));
            shade_pixel(a.fr, in[p], x, depth, focal, colour);
        }
        const size_t idx = (size_t)y * (size_t)a.W + (size_t)x;
        if (a.depth) a.depth[idx] = depth;
        if (a.colours) {
            a.colours[3 * idx] = colour.x;
            a.colours[3 * idx + 1] = colour.y;
            a.colours[3 * idx + 2] = colour.z;
        }
        if (a.focal) a.focal[idx] = focal;
        if (a.winner) a.winner[idx] = winner;
        // CalculateDOF without depth of field + PutPixelSDL (:516-526), fused
        if (a.surface) a.surface[idx] = inside_border(x, y, a.W, a.H) ? pack_xrgb(colour.x, colour.y, colour.z) : 0u;
    }
}

// ---- host side ------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Returns cudaErrorInvalidValue when a triangle exceeds the row/coordinate limits (-> B2R_E_CAPACITY).
// Scenes of few triangles (T * band height row slots within this budget: ~100 MB of row records, ~200 MB of edge
// samples) take the large-triangle path with fixed-capacity slots: nothing is read back, the frame is a plain
// sequence of launches.  Larger scenes size the buffers from counters read back after the first kernel.
constexpr size_t kBandSlotLimit = 2u << 20;

// threads per CTA of ras_edges: one warp per CTA while that leaves SMs without one -- each thread issues one scattered
// store per step of its chain, and the few warps of a 30-triangle scene would otherwise queue on four SMs' LSUs
static inline int edges_block(int threads, int smCount) { return threads <= smCount * 32 * 4 ? 32 : 128; }

// rows per warp of ras_rows: enough warps to fill the machine (about 48 per SM) before a warp takes more than one row
static inline int rows_per_warp(size_t nRows, int smCount) {
    const size_t want = nRows / ((size_t)smCount * 48);
    int r = 1;
    while (r < 32 && (size_t)(2 * r) <= want) r *= 2;
    return r;
}

}  // namespace sl

using namespace sl;

cudaError_t launch_ras_draw_sortlast(Ctx* c, const RasLaunch& a0, cudaStream_t s) {
    RasLaunch a = a0;
    const int T = a.T;
    const int bandH = a.y1 - a.y0;
    cudaError_t e;
    // scratch: [counters 64 B][bigCounts uint2 x T][excl uint2 x T][blockSums][totals]; triInfo int2 x T separately
    const int nbMax = (T + kScanBlock - 1) / kScanBlock + 1;
    const size_t offCtr = 0, offCounts = 256, offExcl = align_up(offCounts + sizeof(uint2) * (size_t)T, 256),
                 offSums = align_up(offExcl + sizeof(uint2) * (size_t)T, 256),
                 offTotals = align_up(offSums + sizeof(uint2) * (size_t)nbMax, 256), scratchBytes = offTotals + 256;
    const void* scratchBefore = c->rasSLScratch.p;
    if ((e = c->rasSLScratch.reserve(scratchBytes)) != cudaSuccess) return e;
    if (c->rasSLScratch.p != scratchBefore &&  // fresh memory: the sticky error flag starts clear
        (e = cudaMemsetAsync(c->rasSLScratch.p, 0, 256, s)) != cudaSuccess)
        return e;
    if ((e = c->rasTri.reserve(sizeof(TriSetup) * (size_t)(T + 1) + sizeof(int2) * (size_t)(T + 1) + 1024)) != cudaSuccess) return e;
    if ((e = c->rasSmall.reserve(sizeof(SmallRow) * (size_t)kSmallRows * (size_t)(T + 1))) != cudaSuccess) return e;
    if ((e = c->rasKeys.reserve(sizeof(unsigned long long) * (size_t)bandH * a.W + 256)) != cudaSuccess) return e;
    unsigned char* sc = c->rasSLScratch.as<unsigned char>();
    RasCounters* ctr = reinterpret_cast<RasCounters*>(sc + offCtr);
    uint2* counts = reinterpret_cast<uint2*>(sc + offCounts);
    uint2* excl = reinterpret_cast<uint2*>(sc + offExcl);
    uint2* sums = reinterpret_cast<uint2*>(sc + offSums);
    uint2* totals = reinterpret_cast<uint2*>(sc + offTotals);
    int2* triInfo = c->rasTri.as<int2>();
    SmallRow* rowRec = c->rasSmall.as<SmallRow>();
    TriSetup* ts = reinterpret_cast<TriSetup*>(c->rasTri.as<unsigned char>() + align_up(sizeof(int2) * (size_t)(T + 1), 256));
    unsigned long long* keys = c->rasKeys.as<unsigned long long>();

    const bool bandSlots = T > 0 && (size_t)T * (size_t)bandH <= kBandSlotLimit && c->optRasVariant != 1;
    a.bandSlots = bandSlots ? 1 : 0;
    // the sticky flag (second half of the struct) survives until the host has reported it
    // ... and so do the per-frame counters, cleared by the previous frame's shade kernel; clear them here only
    // after a draw that did not get that far (or on a fresh buffer, above)
    if (c->rasSLCtrDirty && (e = cudaMemsetAsync(ctr, 0, offsetof(RasCounters, sticky), s)) != cudaSuccess) return e;
    c->rasSLCtrDirty = true;
    // depthBuffer = 0 (:188): the shade pass of the previous frame leaves the key buffer cleared; clear it here
    // only when the buffer is new or was last used for a different band size
    const size_t keyBytes = sizeof(unsigned long long) * (size_t)bandH * a.W;
    if (c->rasKeysClean != keyBytes || c->rasKeysCleanPtr != (void*)keys) {
        if ((e = cudaMemsetAsync(keys, 0, keyBytes, s)) != cudaSuccess) return e;
    }
    c->rasKeysClean = 0;
    RasCounters host{};
    const RowRec* rowsPtr = nullptr;
    if (T > 0) {
        const size_t smem = sizeof(int) * 8 * kWindowRows * kSmallThreads;
        if (!c->rasSmallAttr) {  // once per context (= per device)
            if ((e = cudaFuncSetAttribute(ras_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
            c->rasSmallAttr = true;
        }
        if (bandSlots) {  // worst case: every triangle is large
            const size_t nSlots = (size_t)T * (size_t)bandH;
            if ((e = c->rasRows.reserve(sizeof(RowRec) * nSlots + sizeof(EdgeSample) * 3 * nSlots + 1024)) != cudaSuccess) return e;
        }
        ras_small_kernel<<<(T + kSmallThreads - 1) / kSmallThreads, kSmallThreads, smem, s>>>(a, keys, ts, counts, triInfo, rowRec, ctr);
        c->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (bandSlots) {
        const size_t nSlots = (size_t)T * (size_t)bandH;
        unsigned char* rb = c->rasRows.as<unsigned char>();
        RowRec* rows = reinterpret_cast<RowRec*>(rb);
        EdgeSample* samples = reinterpret_cast<EdgeSample*>(rb + align_up(sizeof(RowRec) * nSlots, 256));
        rowsPtr = rows;
        const int eb = edges_block(15 * T, c->smCount);
        ras_edges_kernel<true><<<(15 * T + eb - 1) / eb, eb, 0, s>>>(ts, T, ctr, samples, nullptr, a.y0, a.y1);
        const int rpw = rows_per_warp(nSlots, c->smCount);
        ras_rows_kernel<true><<<(unsigned)((nSlots / rpw + 1 + 7) / 8), 256, 0, s>>>(ts, ctr, samples, nullptr, 0u, rows, keys, a.W, a.y0,
                                                                                    a.y1, a.stats, rpw);
        c->launches += 2;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        c->rasErrPending = true;  // checked by the caller's next synchronising call (ras_take_error)
        c->rasErrCtr = ctr;
    } else if (T > 0) {
        // How many large triangles / rows / edge samples: 32 bytes back to size the big path -- once per (scene,
        // culling flags, frame params, band).  The counts are a pure function of those, so later frames of the same
        // state reuse them and the draw only enqueues.
        Ctx::RasSizes& z = c->rasSizesSL;
        if (!(z.valid && z.gen == c->rasGen && z.y0 == a.y0 && z.y1 == a.y1)) {
            if ((e = cudaMemcpyAsync(c->pinned, ctr, sizeof(RasCounters), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
            host = *reinterpret_cast<RasCounters*>(c->pinned);
            if (host.err) {
                cudaMemsetAsync(&ctr->sticky, 0, sizeof(unsigned), s);  // reported right here
                return cudaErrorInvalidValue;
            }
            z.valid = true;
            z.gen = c->rasGen;
            z.y0 = a.y0;
            z.y1 = a.y1;
            z.nBig = host.nBig;
            z.bigRows = host.bigRows;
            z.bigSamples = host.bigSamples;
        } else {
            host.nBig = z.nBig;
            host.bigRows = z.bigRows;
            host.bigSamples = z.bigSamples;
            c->rasErrPending = true;  // cannot be set for a state that drew cleanly before; kept for symmetry
            c->rasErrCtr = ctr;
        }
    }
    if (host.nBig > 0) {
        const int nBig = (int)host.nBig;
        const unsigned nRows = host.bigRows, nSamples = host.bigSamples;
        const int nb = (nBig + kScanBlock - 1) / kScanBlock;
        if ((e = c->rasRows.reserve(sizeof(RowRec) * (size_t)nRows + sizeof(unsigned) * (size_t)nRows +
                                    sizeof(EdgeSample) * (size_t)nSamples + 1024)) != cudaSuccess)
            return e;
        unsigned char* rb = c->rasRows.as<unsigned char>();
        RowRec* rows = reinterpret_cast<RowRec*>(rb);
        unsigned* owner = reinterpret_cast<unsigned*>(rb + align_up(sizeof(RowRec) * (size_t)nRows, 256));
        EdgeSample* samples = reinterpret_cast<EdgeSample*>(reinterpret_cast<unsigned char*>(owner) +
                                                            align_up(sizeof(unsigned) * (size_t)nRows, 256));
        rowsPtr = rows;
        scan_blocks_kernel<<<nb, kScanBlock, 0, s>>>(counts, excl, sums, nBig);
        scan_sums_kernel<<<1, kScanBlock, 0, s>>>(sums, nb, totals);
        scan_apply_kernel<<<nb, kScanBlock, 0, s>>>(excl, sums, ts, triInfo, nBig);
        const int eb = edges_block(15 * nBig, c->smCount);
        ras_edges_kernel<false><<<(15 * nBig + eb - 1) / eb, eb, 0, s>>>(ts, nBig, ctr, samples, owner, a.y0, a.y1);
        const int rpw = rows_per_warp(nRows, c->smCount);
        ras_rows_kernel<false><<<(nRows / rpw + 1 + 7) / 8, 256, 0, s>>>(ts, ctr, samples, owner, nRows, rows, keys, a.W, a.y0, a.y1, a.stats, rpw);
        c->launches += 5;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    {
        dim3 grid((a.W + kShadeThreads * kShadePixels - 1) / (kShadeThreads * kShadePixels), bandH);
        ras_shade_kernel<<<grid, kShadeThreads, 0, s>>>(a, ts, triInfo, rowRec, rowsPtr, keys, ctr);
        c->launches++;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) {
        c->rasSLCtrDirty = false;
        c->rasKeysClean = keyBytes;
        c->rasKeysCleanPtr = keys;
    }
    return e;
}

}  // namespace b2r
