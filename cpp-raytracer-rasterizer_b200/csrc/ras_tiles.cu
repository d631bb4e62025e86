// ras_tiles.cu -- the rasteriser hot path on sm_100a, screen-tile form (B2R_OPT_RAS_VARIANT = 2), plus what both
// pipelines share on the host side (culling, the 64-byte scene copy, the deferred error flag, the dispatcher).
//
// Replaces Draw() -> DrawPolygon() -> VertexShader / ComputePolygonRows /
// Interpolate / DrawRows / DrawLineSDL / Bresenham / PixelShader of
// rasteriser/Source/rasteriser.cpp:461-482, :532-546, :549-589, :592-672,
// :674-768, and the culling block of Update() (:385-447).
//
// The reference draws triangles serially in index order with a strict
// `zinv > depthBuffer` test against a buffer cleared to 0 (:606, :188).  The
// final image therefore only depends on, per pixel, the fragment with the
// largest zinv, the lowest triangle index among exact ties, and only
// fragments with zinv > 0.  That rule is order-free, so the frame is built
// sort-middle: triangles are binned to 32x32 screen tiles and one CTA per tile
// resolves depth and ownership in shared memory, then runs PixelShader once per
// pixel for the winner (PixelShader's writes are simply overwritten by later
// winners in the reference, so shading only the final one gives identical arrays).
//
// All arithmetic that decides coverage, depth or colour is in reference order,
// non-fused.  Interpolate's serial float accumulation along each edge (:632-635)
// decides both coverage (via int(current.x)) and the zinv values compared in the
// depth test, so it is replayed step by step, never re-associated.
//
// Pipeline:
//   ras_setup     1 thread / triangle: VertexShader x3, row count, conservative tile rectangle, per-tile
//                 reference counts.  Triangles of more than 20 rows (or wider than 128 tiles) go to the
//                 large-triangle list, whose edge walks are done once by ras_edges into HBM.
//                 The last CTA to finish turns the counts into list offsets (exclusive scan).
//   ras_bin       1 thread / triangle: writes the triangle into the list of every tile of its rectangle; lanes
//                 of a warp that hit the same tile share one atomic (match + prefix popcount).
//   ras_edges     large triangles only: 1 thread / (triangle, edge, chain), stores every edge sample.
//   ras_tile      1 CTA / tile, everything in shared memory:
//                   A1  1 thread / listed triangle: VertexShader x3 again (cheaper than a 60-byte round trip
//                       through HBM), rows inside the tile, block prefix sum -> row items
//                   W   1 thread / (triangle, edge): Interpolate's serial walk, samples of the tile's rows kept
//                   A3  1 thread / row item: ComputePolygonRows' left/right resolve, DrawLineSDL's span clipped
//                       to the tile, atomicMax of the zinv bits on the 32-bit depth tile
//                   B1  same thread: fragments that hold the final depth take atomicMin(owner, triangle index)
//                   B2  same thread: the owning fragment leaves its interpolated pos3d.xy for the pixel
//                   C   1 thread / pixel: PixelShader of the winner, coalesced stores of depth / colour / ...
//                 The 64-bit (depth, index) order is thus resolved with native 32-bit shared-memory atomics;
//                 no key, span or row record ever reaches HBM.
#include <limits.h>

#include <algorithm>

#include "b2r_internal.h"
#include "exact.cuh"
#include "pixel_pack.cuh"
#include "ras_common.cuh"
#include "ras_device.cuh"

namespace b2r {

constexpr int kTileShift = 5;
constexpr int kTileW = 1 << kTileShift, kTileH = 1 << kTileShift, kTilePix = kTileW * kTileH;
constexpr int kTileThreads = 256;
constexpr int kSmallRows = 20;          // triangles up to this many polygon rows are walked inside ras_tile
constexpr int kMaxSmallTilesX = 128;    // ... unless they are wider than this many tiles
constexpr unsigned kBigRef = 0x80000000u;  // tile-list entry / triWord: large-triangle slot in the low bits
constexpr unsigned kNoRect = 0x7FFFFFFFu;  // triWord: nothing to draw (culled, off screen, outside the band)
constexpr unsigned kSmallTri = 0xFFFFFFFFu;
constexpr int kSlice = 512;             // tile-list entries per job (CTA) of ras_tile; longer lists are split and merged

constexpr int kMaxRowsPerTriangle = 1 << 22;
constexpr int kCoordLimit = 1 << 24;

// Large triangle: projected vertices (VertexShader output), where its edge samples live, its tile rectangle.
// pos3d.z is omitted everywhere: it is pos.z/pos.z == 1.0f exactly for every triangle that passes the coordinate
// limits (a NaN there makes x,y INT_MIN -> B2R_E_CAPACITY), Interpolate's z step is then (1-1)/n == 0 and
// Bresenham's is 0/pixels == 0, so that chain is 1.0f for every fragment.
struct TriSetup {  // 24 words
    int vx[3], vy[3];
    float vz[3], vpx[3], vpy[3];
    int minY, rows;
    unsigned sampleBase;
    int tri;  // index in the caller's triangle array (draw order)
    int tx0, tx1, ty0, ty1;
    int pad;
};

struct EdgeSample {  // one Interpolate() result entry (:628-631) without y (implied) and pos3d.z (== 1)
    int x;
    float zinv, px, py;
};


// How far int(current.x) of an edge walk can lie outside the vertices' x range: the chain a.x + step + step + ...
// drifts from the exact line by at most one ulp of the largest magnitude per step (half from the addition, half
// from the rounded step), plus the truncation.
__device__ __forceinline__ int coord_margin(int steps, int maxAbs) {
    return 2 + (int)((float)steps * (float)maxAbs * 2.4e-7f);  // 2^-22 = 2.38e-7
}

// The three vertices: 9 floats at the head of the 64-byte record (b2r_set_triangles repacks 60-byte scenes once), three
// 128-bit loads.
__device__ __forceinline__ void load_vertices(const RasLaunch& a, int i, float* t /* 12 */) {
    const float4* q = reinterpret_cast<const float4*>(a.raw + (size_t)i * 64);
    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
    t[0] = q0.x; t[1] = q0.y; t[2] = q0.z; t[3] = q0.w; t[4] = q1.x; t[5] = q1.y; t[6] = q1.z; t[7] = q1.w;
    t[8] = q2.x; t[9] = q2.y; t[10] = q2.z; t[11] = q2.w;
}

// ---- stage 1: set-up, classification, per-tile counts; stage 2 (tile offsets) in its last block ----
// Lanes of a warp that touch the same tile share one atomic: __match_any_sync groups them and the group's first lane
// adds popc(group).
__device__ __forceinline__ void count_tile(unsigned* __restrict__ tileCount, int tile, int lane) {
    const unsigned peers = __match_any_sync(__activemask(), tile);
    if (lane == __ffs(peers) - 1) atomicAdd(&tileCount[tile], (unsigned)__popc(peers));
}

// One CTA of 256 threads, each owning a contiguous run of tiles: exclusive scan of the per-tile counts (list offsets) and
// the job table of ras_tile -- a tile whose list is longer than kSlice entries is drawn by several CTAs (jobs), each
// taking kSlice entries, which meet in a partial-result buffer (slot mslot + slice).  The counts are zeroed on the way
// (ras_bin uses them as cursors).  tileOffset[nTiles] = total entries, tileOffset[nTiles + 1] = total jobs.
// FROM_OFFSETS: only the job table and the partial slots are (re)built, from offsets that are already there.
template <bool FROM_OFFSETS>
__device__ void tile_scan_block(unsigned* __restrict__ tileCount, unsigned* __restrict__ tileOffset, unsigned* __restrict__ tileMslot,
                                uint2* __restrict__ jobTile, unsigned jobCap, int nTiles, RasCounters* __restrict__ ctr) {
    constexpr int kChunk = 256 * 16;  // tiles staged in shared memory at a time (coalesced loads, all in flight)
    __shared__ unsigned cnt[kChunk];
    __shared__ unsigned warpSums[3][8];
    __shared__ unsigned carry[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 3) carry[tid] = 0u;
    for (int c0 = 0; c0 < nTiles; c0 += kChunk) {
        const int nc = min(kChunk, nTiles - c0);
        __syncthreads();
        for (int t = tid; t < nc; t += 256)
            cnt[t] = FROM_OFFSETS ? __ldcg(&tileOffset[c0 + t + 1]) - __ldcg(&tileOffset[c0 + t]) : __ldcg(&tileCount[c0 + t]);
        __syncthreads();
        const int per = (nc + 255) / 256, t0 = min(tid * per, nc), t1 = min(t0 + per, nc);
        unsigned sum[3] = {0u, 0u, 0u};  // entries, jobs, partial slots
        for (int t = t0; t < t1; ++t) {
            const unsigned v = cnt[t], k = max(1u, (v + kSlice - 1) / kSlice);
            sum[0] += v;
            sum[1] += k;
            sum[2] += (k > 1u) ? k : 0u;
        }
        unsigned run[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            unsigned inc = sum[c];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned x = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += x;
            }
            if (lane == 31) warpSums[c][warp] = inc;
            run[c] = inc - sum[c];
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            run[c] += carry[c];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < warp) run[c] += warpSums[c][k];
        }
        for (int t = t0; t < t1; ++t) {
            const unsigned v = cnt[t], k = max(1u, (v + kSlice - 1) / kSlice);
            if (!FROM_OFFSETS) {
                tileOffset[c0 + t] = run[0];
                tileCount[c0 + t] = 0u;
            }
            tileMslot[c0 + t] = run[2];
            for (unsigned sl = 0; sl < k; ++sl)
                if (run[1] + sl < jobCap) jobTile[run[1] + sl] = make_uint2((unsigned)(c0 + t), sl);  // a short table is rebuilt by ras_jobs
            run[0] += v;
            run[1] += k;
            run[2] += (k > 1u) ? k : 0u;
        }
        __syncthreads();
        if (tid == 255) {  // owns the chunk's last tiles: its running sums are the chunk totals
            carry[0] = run[0];
            carry[1] = run[1];
            carry[2] = run[2];
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (!FROM_OFFSETS) {
            tileOffset[nTiles] = carry[0];
            ctr->totalRefs = carry[0];
            ctr->totalJobs = carry[1];
        }
        tileOffset[nTiles + 1] = carry[1];
    }
}

// Rebuilds the job table after the host has sized it (first frame of a new scene / camera / band only).
__global__ void __launch_bounds__(256) ras_jobs_kernel(unsigned* tileOffset, unsigned* tileMslot, uint2* jobTile, unsigned jobCap,
                                                       int nTiles) {
    tile_scan_block<true>(nullptr, tileOffset, tileMslot, jobTile, jobCap, nTiles, nullptr);
}

__global__ void __launch_bounds__(256) ras_setup_kernel(RasLaunch a, int tilesX, int nTiles, unsigned* __restrict__ triWord,
                                                        TriSetup* __restrict__ bigTs, uint2* __restrict__ bigCounts,
                                                        unsigned* __restrict__ tileCount, unsigned* __restrict__ tileOffset,
                                                        unsigned* __restrict__ tileMslot, uint2* __restrict__ jobTile,
                                                        unsigned jobCap, RasCounters* __restrict__ ctr) {
    const int lane = threadIdx.x & 31;
    unsigned long long nRows = 0, nDrawn = 0;
    // a CTA walks blocks of 256 triangles (grid = a few CTAs per SM): one fence + ticket per CTA at the end
    for (int blk = blockIdx.x; blk * 256 < a.T; blk += gridDim.x) {
    const int i = blk * 256 + threadIdx.x;
    unsigned word = kNoRect;
    int tx0 = 0, tx1 = -1, ty0 = 0, ty1 = -1;
    bool big = false;
    float t[12];
    unsigned char isCulled = 1;
    if (i < a.T) {  // both loads in flight together: the flag does not gate the vertex fetch
        isCulled = a.culled ? a.culled[i] : (unsigned char)0;
        load_vertices(a, i, t);
    }
    if (!isCulled) {  // :470
        RPixel v[3];
        int maxY = INT_MIN, minY = INT_MAX, maxX = INT_MIN, minX = INT_MAX;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = vertex_shader<true>(a.fr, mk3(t[3 * k], t[3 * k + 1], t[3 * k + 2]));  // :760-761
            maxY = max(maxY, v[k].y);
            minY = min(minY, v[k].y);
            maxX = max(maxX, v[k].x);
            minX = min(minX, v[k].x);
            bad = bad || v[k].x <= -kCoordLimit || v[k].x >= kCoordLimit || v[k].y <= -kCoordLimit || v[k].y >= kCoordLimit;
        }
        const int rows = maxY - minY + 1;  // :682
        if (bad || rows > kMaxRowsPerTriangle) {
            atomicExch(&ctr->err, 1u);  // the reference would try to allocate/walk an absurd row count
            atomicExch(&ctr->sticky, 1u);
        } else {
            nDrawn += 1;
            nRows += (unsigned long long)rows;
            // conservative rectangle of the fragments (x = lx+1 .. rx of every row), clipped to the screen and the band
            const int m = coord_margin(rows, max(abs(minX), abs(maxX)));
            const int xa = max(minX - m, 0), xb = min(maxX + m, a.W - 1);
            const int ya = max(minY, a.y0), yb = min(maxY, a.y1 - 1);  // DrawRows :743 + the band
            if (xa <= xb && ya <= yb) {
                tx0 = xa >> kTileShift;
                tx1 = xb >> kTileShift;
                ty0 = (ya - a.y0) >> kTileShift;
                ty1 = (yb - a.y0) >> kTileShift;
                big = rows > kSmallRows || tx1 - tx0 >= kMaxSmallTilesX;
                if (big) {
                    const unsigned samples = (unsigned)(abs(v[0].y - v[1].y) + abs(v[1].y - v[2].y) + abs(v[2].y - v[0].y) + 3);
                    const unsigned slot = atomicAdd(&ctr->nBig, 1u);
                    atomicAdd(&ctr->bigRows, (unsigned)rows);
                    atomicAdd(&ctr->bigSamples, samples);
                    TriSetup s;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        s.vx[k] = v[k].x;
                        s.vy[k] = v[k].y;
                        s.vz[k] = v[k].zinv;
                        s.vpx[k] = v[k].p.x;
                        s.vpy[k] = v[k].p.y;
                    }
                    s.minY = minY;
                    s.rows = rows;
                    // fixed-capacity slots (small scenes): only rows of the band are kept, bandH samples per edge,
                    // addressed by y - y0 (no scan, no readback); otherwise the scan fills sampleBase in
                    s.sampleBase = a.bandSlots ? slot * 3u * (unsigned)(a.y1 - a.y0) : 0u;
                    s.tri = i;
                    s.tx0 = tx0; s.tx1 = tx1; s.ty0 = ty0; s.ty1 = ty1;
                    s.pad = 0;
                    bigTs[slot] = s;
                    if (!a.bandSlots) bigCounts[slot] = make_uint2((unsigned)rows, samples);
                    word = kBigRef | slot;
                } else {
                    word = (unsigned)tx0 | ((unsigned)ty0 << 10) | ((unsigned)(tx1 - tx0) << 20) | ((unsigned)(ty1 - ty0) << 27);
                    for (int ty = ty0; ty <= ty1; ++ty)
                        for (int tx = tx0; tx <= tx1; ++tx) count_tile(tileCount, ty * tilesX + tx, lane);
                }
            }
        }
    }
    if (i < a.T) triWord[i] = word;
    // large triangles: the warp counts the tiles of each rectangle together
    unsigned bigMask = __ballot_sync(0xffffffffu, big);
    while (bigMask) {
        const int src = __ffs(bigMask) - 1;
        bigMask &= bigMask - 1;
        const int bx0 = __shfl_sync(0xffffffffu, tx0, src), bx1 = __shfl_sync(0xffffffffu, tx1, src);
        const int by0 = __shfl_sync(0xffffffffu, ty0, src), by1 = __shfl_sync(0xffffffffu, ty1, src);
        const int w = bx1 - bx0 + 1, n = w * (by1 - by0 + 1);
        for (int k = lane; k < n; k += 32) atomicAdd(&tileCount[(by0 + k / w) * tilesX + bx0 + k % w], 1u);
    }
    }
    if (a.stats) {
        for (int off = 16; off > 0; off >>= 1) {
            nRows += __shfl_xor_sync(0xffffffffu, nRows, off);
            nDrawn += __shfl_xor_sync(0xffffffffu, nDrawn, off);
        }
        if (lane == 0) {
            if (nRows) atomicAdd(a.stats + B2R_STAT_RAS_ROWS, nRows);
            if (nDrawn) atomicAdd(a.stats + B2R_STAT_RAS_TRIANGLES, nDrawn);
        }
    }
    // the last CTA to get here turns the tile counts into list offsets (no separate launch)
    __shared__ bool isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // cumulative: orders this CTA's counts (made visible to this thread by the barrier) before the ticket
        isLast = atomicAdd(&ctr->blocksDone, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (isLast) {
        __threadfence();
        tile_scan_block<false>(tileCount, tileOffset, tileMslot, jobTile, jobCap, nTiles, ctr);
    }
}

// ---- stage 3: binning -------------------------------------------------------------------------------
// Lanes of a warp that append to the same tile list share one atomicAdd: __match_any_sync groups them, the group's
// first lane reserves popc(group) entries and each lane's position is the prefix popcount of the group below it.
__device__ __forceinline__ void bin_append(unsigned* __restrict__ cursor, const unsigned* __restrict__ tileOffset,
                                           unsigned* __restrict__ refs, int tile, unsigned entry, int lane) {
    const unsigned peers = __match_any_sync(__activemask(), tile);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&cursor[tile], (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    B2R_BOUND(base + (unsigned)__popc(peers & ((1u << lane) - 1u)), tileOffset[tile + 1] - tileOffset[tile]);
    refs[tileOffset[tile] + base + (unsigned)__popc(peers & ((1u << lane) - 1u))] = entry;
}

__global__ void __launch_bounds__(256) ras_bin_kernel(const unsigned* __restrict__ triWord, int T, int tilesX,
                                                      const TriSetup* __restrict__ bigTs, unsigned* __restrict__ cursor,
                                                      const unsigned* __restrict__ tileOffset, unsigned* __restrict__ refs) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned w = (i < T) ? triWord[i] : kNoRect;
    const bool big = (w & kBigRef) != 0u;
    if (!big && w != kNoRect) {
        const int tx0 = (int)(w & 1023u), ty0 = (int)((w >> 10) & 1023u);
        const int nx = (int)((w >> 20) & 127u) + 1, ny = (int)((w >> 27) & 15u) + 1;
        for (int dy = 0; dy < ny; ++dy)
            for (int dx = 0; dx < nx; ++dx) bin_append(cursor, tileOffset, refs, (ty0 + dy) * tilesX + tx0 + dx, (unsigned)i, lane);
    }
    unsigned bigMask = __ballot_sync(0xffffffffu, big);
    while (bigMask) {
        const int src = __ffs(bigMask) - 1;
        bigMask &= bigMask - 1;
        const unsigned entry = __shfl_sync(0xffffffffu, w, src);
        const TriSetup* s = bigTs + (entry & ~kBigRef);
        const int bx0 = s->tx0, by0 = s->ty0, bw = s->tx1 - bx0 + 1, n = bw * (s->ty1 - by0 + 1);
        for (int k = lane; k < n; k += 32) {  // distinct tiles per lane: plain atomics
            const int tile = (by0 + k / bw) * tilesX + bx0 + k % bw;
            refs[tileOffset[tile] + atomicAdd(&cursor[tile], 1u)] = entry;
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
                                  TriSetup* __restrict__ ts, int n) {
    int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i >= n) return;
    uint2 o = add2(ex[i], blockSums[blockIdx.x]);
    ts[i].sampleBase = o.y;
}

// ---- large triangles: Interpolate (:615-637), one thread per (triangle, edge, chain) -----------------
// The four accumulation chains of an edge (x, zinv, pos3d.x, pos3d.y) are independent of each other, so each runs its
// own serial loop -- 12 threads per triangle -- and writes its field of every sample.  A 2160-row edge is latency-bound:
// one dependent FADD per step and thread.
// BAND: fixed-capacity slots (see ras_setup): the number of listed triangles is read from the device counter, the walk
// still runs over every step (the accumulation is serial) but only samples on rows of the band are stored, at y - y0.
template <bool BAND>
__global__ void ras_edges_kernel(const TriSetup* __restrict__ ts, int T /* listed large triangles (upper bound if BAND) */,
                                 const RasCounters* __restrict__ ctr, EdgeSample* __restrict__ samples, int y0, int y1) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid / 12, rem = gid - 12 * i, e = rem >> 2, ch = rem & 3;
    if (i >= (BAND ? (int)ctr->nBig : T)) return;
    const TriSetup s = ts[i];
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
        step = xdiv_step(xsub(s.vz[j], cur), div);
    } else if (ch == 2) {
        cur = s.vpx[e];
        step = xdiv_step(xsub(s.vpx[j], cur), div);
    } else {
        cur = s.vpy[e];
        step = xdiv_step(xsub(s.vpy[j], cur), div);
    }
    float* out = reinterpret_cast<float*>(samples + off) + ch;  // field ch of sample 0; samples are 4 words apart
    int kLo = 0, kHi = n - 1, stride = 4;
    if (BAND) {
        // step k lies on row vy[e] + sgn*k: the steps before the band only accumulate, the steps inside it are
        // stored, the steps after it are not needed by anyone
        const int ya = s.vy[e], sgn = (s.vy[j] > ya) - (s.vy[j] < ya);
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
        out += 4 * (ya + sgn * kLo - y0);
        stride = 4 * sgn;
    }
    for (int k = 0; k < kLo; ++k) cur = xadd(cur, step);  // :626-636 -- serial float accumulation, order matters
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
}

// ---- the tile kernel ----------------------------------------------------------------------------------
struct TileArgs {
    const unsigned* tileOffset;  // nTiles + 1
    const unsigned* refs;        // tile lists: triangle index, or kBigRef | large-triangle slot
    unsigned* tileCount;         // re-armed (zeroed) for the next frame
    const TriSetup* bigTs;
    const EdgeSample* samples;   // edge samples of the large triangles
    RasCounters* ctr;
    const unsigned* tileMslot;   // first partial-result slot of a tile drawn by several jobs
    const uint2* jobTile;        // job -> (tile, slice of its list)
    unsigned* tileDone;          // jobs of the tile that have delivered their partial result
    uint4* partials;             // per slot, per pixel: (zinv bits, triangle, pos3d.x, pos3d.y)
    int tilesX, nTiles;
    int nEpochs;                 // ceil(T / 2^20), see the depth key below
};

// Depth key of a pixel, one 64-bit word in shared memory updated with atomicMax:
//   high word  zinv bits (zinv > 0, so unsigned order == float order): larger zinv wins (:606)
//   low word   (0xFFFFF - (triangle & 0xFFFFF)) << 12 | row item: among equal zinv the LOWER triangle index wins -- the
//              reference's strict `>` in draw order -- and the low 12 bits name the row item (this round's thread)
//              whose span produced the fragment, so that the pixel can fetch its attributes right after the round.
// The triangle field holds 20 bits.  Scenes of more than 2^20 triangles are drawn in epochs of 2^20 consecutive
// indices, lowest first (every epoch walks the tile's list and skips the other epochs' triangles); between epochs
// the low word of every covered pixel is raised to 0xFFFFFFFF, so a later epoch (higher indices) can only take a
// pixel with a strictly larger zinv.  Within a tile a triangle appears in exactly one round of one epoch and has at
// most one fragment per pixel, so a pixel's low word changes whenever its owner does.
constexpr int kTriBits = 20, kSlotBits = 12;
constexpr unsigned kTriMask = (1u << kTriBits) - 1u;

struct RowRec {  // what PixelShader needs from the winning row: Bresenham's pos3d.xy start and step (:649, :668)
    int lx;
    float lpx, psx, lpy, psy;
    unsigned tri;
};

constexpr int kPixPerThread = kTilePix / kTileThreads;  // 4: pixel p = tid + 256 i, a warp covers one tile row
constexpr int kRoundRows = 384;                         // row items per round (up to two per thread in A3)

struct __align__(16) TileShared {
    float4 samples[3 * kRoundRows];    // W -> A3: (int x, zinv, pos3d.x, pos3d.y) of row item r, edge e at [3r + e]
    unsigned long long key[kTilePix];  // depth keys, 0 = nothing drawn (depthBuffer = 0, :188)
    RowRec rec[kRoundRows];            // A3 -> G: the round's row items
    // the triangles of the current batch, one per thread (VertexShader output)
    int vy[3][kTileThreads], vx[3][kTileThreads];
    float vz[3][kTileThreads], vpx[3][kTileThreads], vpy[3][kTileThreads];
    unsigned tri[kTileThreads];    // index in the caller's array
    unsigned sbase[kTileThreads];  // large triangle: where its edge samples start; kSmallTri otherwise
    int yA[kTileThreads];          // first row inside the tile
    unsigned rowBase[kTileThreads + 4];  // exclusive prefix sum of the rows inside the tile
    unsigned short rowTri[kRoundRows];   // row item of the current round -> triangle slot of the batch
    unsigned warpSums[8];
};
static_assert(sizeof(TileShared) <= 55 * 1024 + 768, "four CTAs per SM");
static_assert(kRoundRows <= (1 << kSlotBits), "row item must fit the key's slot field");

__global__ void __launch_bounds__(kTileThreads, 4) ras_tile_kernel(RasLaunch a, TileArgs g) {
    extern __shared__ __align__(16) unsigned char tileSmem[];
    TileShared& S = *reinterpret_cast<TileShared*>(tileSmem);
    __shared__ bool lastJob;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (blockIdx.x >= g.tileOffset[g.nTiles + 1]) return;  // the grid is sized for the worst case
    const uint2 job = g.jobTile[blockIdx.x];
    const int tile = (int)job.x, slice = (int)job.y;
    const int tileX = tile % g.tilesX, tileY = tile / g.tilesX;
    const int X0 = tileX << kTileShift, X1 = min(X0 + kTileW, a.W);
    const int Y0 = a.y0 + (tileY << kTileShift), Y1 = min(Y0 + kTileH, a.y1);
    const unsigned listLen = g.tileOffset[tile + 1] - g.tileOffset[tile];
    const unsigned beg = g.tileOffset[tile] + (unsigned)slice * kSlice, nRefs = min((unsigned)kSlice, listLen - (unsigned)slice * kSlice);
    const unsigned nJobs = max(1u, (listLen + kSlice - 1) / kSlice);
    // last kernel of the frame: leave the per-frame counters and the tile counts clear for the next one
    if (tid == 0 && slice == 0) {
        g.tileCount[tile] = 0u;
        if (tile == 0) g.ctr->nBig = g.ctr->bigRows = g.ctr->bigSamples = g.ctr->err = g.ctr->totalRefs = g.ctr->blocksDone = g.ctr->totalJobs = 0u;
    }
    unsigned long long nTests = 0;
    // the winner of this thread's pixels so far: low key word seen last, triangle, interpolated pos3d.xy
    unsigned seenLo[kPixPerThread], wtri[kPixPerThread];
    float posx[kPixPerThread], posy[kPixPerThread];
#pragma unroll
    for (int i = 0; i < kPixPerThread; ++i) {
        seenLo[i] = 0u;
        wtri[i] = 0u;
        posx[i] = posy[i] = 0.f;
    }
    if (nRefs != 0u) {
#pragma unroll
        for (int i = 0; i < kPixPerThread; ++i) S.key[tid + kTileThreads * i] = 0ull;
    }
    const int bandH = a.y1 - a.y0;
    const int nEpochs = nRefs ? g.nEpochs : 0;
    for (int epoch = 0; epoch < nEpochs; ++epoch) {
        for (unsigned b0 = 0; b0 < nRefs; b0 += kTileThreads) {
            const int nb = (int)min((unsigned)kTileThreads, nRefs - b0);
            // ---- A1: one listed triangle per thread ----
            unsigned cnt = 0;
            if (tid < nb) {
                const unsigned ref = g.refs[beg + b0 + tid];
                int minY = 0, maxY = -1;
                if (ref & kBigRef) {
                    const TriSetup* s = g.bigTs + (ref & ~kBigRef);
                    if (((unsigned)s->tri >> kTriBits) == (unsigned)epoch) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            S.vx[k][tid] = s->vx[k];
                            S.vy[k][tid] = s->vy[k];
                            S.vz[k][tid] = s->vz[k];
                            S.vpx[k][tid] = s->vpx[k];
                            S.vpy[k][tid] = s->vpy[k];
                        }
                        minY = s->minY;
                        maxY = minY + s->rows - 1;
                        S.tri[tid] = (unsigned)s->tri;
                        S.sbase[tid] = s->sampleBase;
                    }
                } else if ((ref >> kTriBits) == (unsigned)epoch) {
                    float t[12];
                    load_vertices(a, (int)ref, t);
                    minY = INT_MAX;
                    maxY = INT_MIN;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const RPixel v = vertex_shader<true>(a.fr, mk3(t[3 * k], t[3 * k + 1], t[3 * k + 2]));  // :760-761
                        S.vx[k][tid] = v.x;
                        S.vy[k][tid] = v.y;
                        S.vz[k][tid] = v.zinv;
                        S.vpx[k][tid] = v.p.x;
                        S.vpy[k][tid] = v.p.y;
                        minY = min(minY, v.y);
                        maxY = max(maxY, v.y);
                    }
                    S.tri[tid] = ref;
                    S.sbase[tid] = kSmallTri;
                }
                const int ya = max(minY, Y0), yb = min(maxY, Y1 - 1);  // DrawRows :743, the band, the tile
                S.yA[tid] = ya;
                cnt = (unsigned)max(yb - ya + 1, 0);
            }
            {   // block-wide exclusive prefix sum of cnt -> rowBase[0..256]
                unsigned inc = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned x = __shfl_up_sync(0xffffffffu, inc, off);
                    if (lane >= off) inc += x;
                }
                if (lane == 31) S.warpSums[warp] = inc;
                __syncthreads();
                unsigned wbase = 0;
#pragma unroll
                for (int k = 0; k < kTileThreads / 32; ++k)
                    if (k < warp) wbase += S.warpSums[k];
                S.rowBase[tid] = wbase + inc - cnt;
                if (tid == kTileThreads - 1) S.rowBase[kTileThreads] = wbase + inc;
                __syncthreads();
            }
            // ---- rounds of whole triangles with at most kRoundRows row items between them ----
            int ja = 0;
            while (ja < nb) {
                const unsigned base = S.rowBase[ja];
                const bool fits = tid >= ja && tid < nb && S.rowBase[tid + 1] - base <= (unsigned)kRoundRows;
                const int nt = __syncthreads_count(fits);  // >= 1: one triangle has at most kTileH rows in the tile
                const int jb = ja + nt;
                const int nrows = (int)(S.rowBase[jb] - base);
                if (nrows == 0) {  // (uniform) nothing of these triangles lies on the tile's rows
                    ja = jb;
                    continue;
                }
                if (fits) {
                    const int r0 = (int)(S.rowBase[tid] - base), r1 = (int)(S.rowBase[tid + 1] - base);
                    for (int r = r0; r < r1; ++r) {
                        B2R_BOUND(r, kRoundRows);
                        S.rowTri[r] = (unsigned short)tid;
                    }
                }
                // ---- W: Interpolate (:615-637) of the small triangles' edges, one (triangle, edge) per thread ----
                for (int t = tid; t < 3 * nt; t += kTileThreads) {
                    const int jr = t / 3, e = t - 3 * jr, j = ja + jr, e2 = (e == 2) ? 0 : e + 1;  // :707
                    if (S.sbase[j] != kSmallTri) continue;
                    const int rj = (int)(S.rowBase[j] - base), cj = (int)(S.rowBase[j + 1] - base) - rj;
                    if (cj == 0) continue;
                    const int yLo = S.yA[j], yHi = yLo + cj - 1;
                    const int ya = S.vy[e][j], yb = S.vy[e2][j];
                    const int n = abs(ya - yb) + 1, sgn = (yb > ya) - (yb < ya);  // :712
                    // step k lies on row ya + sgn*k: the steps before the tile's rows only accumulate, the steps
                    // inside are stored, the steps after them are not needed by anyone
                    int kLo, kHi;
                    if (sgn > 0) {
                        kLo = yLo - ya;
                        kHi = yHi - ya;
                    } else if (sgn < 0) {
                        kLo = ya - yHi;
                        kHi = ya - yLo;
                    } else {
                        kLo = (ya >= yLo && ya <= yHi) ? 0 : 1;
                        kHi = 0;
                    }
                    kLo = max(kLo, 0);
                    kHi = min(kHi, n - 1);
                    if (kLo > kHi) continue;
                    const Recip div = recip_make((float)max(n - 1, 1));  // :622
                    const int ax = S.vx[e][j];
                    float cx = (float)ax, cz = S.vz[e][j], cpx = S.vpx[e][j], cpy = S.vpy[e][j];  // fPixel(Pixel&)
                    const float dx = (float)(S.vx[e2][j] - ax), dz = xsub(S.vz[e2][j], cz);  // Pixel operator-
                    const float dpx = xsub(S.vpx[e2][j], cpx), dpy = xsub(S.vpy[e2][j], cpy);
                    bool ok = div.ok;
                    float sx = xdiv_step_fast(dx, div, ok), sz = xdiv_step_fast(dz, div, ok);  // fPixel operator/
                    float spx = xdiv_step_fast(dpx, div, ok), spy = xdiv_step_fast(dpy, div, ok);
                    if (!ok) {
                        sx = xdiv_step(dx, div.b);
                        sz = xdiv_step(dz, div.b);
                        spx = xdiv_step(dpx, div.b);
                        spy = xdiv_step(dpy, div.b);
                    }
#pragma unroll 1
                    for (int k = 0; k < kLo; ++k) {  // :626-636, serial accumulation
                        cx = xadd(cx, sx);
                        cz = xadd(cz, sz);
                        cpx = xadd(cpx, spx);
                        cpy = xadd(cpy, spy);
                    }
                    int idx = 3 * (rj + (ya + sgn * kLo - yLo)) + e;
                    const int stride = 3 * sgn;
#pragma unroll 1
                    for (int k = kLo; k <= kHi; ++k, idx += stride) {
                        // int(current.x): the chain stays within coord_margin of the vertex range, which passed the
                        // +-2^24 limit, so the plain truncating conversion equals the x86 one
                        B2R_BOUND(idx, 3 * kRoundRows);
                        S.samples[idx] = make_float4(__int_as_float(__float2int_rz(cx)), cz, cpx, cpy);
                        cx = xadd(cx, sx);
                        cz = xadd(cz, sz);
                        cpx = xadd(cpx, spx);
                        cpy = xadd(cpy, spy);
                    }
                }
                __syncthreads();
                // ---- A3: ComputePolygonRows' resolve (:705-733) + DrawRows / DrawLineSDL / Bresenham (dy == 0) ----
                for (int ri = tid; ri < nrows; ri += kTileThreads) {
                    const int j = S.rowTri[ri];
                    const int y = S.yA[j] + ri - (int)(S.rowBase[j] - base);
                    const unsigned sb = S.sbase[j];
                    int lx = INT_MAX, rx = -INT_MAX;  // :696-697
                    float lz = 0.f, rz = 0.f, lpx = 0.f, lpy = 0.f, rpx = 0.f, rpy = 0.f;
                    unsigned off = sb;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {  // edges 0->1, 1->2, 2->0 in order; strict </> so the first edge to
                        const int e2 = (e == 2) ? 0 : e + 1;  // reach an extreme x keeps its attributes (:718, :726)
                        const int ya = S.vy[e][j], yb = S.vy[e2][j];
                        if (y >= min(ya, yb) && y <= max(ya, yb)) {
                            float4 q;
                            if (sb == kSmallTri) {
                                q = S.samples[3 * ri + e];
                            } else {
                                const unsigned at = a.bandSlots ? sb + (unsigned)(e * bandH + (y - a.y0)) : off + (unsigned)abs(y - ya);
                                q = *reinterpret_cast<const float4*>(g.samples + at);
                            }
                            const int x = __float_as_int(q.x);
                            if (x < lx) {
                                lx = x; lz = q.y; lpx = q.z; lpy = q.w;
                            }
                            if (x > rx) {
                                rx = x; rz = q.y; rpx = q.z; rpy = q.w;
                            }
                        }
                        off += (unsigned)(abs(ya - yb) + 1);
                    }
                    const int pixels = rx - lx;                                         // :598
                    const int i0 = max(0, X0 - lx - 1), i1 = min(pixels, X1 - lx - 1);  // :663 keeps 0 <= x < W; here: the tile's columns
                    if (i1 > i0) {
                        const Recip fdx = recip_make((float)pixels);
                        const float dz = xsub(rz, lz), dpx = xsub(rpx, lpx), dpy = xsub(rpy, lpy);
                        bool ok = fdx.ok;
                        float zstep = xdiv_step_fast(dz, fdx, ok);  // :648 (constant-depth rows: 0/n)
                        RowRec r;
                        r.lx = lx;
                        r.lpx = lpx;
                        r.lpy = lpy;
                        r.psx = xdiv_step_fast(dpx, fdx, ok);       // :649
                        r.psy = xdiv_step_fast(dpy, fdx, ok);
                        if (!ok) {
                            zstep = xdiv_step(dz, fdx.b);
                            r.psx = xdiv_step(dpx, fdx.b);
                            r.psy = xdiv_step(dpy, fdx.b);
                        }
                        r.tri = S.tri[j];
                        B2R_BOUND(ri, kRoundRows);
                        S.rec[ri] = r;
                        const unsigned lo = ((kTriMask - (r.tri & kTriMask)) << kSlotBits) | (unsigned)ri;
                        unsigned long long* krow = S.key + ((y - Y0) * kTileW - X0 + lx + 1);  // pixel of fragment q: krow[q]
                        B2R_BOUND((y - Y0) * kTileW - X0 + lx + 1 + i0, kTilePix);
                        B2R_BOUND((y - Y0) * kTileW - X0 + lx + i1, kTilePix);
                        for (int q = i0; q < i1; ++q) {
                            const float zinv = xadd(lz, xmul(zstep, (float)q));  // :667
                            if (zinv > 0.0f)                                     // :606 against a buffer cleared to 0 (:188)
                                atomicMax(krow + q, ((unsigned long long)__float_as_uint(zinv) << 32) | lo);
                        }
                        nTests += (unsigned long long)(i1 - i0);
                    }
                }
                __syncthreads();
                // ---- G: every pixel whose owner changed in this round fetches the owner's row (:667-668) ----
#pragma unroll
                for (int i = 0; i < kPixPerThread; ++i) {
                    const unsigned lo = (unsigned)S.key[tid + kTileThreads * i];
                    if (lo != seenLo[i]) {
                        seenLo[i] = lo;
                        B2R_BOUND(lo & ((1u << kSlotBits) - 1u), kRoundRows);
                        const RowRec* r = &S.rec[lo & ((1u << kSlotBits) - 1u)];
                        const float fi = (float)(X0 + lane - r->lx - 1);
                        posx[i] = xadd(r->lpx, xmul(r->psx, fi));
                        posy[i] = xadd(r->lpy, xmul(r->psy, fi));
                        wtri[i] = r->tri;
                    }
                }
                // no barrier here: the next round's first writes to anything read above come after its own barriers
                ja = jb;
            }
            __syncthreads();
        }
        if (epoch + 1 < nEpochs) {  // freeze: later epochs hold higher triangle indices and lose every tie
#pragma unroll
            for (int i = 0; i < kPixPerThread; ++i) {
                unsigned long long* k = &S.key[tid + kTileThreads * i];
                if (*k != 0ull) {
                    *k |= 0xFFFFFFFFull;
                    seenLo[i] = 0xFFFFFFFFu;
                }
            }
            __syncthreads();
        }
    }
    if (a.stats) {
        for (int off = 16; off > 0; off >>= 1) nTests += __shfl_xor_sync(0xffffffffu, nTests, off);
        if (lane == 0 && nTests) atomicAdd(a.stats + B2R_STAT_RAS_DEPTH_TESTS, nTests);
    }
    unsigned wbits[kPixPerThread];
#pragma unroll
    for (int i = 0; i < kPixPerThread; ++i) wbits[i] = nRefs ? (unsigned)(S.key[tid + kTileThreads * i] >> 32) : 0u;
    if (nJobs > 1u) {
        // The tile's list was split: every job leaves its per-pixel winners in its slot; the last one to arrive merges
        // them -- larger zinv first, lower triangle index among equals (:606) -- and shades.
        uint4* mine = g.partials + ((size_t)g.tileMslot[tile] + (size_t)slice) * kTilePix;
#pragma unroll
        for (int i = 0; i < kPixPerThread; ++i)
            mine[tid + kTileThreads * i] = make_uint4(wbits[i], wtri[i], __float_as_uint(posx[i]), __float_as_uint(posy[i]));
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            lastJob = atomicAdd(&g.tileDone[tile], 1u) == nJobs - 1u;
            if (lastJob) g.tileDone[tile] = 0u;  // re-armed for the next frame
        }
        __syncthreads();
        if (!lastJob) return;
        __threadfence();
        const uint4* all = g.partials + (size_t)g.tileMslot[tile] * kTilePix;
        for (unsigned sl = 0; sl < nJobs; ++sl) {
            if (sl == (unsigned)slice) continue;
#pragma unroll
            for (int i = 0; i < kPixPerThread; ++i) {
                const uint4 o = __ldcg(&all[(size_t)sl * kTilePix + tid + kTileThreads * i]);
                if (o.x > wbits[i] || (o.x == wbits[i] && o.x != 0u && o.y < wtri[i])) {
                    wbits[i] = o.x;
                    wtri[i] = o.y;
                    posx[i] = __uint_as_float(o.z);
                    posy[i] = __uint_as_float(o.w);
                }
            }
        }
    }
    // ---- C: PixelShader (:549-589) for the depth winner of every pixel, CalculateDOF/PutPixelSDL fused ----
    const int x = X0 + lane;
    if (x < X1) {
#pragma unroll
        for (int i = 0; i < kPixPerThread; ++i) {
            const int y = Y0 + warp + (kTileThreads / 32) * i;
            if (y >= Y1) break;
            float depth = 0.f, focal = 0.f;
            V3 colour = mk3(0.f, 0.f, 0.f);
            int winner = -1;
            const unsigned bits = wbits[i];
            if (bits != 0u) {
                winner = (int)wtri[i];
                depth = __uint_as_float(bits);
                const float4* t = reinterpret_cast<const float4*>(a.raw + (size_t)wtri[i] * 64);
                const float4 n4 = t[2], c4 = t[3];  // (v2.z, normal), (colour, isCulled)
                const V3 normal = mk3(n4.y, n4.z, n4.w), color = mk3(c4.x, c4.y, c4.z);
                pixel_shader_core<true>(a.fr, depth, mk3(posx[i], posy[i], 1.0f), normal, color, focal, colour);
            }
            const unsigned idx = (unsigned)y * (unsigned)a.W + (unsigned)x;  // W, H <= 32768
            if (a.depth) a.depth[idx] = depth;
            if (a.colours) {
                float* c3 = a.colours + (size_t)idx * 3;
                c3[0] = colour.x;
                c3[1] = colour.y;
                c3[2] = colour.z;
            }
            if (a.focal) a.focal[idx] = focal;
            if (a.winner) a.winner[idx] = winner;
            // CalculateDOF without depth of field + PutPixelSDL (:516-526), fused
            if (a.surface) a.surface[idx] = inside_border(x, y, a.W, a.H) ? pack_xrgb(colour.x, colour.y, colour.z) : 0u;
        }
    }
}

// ---- culling block of Update() (:385-447) ------------------------------------
struct CullConsts {
    float m00, m11, m22, m32;
};

__global__ void ras_cull_kernel(const unsigned char* __restrict__ raw, int stride, int T, const DevFrame* __restrict__ f,
                                CullConsts cc, int backface, int frustum, unsigned char* __restrict__ culled) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const float* t = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    const V3 cam = mk3(f->cam[0], f->cam[1], f->cam[2]);
    int cull = 0;
    if (backface) {  // :408-414
        if (xdot3(xsub3(mk3(t[0], t[1], t[2]), cam), mk3(t[9], t[10], t[11])) > 0.0f) cull = 1;
    }
    if (frustum && !cull) {  // :416-446
        bool anyInside = false;
        for (int k = 0; k < 3; ++k) {
            const V3 v = xvec_mat(xsub3(mk3(t[3 * k], t[3 * k + 1], t[3 * k + 2]), cam), f->R);  // :423-425
            // vec4(v,1) * transform, all sixteen terms of glm/detail/type_mat4x4.inl:664-675
            const float v3 = 1.0f, z0 = 0.0f;
            float X = xadd(xadd(xadd(xmul(cc.m00, v.x), xmul(z0, v.y)), xmul(z0, v.z)), xmul(z0, v3));
            float Y = xadd(xadd(xadd(xmul(z0, v.x), xmul(cc.m11, v.y)), xmul(z0, v.z)), xmul(z0, v3));
            float Z = xadd(xadd(xadd(xmul(z0, v.x), xmul(z0, v.y)), xmul(cc.m22, v.z)), xmul(z0, v3));
            float Wc = xadd(xadd(xadd(xmul(z0, v.x), xmul(z0, v.y)), xmul(cc.m32, v.z)), xmul(z0, v3));
            X = xdiv(X, Wc);  // :435-437
            Y = xdiv(Y, Wc);
            Z = xdiv(Z, Wc);
            anyInside = anyInside || (X >= -1.0f && X <= 1.0f && Y >= -1.0f && Y <= 1.0f && Z >= 0.0f && Z <= 1.0f);  // :453
        }
        if (!anyInside) cull = 1;  // :444-445
    }
    culled[i] = (unsigned char)cull;
}

// ---- 60-byte scenes (the raytracer's Triangle) as 64-byte records, once per b2r_set_triangles ------------
__global__ void ras_repack_kernel(const unsigned char* __restrict__ raw, int stride, int T, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const float* r = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    float v[16];
#pragma unroll
    for (int k = 0; k < 15; ++k) v[k] = r[k];
    v[15] = 0.f;  // isCulled lives in its own array (b2r_ras_cull / b2r_set_culled)
#pragma unroll
    for (int k = 0; k < 4; ++k) out[(size_t)i * 4 + k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

cudaError_t launch_ras_repack(Ctx* c, cudaStream_t s) {
    if (c->stride == 64 || c->T == 0) return cudaSuccess;
    cudaError_t e = c->raw64.reserve((size_t)c->T * 64);
    if (e != cudaSuccess) return e;
    ras_repack_kernel<<<(c->T + 255) / 256, 256, 0, s>>>(c->raw.as<unsigned char>(), c->stride, c->T, c->raw64.as<float4>());
    c->launches++;
    return cudaGetLastError();
}

// ---- host side ------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

cudaError_t launch_ras_cull(Ctx* c, unsigned char* d_culled, cudaStream_t s) {
    // Frame constants of :385-402, evaluated on the host in reference order (float, libm acosf/tanf like the
    // reference's cos/sin/acos/tan calls -- none of it is per-triangle work).
    const b2r_frame_params& p = c->params;
    V3 fv = xnormalize3(xvec_mat(mk3(0.f, 0.f, 1.0f), p.cameraRot));                    // :385
    float nearZ = xadd(p.cameraPos[2], xmul(fv.z, 0.1f)), farZ = xadd(p.cameraPos[2], xmul(fv.z, 15.0f));  // :386
    float w = (float)c->W, h = (float)c->H;
    V3 tv = mk3(0.0f, -h / 2.0f, p.focalLength), bv = mk3(0.0f, h / 2.0f, p.focalLength);  // :392-393
    float cy = xdot3(tv, bv) / (xsqrt(xdot3(tv, tv)) * xsqrt(xdot3(bv, bv)));           // :394
    float rfovy = acosf(cy);                                                            // :395
    float aspect = w / h;
    CullConsts cc;
    cc.m00 = (1.0f / tanf(rfovy / 2.0f)) / aspect;  // :398
    cc.m11 = (1.0f / tanf(rfovy / 2.0f));           // :399
    cc.m22 = farZ / (farZ - nearZ);                 // :400
    cc.m32 = 1.0f;                                  // :401-402
    if (c->T == 0) return cudaSuccess;
    ras_cull_kernel<<<(c->T + 255) / 256, 256, 0, s>>>(c->raw.as<unsigned char>(), c->stride, c->T,
                                                       c->frame.as<DevFrame>(), cc, p.backfaceCulling,
                                                       p.frustumCulling, d_culled);
    c->launches++;
    return cudaGetLastError();
}

// ---- host side ------------------------------------------------------------------
// Returns cudaErrorInvalidValue when a triangle exceeds the row/coordinate limits (-> B2R_E_CAPACITY).
// Scenes of few triangles (T * band height within this budget) give every large triangle a fixed-capacity slot of edge
// samples and size the tile lists for the worst case: nothing is read back, the frame is a plain sequence of launches.
// Larger scenes size both from counters read back after the first two kernels -- once per (scene, culling flags, frame
// params, band): the counts are a pure function of those, so later frames of the same state run without the readback.
constexpr size_t kBandSlotLimit = 2u << 20;  // edge-sample slots (T * band height)
constexpr size_t kBandRefLimit = 1u << 20;   // tile-list entries (T * tiles)

cudaError_t ras_take_error(Ctx* c) {
    if (!c->rasErrPending) return cudaSuccess;
    c->rasErrPending = false;
    RasCounters* ctr = reinterpret_cast<RasCounters*>(c->rasErrCtr);  // of the pipeline that drew last
    if (!ctr) return cudaSuccess;
    unsigned* flag = reinterpret_cast<unsigned*>(c->pinned);
    cudaError_t e = cudaMemcpyAsync(flag, &ctr->sticky, sizeof *flag, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
    if (!*flag) return cudaSuccess;
    e = cudaMemsetAsync(&ctr->sticky, 0, sizeof *flag, c->stream);
    return e != cudaSuccess ? e : cudaErrorInvalidValue;
}

cudaError_t launch_ras_draw_sortlast(Ctx* c, const RasLaunch& a0, cudaStream_t s);
static cudaError_t launch_ras_draw_tiles(Ctx* c, const RasLaunch& a0, cudaStream_t s);

// B2R_OPT_RAS_VARIANT: 0 sort-last pipeline (default), 1 the same without fixed-capacity slots, 2 screen tiles,
// 3 screen tiles without fixed-capacity slots.  Every variant produces the same bits.
cudaError_t launch_ras_draw(Ctx* c, const RasLaunch& a, cudaStream_t s) {
    return c->optRasVariant >= 2 ? launch_ras_draw_tiles(c, a, s) : launch_ras_draw_sortlast(c, a, s);
}

static cudaError_t launch_ras_draw_tiles(Ctx* c, const RasLaunch& a0, cudaStream_t s) {
    RasLaunch a = a0;
    const int T = a.T;
    const int bandH = a.y1 - a.y0;
    const int tilesX = (a.W + kTileW - 1) >> kTileShift, tilesY = (bandH + kTileH - 1) >> kTileShift, nTiles = tilesX * tilesY;
    cudaError_t e;
    // scratch: [counters 256 B][tileCount][tileOffset][tileMslot][tileDone][bigCounts uint2 x T][excl uint2 x T][blockSums][totals]
    const int nbMax = (T + kScanBlock - 1) / kScanBlock + 1;
    const size_t offCtr = 0, offTileCount = 256, offTileOffset = align_up(offTileCount + sizeof(unsigned) * (size_t)nTiles, 256),
                 offMslot = align_up(offTileOffset + sizeof(unsigned) * (size_t)(nTiles + 2), 256),
                 offDone = align_up(offMslot + sizeof(unsigned) * (size_t)nTiles, 256),
                 offCounts = align_up(offDone + sizeof(unsigned) * (size_t)nTiles, 256),
                 offExcl = align_up(offCounts + sizeof(uint2) * (size_t)T, 256),
                 offSums = align_up(offExcl + sizeof(uint2) * (size_t)T, 256),
                 offTotals = align_up(offSums + sizeof(uint2) * (size_t)nbMax, 256), scratchBytes = offTotals + 256;
    const void* scratchBefore = c->rasScratch.p;
    if ((e = c->rasScratch.reserve(scratchBytes)) != cudaSuccess) return e;
    unsigned char* sc = c->rasScratch.as<unsigned char>();
    if (c->rasScratch.p != scratchBefore) {  // fresh memory: the sticky error flag starts clear
        if ((e = cudaMemsetAsync(sc, 0, 256, s)) != cudaSuccess) return e;
        c->rasCtrDirty = true;
    }
    if ((e = c->rasTri.reserve(sizeof(unsigned) * (size_t)(T + 1) + 256 + sizeof(TriSetup) * (size_t)(T + 1))) != cudaSuccess) return e;
    RasCounters* ctr = reinterpret_cast<RasCounters*>(sc + offCtr);
    unsigned* tileCount = reinterpret_cast<unsigned*>(sc + offTileCount);
    unsigned* tileOffset = reinterpret_cast<unsigned*>(sc + offTileOffset);
    unsigned* tileMslot = reinterpret_cast<unsigned*>(sc + offMslot);
    unsigned* tileDone = reinterpret_cast<unsigned*>(sc + offDone);
    uint2* counts = reinterpret_cast<uint2*>(sc + offCounts);
    uint2* excl = reinterpret_cast<uint2*>(sc + offExcl);
    uint2* sums = reinterpret_cast<uint2*>(sc + offSums);
    uint2* totals = reinterpret_cast<uint2*>(sc + offTotals);
    unsigned* triWord = c->rasTri.as<unsigned>();
    TriSetup* ts = reinterpret_cast<TriSetup*>(c->rasTri.as<unsigned char>() + align_up(sizeof(unsigned) * (size_t)(T + 1), 256));

    const bool bandSlots = (size_t)T * (size_t)bandH <= kBandSlotLimit && (size_t)T * (size_t)nTiles <= kBandRefLimit &&
                           c->optRasVariant != 3;
    a.bandSlots = bandSlots ? 1 : 0;
    // the per-frame counters and the tile counts are left clear by the previous frame's tile kernel; clear them here
    // only after a draw that did not get that far, on a fresh buffer, or when the tile grid changed shape
    // (the sticky flag, second half of the counter block, survives until the host has reported it)
    if (c->rasCtrDirty || c->rasTilesClean != (size_t)nTiles) {
        if ((e = cudaMemsetAsync(ctr, 0, offsetof(RasCounters, sticky), s)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(tileCount, 0, offTileOffset - offTileCount, s)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(tileDone, 0, offCounts - offDone, s)) != cudaSuccess) return e;
    }
    c->rasCtrDirty = true;

    // job table of ras_tile: a tile whose list is longer than kSlice is split over ceil(len / kSlice) CTAs
    unsigned nBig = 0, nSamples = 0;
    size_t refCap, sampleCap;
    Ctx::RasSizes& z = c->rasSizes;
    const bool sized = z.valid && z.gen == c->rasGen && z.y0 == a.y0 && z.y1 == a.y1;
    if (bandSlots) refCap = (size_t)T * (size_t)nTiles;             // worst case: every triangle in every tile
    else refCap = sized ? z.totalRefs : (size_t)T + (size_t)T / 2;  // known, or a guess corrected below
    size_t jobCap = (size_t)nTiles + refCap / kSlice + 1;
    if ((e = c->rasJobs.reserve(sizeof(uint2) * jobCap)) != cudaSuccess) return e;
    jobCap = c->rasJobs.cap / sizeof(uint2);

    ras_setup_kernel<<<std::max(1, std::min((T + 255) / 256, c->smCount * 5)), 256, 0, s>>>(a, tilesX, nTiles, triWord, ts, counts, tileCount, tileOffset,
                                                               tileMslot, c->rasJobs.as<uint2>(), (unsigned)jobCap, ctr);
    c->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;

    if (bandSlots) {
        sampleCap = (size_t)T * 3 * (size_t)bandH;
        c->rasErrPending = true;  // capacity flag: checked by the caller's next synchronising call (ras_take_error)
        c->rasErrCtr = ctr;
    } else {
        if (!sized) {
            // how many large triangles / edge samples / tile-list entries / jobs: 32 bytes back to size the buffers
            if ((e = cudaMemcpyAsync(c->pinned, ctr, sizeof(RasCounters), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
            const RasCounters host = *reinterpret_cast<RasCounters*>(c->pinned);
            if (host.err) {
                cudaMemsetAsync(&ctr->sticky, 0, sizeof(unsigned), s);  // reported right here
                return cudaErrorInvalidValue;
            }
            z.valid = true;
            z.gen = c->rasGen;
            z.y0 = a.y0;
            z.y1 = a.y1;
            z.nBig = host.nBig;
            z.bigSamples = host.bigSamples;
            z.totalRefs = host.totalRefs;
            if (host.totalJobs > jobCap) {  // the guess was short: size the table and rebuild it from the offsets
                jobCap = (size_t)host.totalJobs;
                if ((e = c->rasJobs.reserve(sizeof(uint2) * jobCap)) != cudaSuccess) return e;
                ras_jobs_kernel<<<1, 256, 0, s>>>(tileOffset, tileMslot, c->rasJobs.as<uint2>(), (unsigned)jobCap, nTiles);
                c->launches++;
            }
        } else {
            c->rasErrPending = true;  // cannot be set for a state that drew cleanly before; kept for symmetry
            c->rasErrCtr = ctr;
        }
        nBig = z.nBig;
        nSamples = z.bigSamples;
        refCap = z.totalRefs;
        sampleCap = nSamples;
    }
    const size_t nJobsMax = std::min(jobCap, (size_t)nTiles + refCap / kSlice + 1);  // grid of ras_tile (surplus CTAs exit)
    if ((e = c->rasRefs.reserve(sizeof(unsigned) * refCap + 256)) != cudaSuccess) return e;
    if ((e = c->rasRows.reserve(sizeof(EdgeSample) * sampleCap + 256)) != cudaSuccess) return e;
    if ((e = c->rasPartials.reserve(sizeof(uint4) * kTilePix * (2 * (refCap / kSlice) + 2))) != cudaSuccess) return e;
    unsigned* refs = c->rasRefs.as<unsigned>();
    EdgeSample* samples = c->rasRows.as<EdgeSample>();

    if (T > 0) {
        ras_bin_kernel<<<(T + 255) / 256, 256, 0, s>>>(triWord, T, tilesX, ts, tileCount, tileOffset, refs);
        c->launches++;
    }
    if (bandSlots) {
        if (T > 0) {
            const int eb = 12 * T <= c->smCount * 32 * 4 ? 32 : 128;  // one warp per CTA while SMs would be left without one
            ras_edges_kernel<true><<<(12 * T + eb - 1) / eb, eb, 0, s>>>(ts, T, ctr, samples, a.y0, a.y1);
            c->launches++;
        }
    } else if (nBig > 0) {
        const int nb = ((int)nBig + kScanBlock - 1) / kScanBlock;
        scan_blocks_kernel<<<nb, kScanBlock, 0, s>>>(counts, excl, sums, (int)nBig);
        scan_sums_kernel<<<1, kScanBlock, 0, s>>>(sums, nb, totals);
        scan_apply_kernel<<<nb, kScanBlock, 0, s>>>(excl, sums, ts, (int)nBig);
        const int eb = 12 * (int)nBig <= c->smCount * 32 * 4 ? 32 : 128;
        ras_edges_kernel<false><<<(12 * (int)nBig + eb - 1) / eb, eb, 0, s>>>(ts, (int)nBig, ctr, samples, a.y0, a.y1);
        c->launches += 4;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    {
        TileArgs g;
        g.tileOffset = tileOffset;
        g.refs = refs;
        g.tileCount = tileCount;
        g.bigTs = ts;
        g.samples = samples;
        g.ctr = ctr;
        g.tileMslot = tileMslot;
        g.jobTile = c->rasJobs.as<uint2>();
        g.tileDone = tileDone;
        g.partials = c->rasPartials.as<uint4>();
        g.tilesX = tilesX;
        g.nTiles = nTiles;
        g.nEpochs = (int)(((long long)T + (1ll << kTriBits) - 1) >> kTriBits);
        if (!c->rasTileAttr) {  // opt in to > 48 KB of shared memory, once per context (= per device)
            if ((e = cudaFuncSetAttribute(ras_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileShared))) != cudaSuccess) return e;
            c->rasTileAttr = true;
        }
        ras_tile_kernel<<<(unsigned)nJobsMax, kTileThreads, sizeof(TileShared), s>>>(a, g);
        c->launches++;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) {
        c->rasCtrDirty = false;
        c->rasTilesClean = (size_t)nTiles;
    }
    return e;
}

}  // namespace b2r
