// b2r_group.cpp -- several GPUs behind one Draw(): the b2r_group_* part of include/b2r.h.
//
// The reference parallelises Draw() over image rows with OpenMP (raytracer.cpp:557, rasteriser.cpp:467); a group does
// the same over the GPUs of one box, in ONE process: one context per device, scene and frame params replicated, the
// caller's host surface as the meeting point (every device copies its own rows over its own PCIe link), and frames of
// an animation handed out round-robin.  There is no data-path collective on any of these paths; the
// device-resident gather over NVLink is b2r_rt_frame_gather_device_async.
//
// Written against the public C ABI plus the CUDA runtime calls a host needs for page-locking; no kernel lives here.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b2r.h"

struct b2r_group {
    int n = 0, W = 0, H = 0;
    std::vector<int> dev;
    std::vector<b2r_ctx*> ctx;
    std::string err;
    void* pinned = nullptr;  // caller surface page-locked on the caller's behalf
    void* seen = nullptr;    // last caller surface looked at (locked by us or already by the caller)
    size_t seenBytes = 0;
};

namespace {

thread_local std::string g_groupCreateError;

int gfail(b2r_group* g, int code, const std::string& what) {
    if (g) g->err = what;
    else g_groupCreateError = what;
    return code;
}

int member_fail(b2r_group* g, int i, int rc, const char* what) {
    char buf[640];
    snprintf(buf, sizeof buf, "%s on device %d: %s", what, g->dev[i], b2r_last_error(g->ctx[i]));
    g->err = buf;
    return rc;
}

// The host surface every device copies into: page-locked once (portable, so that every device's copy engine can use
// it), remembered until the caller passes a different buffer.  Failure to lock is not an error: the copies then go
// through the driver's staging buffer.
void pin_for_group(b2r_group* g, void* p, size_t bytes) {
    if (g->seen == p && g->seenBytes >= bytes) return;  // this buffer has been dealt with (locked here, or by the caller)
    if (g->pinned) cudaHostUnregister(g->pinned);
    g->pinned = nullptr;
    g->seen = p;
    g->seenBytes = bytes;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) {
        cudaGetLastError();
        return;  // already page-locked by the caller (cudaHostAlloc / cudaHostRegister)
    }
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess) g->pinned = p;
    cudaGetLastError();  // not lockable: fine, the copies go through the driver's staging buffer
}

// ---- writer threads of b2r_group_rt_frames: BMP files (or copies into the caller's array) off the GPU threads ----
struct FrameJob {
    int frame;
    int slot;
};

struct Writers {
    std::mutex m;
    std::condition_variable cvJob, cvDone;
    std::deque<FrameJob> jobs;
    std::vector<char> slotBusy;
    std::vector<std::thread> threads;
    bool stop = false;
    int failed = 0;
    int pending = 0;  // jobs queued or being written
};

}  // namespace

extern "C" {

int b2r_group_create(b2r_group** out, const int* devices, int n, int width, int height) {
    if (!out || !devices || n < 1 || n > B2R_MAX_PEERS) return gfail(nullptr, B2R_E_INVALID, "b2r_group_create: 1..8 devices");
    *out = nullptr;
    b2r_group* g = new (std::nothrow) b2r_group();
    if (!g) return B2R_E_CUDA;
    g->n = n;
    g->W = width;
    g->H = height;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) {
                b2r_group_destroy(g);
                return gfail(nullptr, B2R_E_INVALID, "b2r_group_create: a device is listed twice");
            }
        b2r_ctx* c = nullptr;
        const int rc = b2r_create(&c, devices[i], width, height);
        if (rc) {
            const std::string why = std::string("b2r_group_create: ") + b2r_last_error(nullptr);
            b2r_group_destroy(g);
            return gfail(nullptr, rc, why);
        }
        g->dev.push_back(devices[i]);
        g->ctx.push_back(c);
    }
    *out = g;
    return B2R_OK;
}

int b2r_group_destroy(b2r_group* g) {
    if (!g) return B2R_OK;
    for (b2r_ctx* c : g->ctx) b2r_destroy(c);
    if (g->pinned) cudaHostUnregister(g->pinned);
    cudaGetLastError();
    delete g;
    return B2R_OK;
}

int b2r_group_size(const b2r_group* g) { return g ? g->n : 0; }

b2r_ctx* b2r_group_ctx(b2r_group* g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : nullptr; }

const char* b2r_group_last_error(const b2r_group* g) { return g ? g->err.c_str() : g_groupCreateError.c_str(); }

int b2r_group_set_triangles(b2r_group* g, const void* triangles, int count, int stride_bytes) {
    if (!g) return B2R_E_INVALID;
    for (int i = 0; i < g->n; ++i)
        if (int rc = b2r_set_triangles(g->ctx[i], triangles, count, stride_bytes)) return member_fail(g, i, rc, "b2r_set_triangles");
    return B2R_OK;
}

int b2r_group_set_frame(b2r_group* g, const b2r_frame_params* params) {
    if (!g) return B2R_E_INVALID;
    for (int i = 0; i < g->n; ++i)
        if (int rc = b2r_set_frame(g->ctx[i], params)) return member_fail(g, i, rc, "b2r_set_frame");
    return B2R_OK;
}

// Draw() of the raytracer, one frame over the group: rows are the parallel axis (raytracer.cpp:557-558).
int b2r_group_rt_frame(b2r_group* g, uint32_t* surface) {
    if (!g || !surface) return gfail(g, B2R_E_INVALID, "b2r_group_rt_frame: null argument");
    pin_for_group(g, surface, (size_t)g->W * g->H * 4);
    int first = B2R_OK;
    for (int i = 0; i < g->n; ++i)  // every device is busy before the first one is waited for
        if (int rc = b2r_rt_frame_part_async(g->ctx[i], i, g->n, surface)) {
            first = member_fail(g, i, rc, "b2r_rt_frame_part_async");
            break;
        }
    for (int i = 0; i < g->n; ++i)
        if (int rc = b2r_synchronize(g->ctx[i]))
            if (!first) first = member_fail(g, i, rc, "b2r_synchronize");
    return first;
}

// Draw() of the rasteriser, sort-first: contiguous row bands, triangle list replicated (rasteriser.cpp:467-478).
int b2r_group_ras_frame(b2r_group* g, uint32_t* surface) {
    if (!g || !surface) return gfail(g, B2R_E_INVALID, "b2r_group_ras_frame: null argument");
    pin_for_group(g, surface, (size_t)g->W * g->H * 4);
    int first = B2R_OK;
    const int base = g->H / g->n, extra = g->H % g->n;
    for (int i = 0; i < g->n; ++i) {
        const int y0 = i * base + (i < extra ? i : extra), y1 = y0 + base + (i < extra ? 1 : 0);
        if (int rc = b2r_ras_frame_part_async(g->ctx[i], y0, y1, surface)) {
            first = member_fail(g, i, rc, "b2r_ras_frame_part_async");
            break;
        }
    }
    for (int i = 0; i < g->n; ++i)
        if (int rc = b2r_synchronize(g->ctx[i]))
            if (!first) first = member_fail(g, i, rc, "b2r_synchronize");
    return first;
}

// An animation: frame f is rendered whole by device f mod n.  Every device owns a small ring of
// page-locked frame buffers; a finished frame is handed to a writer thread -- which writes the BMP file, or copies
// the pixels into the caller's array -- and the device goes on with its next frame.
int b2r_group_rt_frames(b2r_group* g, const b2r_frame_params* frames, int nframes, uint32_t* surfaces, const char* bmp_pattern) {
    if (!g || !frames || nframes < 0 || (!surfaces && !bmp_pattern))
        return gfail(g, B2R_E_INVALID, "b2r_group_rt_frames: needs frames and a destination (surfaces or bmp_pattern)");
    if (nframes == 0) return B2R_OK;
    const bool bmp = bmp_pattern != nullptr;
    const size_t npx = (size_t)g->W * g->H;
    const size_t slotBytes = bmp ? b2r_bmp_payload_bytes(g->W, g->H) : npx * 4;
    // Per device: up to kInFlight frames on the GPU, the other slots of its ring with the writer threads.  Writing a 25 MB
    // file takes several milliseconds (a frame is traced and copied in well under one), so what sets the pace is how
    // many files are being written at once (one GPU, 4K, /dev/shm: 3 slots 130-195 frames/s, 8 slots 520, 12 slots 550).
    const int kInFlight = 2;
    int kRing = 8, nSlots = 0;
    std::vector<void*> slots;
    int rcAll = B2R_OK;
    for (;;) {  // page-locked memory is a limited resource: fall back to a shorter ring rather than fail
        nSlots = g->n * kRing;
        slots.assign(nSlots, nullptr);
        bool ok = true;
        for (int s = 0; s < nSlots && ok; ++s)
            if (cudaHostAlloc(&slots[s], slotBytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
            }
        if (ok) break;
        for (void* p : slots)
            if (p) cudaFreeHost(p);
        if (kRing == 3) {
            slots.assign(nSlots, nullptr);
            rcAll = gfail(g, B2R_E_CUDA, "b2r_group_rt_frames: cudaHostAlloc of the frame slots failed");
            break;
        }
        kRing = kRing > 4 ? 4 : 3;
    }
    Writers w;
    w.slotBusy.assign(nSlots, 0);
    const int W = g->W, H = g->H;
    auto work = [&]() {
        for (;;) {
            FrameJob j;
            {
                std::unique_lock<std::mutex> lk(w.m);
                w.cvJob.wait(lk, [&] { return w.stop || !w.jobs.empty(); });
                if (w.jobs.empty()) return;
                j = w.jobs.front();
                w.jobs.pop_front();
            }
            int rc = B2R_OK;
            if (bmp) {
                char name[1024];
                snprintf(name, sizeof name, bmp_pattern, j.frame);
                rc = b2r_write_bmp(name, (const uint8_t*)slots[j.slot], W, H);
            }
            if (surfaces) {
                if (bmp) {  // both asked for: the pixels of the array come from the 24-bit payload (bottom-up, padded rows)
                    const size_t rowBytes = ((size_t)W * 3 + 3) & ~(size_t)3;
                    for (int y = 0; y < H; ++y) {
                        const uint8_t* src = (const uint8_t*)slots[j.slot] + (size_t)(H - 1 - y) * rowBytes;
                        uint32_t* dst = surfaces + (size_t)j.frame * npx + (size_t)y * W;
                        for (int x = 0; x < W; ++x) dst[x] = ((uint32_t)src[3 * x + 2] << 16) | ((uint32_t)src[3 * x + 1] << 8) | src[3 * x];
                    }
                } else {
                    memcpy(surfaces + (size_t)j.frame * npx, slots[j.slot], npx * 4);
                }
            }
            {
                std::lock_guard<std::mutex> lk(w.m);
                if (rc) w.failed = rc;
                w.slotBusy[j.slot] = 0;
                --w.pending;
            }
            w.cvDone.notify_all();
        }
    };
    if (!rcAll) {
        unsigned hw = std::thread::hardware_concurrency();
        const int nThreads = (int)std::max(1u, std::min(hw ? hw : 4u, (unsigned)((kRing - kInFlight) * g->n)));
        for (int t = 0; t < nThreads; ++t) w.threads.emplace_back(work);
    }
    // per device: frames in flight as (frame, slot), oldest first
    std::vector<std::deque<FrameJob>> flight(g->n);
    std::vector<int> next(g->n), ringPos(g->n, 0);
    for (int i = 0; i < g->n; ++i) next[i] = i;
    auto retire = [&](int i) -> int {  // wait for device i's oldest frame and hand it to the writers
        // one stream per device: synchronising it retires every frame in flight there
        if (int rc = b2r_synchronize(g->ctx[i])) return member_fail(g, i, rc, "b2r_synchronize");
        {
            std::lock_guard<std::mutex> lk(w.m);
            for (const FrameJob& f : flight[i]) {
                w.jobs.push_back(f);
                ++w.pending;
            }
        }
        flight[i].clear();
        w.cvJob.notify_all();
        return B2R_OK;
    };
    bool more = !rcAll;
    while (more && !rcAll) {
        more = false;
        for (int i = 0; i < g->n && !rcAll; ++i) {
            if (next[i] >= nframes) continue;
            more = true;
            const int slot = i * kRing + ringPos[i];
            if ((int)flight[i].size() == kInFlight) rcAll = retire(i);
            if (rcAll) break;
            {   // the slot's previous frame must have been written out
                std::unique_lock<std::mutex> lk(w.m);
                w.cvDone.wait(lk, [&] { return !w.slotBusy[slot]; });
                w.slotBusy[slot] = 1;
            }
            const int f = next[i];
            int rc = b2r_set_frame(g->ctx[i], &frames[f]);
            if (!rc) rc = bmp ? b2r_rt_frame_bgr8_async(g->ctx[i], (uint8_t*)slots[slot])
                              : b2r_rt_frame_part_async(g->ctx[i], 0, 1, (uint32_t*)slots[slot]);
            if (rc) {
                rcAll = member_fail(g, i, rc, "frame of b2r_group_rt_frames");
                break;
            }
            flight[i].push_back({f, slot});
            ringPos[i] = (ringPos[i] + 1) % kRing;
            next[i] += g->n;
        }
    }
    for (int i = 0; i < g->n; ++i)
        if (!flight[i].empty()) {
            const int rc = retire(i);
            if (rc && !rcAll) rcAll = rc;
        }
    {
        std::unique_lock<std::mutex> lk(w.m);
        w.cvDone.wait(lk, [&] { return w.pending == 0; });  // (slots of frames that failed before reaching a writer stay marked)
        w.stop = true;
    }
    w.cvJob.notify_all();
    for (std::thread& t : w.threads) t.join();
    for (void* p : slots)
        if (p) cudaFreeHost(p);
    if (!rcAll && w.failed) rcAll = gfail(g, w.failed, "b2r_group_rt_frames: writing a BMP file failed");
    return rcAll;
}

}  // extern "C"
