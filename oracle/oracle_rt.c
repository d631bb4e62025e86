/* oracle_rt.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference raytracer's per-pixel hot path,
 * plain C, runtime screen size.  It is the checker for the CUDA path and is
 * itself pinned against the compiled reference (oracle/_ref) by
 * tests/test_oracle_vs_reference.py and against tests/golden/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call it; the product library never does.
 *
 * Follows, function by function (all citations: raytracer/Source/raytracer.cpp):
 *   oracle_rt_closest_intersection   ClosestIntersection   :202-257
 *   oracle_rt_direct_light           DirectLight           :265-327
 *   oracle_rt_draw                   Draw (pixel loop)     :547-603  + reset :335-339
 * Deviations, each mechanical: P2 row stride is the
 * screen width (reference: y*SCREEN_HEIGHT, identical for square screens);
 * P5 never-hit pixels report triangleIndex -1, position 0.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (no -march, no -ffast-math).
 */
#include <float.h>
#include <stdint.h>
#include <string.h>

#include "../include/b2r.h"
#include "oracle_math.h"

#define ORACLE_API __attribute__((visibility("default")))

typedef struct {
    ovec3 position;
    float distance;
    int32_t triangleIndex;
} oisect; /* == struct Intersection, :91-96 */

static inline ovec3 ld3(const float* p) { return ov(p[0], p[1], p[2]); }

/* Brute force over every triangle in index order (:208).  tris15: v0 v1 v2 normal color. */
static int closest_intersection(ovec3 start, ovec3 dir, const float* tris15, int T, oisect* closest,
                                int isLight, float* focalSlot, float dofFocal) {
    int any = 0;
    for (int i = 0; i < T; ++i) {
        const float* t = tris15 + 15 * i;
        ovec3 v0 = ld3(t), v1 = ld3(t + 3), v2 = ld3(t + 6);
        ovec3 e1 = osub(v1, v0);      /* :216 */
        ovec3 e2 = osub(v2, v0);      /* :217 */
        ovec3 b = osub(start, v0);    /* :218 */
        ovec3 e1e2 = ocross(e1, e2);  /* :225 */
        ovec3 be2 = ocross(b, e2);    /* :226 */
        ovec3 e1b = ocross(e1, b);    /* :227 */
        ovec3 nd = oneg(dir);         /* :229 */
        /* :231-234 hand-written dots, left to right */
        float e1e2b = e1e2.x * b.x + e1e2.y * b.y + e1e2.z * b.z;
        float e1e2d = e1e2.x * nd.x + e1e2.y * nd.y + e1e2.z * nd.z;
        float be2d = be2.x * nd.x + be2.y * nd.y + be2.z * nd.z;
        float e1bd = e1b.x * nd.x + e1b.y * nd.y + e1b.z * nd.z;
        float tt = e1e2b / e1e2d, u = be2d / e1e2d, v = e1bd / e1e2d; /* :237 */
        if (u + v <= 1.0f && u >= 0.0f && v >= 0.0f && tt >= 0.0f) {  /* :239 */
            ovec3 pos = oadd(oadd(v0, oscale(e1, u)), oscale(e2, v)); /* :241  v0 + (u*e1) + (v*e2) */
            float distance = odistance(start, pos);                   /* :242 */
            if (closest->distance >= distance) {                      /* :243 ties -> later index */
                closest->position = pos;
                closest->distance = distance;
                closest->triangleIndex = i;
                if (!isLight) *focalSlot = distance - dofFocal;       /* :248-249 */
            }
            any = 1;                                                  /* :251 */
        }
    }
    return any;
}

static ovec3 direct_light(const oisect* hit, const float* tris15, int T, const b2r_frame_params* fp) {
    ovec3 result = ov(0, 0, 0), result2 = ov(0, 0, 0);
    int samples = fp->softShadowsEnabled ? fp->softShadowsSamples : 1; /* :272-275 */
    const float* tri = tris15 + 15 * hit->triangleIndex;
    for (int k = 0; k < fp->numLights; ++k) {
        for (int s = 0; s < samples; ++s) {
            const b2r_light* L = &fp->lights[k];
            ovec3 lightColor = oscale(ld3(L->color), L->intensity);   /* :282 */
            ovec3 position = (samples != 1)
                                 ? ld3(fp->randomPositions + 3 * (k * fp->softShadowsSamples + s)) /* :286 */
                                 : ld3(L->position);                                               /* :290 */
            float r = odistance(hit->position, position);             /* :294 */
            float A = (float)(4 * M_PI * (double)(r * r));            /* :295 float*float, then double */
            ovec3 P = odivs(lightColor, (float)samples);              /* :296 */
            ovec3 rDir = onormalize(osub(position, hit->position));   /* :298 */
            ovec3 nDir = onormalize(ld3(tri + 9));                    /* :300 */
            ovec3 B = odivs(P, A);                                    /* :301 */
            ovec3 D = oscale(B, omaxf(odot(rDir, nDir), 0.0f));       /* :304 */
            oisect j;
            j.position = ov(0, 0, 0);
            j.distance = FLT_MAX;                                     /* :308 */
            j.triangleIndex = -1;
            float unused;
            if (closest_intersection(position, oneg(rDir), tris15, T, &j, 1, &unused, 0.0f)) { /* :310 */
                if (j.distance < r * 0.99f) D = ov(0, 0, 0);          /* :313-314 */
            }
            result = oadd(result, D);                                 /* :319 */
        }
        result2 = oadd(result2, result);                              /* :322 (result not reset) */
    }
    return omul(result2, ld3(tri + 12));                              /* :325-326 */
}

/* ---- exported sub-stage entry points ------------------------------------- */
ORACLE_API int oracle_rt_closest_intersection(const float start[3], const float dir[3],
                                              const float* tris15, int T, b2r_intersection* io,
                                              int isLight, float dofFocal, float* focalOut) {
    oisect c;
    memcpy(&c, io, sizeof c);
    float slot = 0.0f;
    int hit = closest_intersection(ld3(start), ld3(dir), tris15, T, &c, isLight, &slot, dofFocal);
    memcpy(io, &c, sizeof c);
    if (focalOut) *focalOut = slot;
    return hit;
}

ORACLE_API void oracle_rt_direct_light(const b2r_intersection* hit, const float* tris15, int T,
                                       const b2r_frame_params* fp, float out[3]) {
    oisect c;
    memcpy(&c, hit, sizeof c);
    ovec3 r = direct_light(&c, tris15, T, fp);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* ---- Draw ------------------------------------------------------------------
 * Rows y0, y0+ystep, ... < y1 of a W x H frame (ystep > 1 only for bounded timing samples).
 * Output arrays are full-frame (may be NULL).
 * counters (may be NULL): [0] primary rays, [1] shadow rays. */
ORACLE_API int oracle_rt_draw(const float* tris15, int T, const b2r_frame_params* fp, int W, int H,
                              int y0, int y1, int ystep, float* pixelColours, b2r_intersection* closestOut,
                              float* focalDistances, int threads, unsigned long long* counters) {
    if (!tris15 || !fp || W <= 0 || H <= 0 || y0 < 0 || y1 > H || y0 > y1 || ystep < 1) return -1;
    const int N = fp->aaEnabled ? fp->aaSamples : 1; /* :551-554 */
    omat3 R;
    memcpy(&R, fp->cameraRot, sizeof R);
    const ovec3 cam = ld3(fp->cameraPos);
    const ovec3 indirect = ld3(fp->indirectLight);
    const int samples = fp->softShadowsEnabled ? fp->softShadowsSamples : 1;
    unsigned long long nPrimary = 0, nShadow = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1) reduction(+ : nPrimary, nShadow)
#endif
    for (int y = y0; y < y1; y += ystep) {
        float x1 = 0.0f, y1f = 0.0f;
        for (int x = 0; x < W; ++x) {
            oisect c; /* per-frame reset :335-339 (+P5) */
            c.position = ov(0, 0, 0);
            c.distance = FLT_MAX;
            c.triangleIndex = -1;
            float focal = 0.0f;
            ovec3 avg = ov(0, 0, 0);
            y1f = (N > 1) ? (float)y - 0.5f : (float)y;                 /* :564-567 */
            for (int z = 0; z < N; ++z) {
                x1 = (N > 1) ? (float)x - 0.5f : (float)x;              /* :571-574 */
                for (int z2 = 0; z2 < N; ++z2) {
                    ovec3 d = ov(x1 - (float)W / 2.0f, y1f - (float)H / 2.0f, fp->focalLength); /* :579 */
                    ovec3 dir = omat_vec(&R, d);
                    ++nPrimary;
                    if (closest_intersection(cam, dir, tris15, T, &c, 0, &focal, fp->dofFocalLength)) { /* :580 */
                        ovec3 D = direct_light(&c, tris15, T, fp);      /* :583 */
                        nShadow += (unsigned long long)fp->numLights * (unsigned long long)samples;
                        ovec3 Tsum = oadd(D, indirect);                 /* :586 */
                        ovec3 p = ld3(tris15 + 15 * c.triangleIndex + 12); /* :587 */
                        avg = oadd(avg, omul(p, Tsum));                 /* :588-591 */
                        x1 += (1.0f / (float)(N - 1));                  /* :593 only advances on a hit */
                    }
                }
                y1f += (1.0f / (float)(N - 1));                         /* :596 */
            }
            avg = odivs(avg, (float)(N * N));                           /* :599 */
            size_t idx = (size_t)y * (size_t)W + (size_t)x;             /* P2 */
            if (pixelColours) {
                pixelColours[3 * idx] = avg.x;
                pixelColours[3 * idx + 1] = avg.y;
                pixelColours[3 * idx + 2] = avg.z;
            }
            if (closestOut) memcpy(&closestOut[idx], &c, sizeof c);
            if (focalDistances) focalDistances[idx] = focal;
        }
    }
    if (counters) {
        counters[0] = nPrimary;
        counters[1] = nShadow;
    }
    return 0;
}
