/* oracle_ras.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference rasteriser's per-triangle /
 * per-pixel hot path, plain C, runtime screen size.  Checker for the CUDA
 * path; itself pinned against the compiled reference (oracle/_ref) and
 * tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call it.
 *
 * Follows (all citations: rasteriser/Source/rasteriser.cpp unless noted):
 *   oracle_ras_vertex_shader         VertexShader          :532-546
 *   oracle_ras_interpolate           Interpolate           :615-637 (+ fPixel, TestModel.h:98-127)
 *   oracle_ras_compute_polygon_rows  ComputePolygonRows    :674-735
 *   draw_row                         DrawLineSDL+Bresenham :592-612, :639-672
 *   oracle_ras_pixel_shader          PixelShader           :549-589
 *   oracle_ras_draw                  Draw + DrawPolygon + DrawRows  :461-479, :738-768, clears :183-192
 *   oracle_ras_cull                  culling block of Update()      :385-447, InCuboid :451-458
 * Deviations, each mechanical: P2 row stride = screen
 * width; P3 line[] entries Bresenham leaves unwritten are skipped (the
 * reference reads them uninitialised); P4 the depth winner's triangle index
 * is recorded.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march, no -ffast-math).
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b2r.h"
#include "oracle_math.h"

#define ORACLE_API __attribute__((visibility("default")))

typedef struct {
    int32_t x, y;
    float zinv;
    ovec3 pos3d;
} opixel; /* == struct Pixel, TestModel.h:34-53 (24 bytes) */

static inline ovec3 ld3(const float* p) { return ov(p[0], p[1], p[2]); }

typedef struct {
    int W, H;
    ovec3 cam;
    omat3 R, Rinv;
    float focal;
    const b2r_frame_params* fp;
} ocam;

static void cam_init(ocam* c, const b2r_frame_params* fp, int W, int H) {
    c->W = W;
    c->H = H;
    c->cam = ld3(fp->cameraPos);
    memcpy(&c->R, fp->cameraRot, sizeof c->R);
    c->Rinv = oinverse(&c->R); /* :559 recomputed per pixel in the reference; same bits every time */
    c->focal = fp->focalLength;
    c->fp = fp;
}

static opixel vertex_shader(const ocam* c, ovec3 v) {
    opixel p;
    ovec3 pos = ovec_mat(osub(v, c->cam), &c->R);            /* :535 */
    p.pos3d = odivs(pos, pos.z);                             /* :538 */
    p.zinv = 1.0f / pos.z;                                   /* :541 */
    p.x = (int32_t)((float)(int)(c->focal * (pos.x * p.zinv)) + ((float)c->W / 2.0f)); /* :544 */
    p.y = (int32_t)((float)(int)(c->focal * (pos.y * p.zinv)) + ((float)c->H / 2.0f)); /* :545 */
    return p;
}

/* N samples from a to b; every attribute advances by repeated float addition. */
static void interpolate(opixel a, opixel b, opixel* out, int N) {
    float div = (float)((N - 1 > 1) ? N - 1 : 1);            /* :622 */
    float sx = (float)(b.x - a.x) / div;                     /* Pixel operator- then fPixel operator/ */
    float sy = (float)(b.y - a.y) / div;
    float sz = (b.zinv - a.zinv) / div;
    ovec3 sp = odivs(osub(b.pos3d, a.pos3d), div);
    float cx = (float)a.x, cy = (float)a.y, cz = a.zinv;     /* fPixel(Pixel&) */
    ovec3 cp = a.pos3d;
    for (int i = 0; i < N; ++i) {                            /* :626-636 */
        out[i].x = (int32_t)cx;
        out[i].y = (int32_t)cy;
        out[i].zinv = cz;
        out[i].pos3d = cp;
        cx += sx;
        cy += sy;
        cz += sz;
        cp = oadd(cp, sp);
    }
}

typedef struct {
    opixel* left;
    opixel* right;
    opixel* edge;
    int cap;
} orows;

static int rows_reserve(orows* r, int rows) {
    if (rows <= r->cap) return 0;
    int cap = rows + rows / 2 + 16;
    opixel* l = (opixel*)realloc(r->left, sizeof(opixel) * cap);
    if (l) r->left = l;
    opixel* rr = (opixel*)realloc(r->right, sizeof(opixel) * cap);
    if (rr) r->right = rr;
    opixel* e = (opixel*)realloc(r->edge, sizeof(opixel) * cap);
    if (e) r->edge = e;
    if (!l || !rr || !e) return -1;
    r->cap = cap;
    return 0;
}

/* Returns ROWS (>=1) or -1 when out of memory. */
static int compute_polygon_rows(const opixel vp[3], orows* r) {
    int maxY = vp[0].y > vp[1].y ? vp[0].y : vp[1].y;
    if (vp[2].y > maxY) maxY = vp[2].y;
    int minY = vp[0].y < vp[1].y ? vp[0].y : vp[1].y;
    if (vp[2].y < minY) minY = vp[2].y;
    long long rowsLL = (long long)maxY - (long long)minY + 1; /* :682 */
    if (rowsLL > (1 << 24)) return -1;
    int ROWS = (int)rowsLL;
    if (rows_reserve(r, ROWS) != 0) return -1;
    memset(r->left, 0, sizeof(opixel) * ROWS);
    memset(r->right, 0, sizeof(opixel) * ROWS);
    for (int i = 0; i < ROWS; ++i) {                         /* :694-698 */
        r->left[i].x = +INT_MAX;
        r->right[i].x = -INT_MAX;
    }
    for (int i = 0; i < 3; ++i) {                            /* :705-734 */
        int j = (i + 1) % 3;
        opixel a = vp[i], b = vp[j];
        a.y -= minY;                                         /* :709-710 */
        b.y -= minY;
        int n = abs(vp[i].y - vp[j].y) + 1;                  /* :712 */
        interpolate(a, b, r->edge, n);
        for (int k = 0; k < n; ++k) {
            const opixel* e = &r->edge[k];
            if (e->x < r->left[e->y].x) {                    /* :718 strict */
                r->left[e->y] = *e;
                r->left[e->y].y = e->y + minY;
            }
            if (e->x > r->right[e->y].x) {                   /* :726 strict */
                r->right[e->y] = *e;
                r->right[e->y].y = e->y + minY;
            }
        }
    }
    return ROWS;
}

static void pixel_shader(const ocam* c, const opixel* p, ovec3 color, ovec3 normal, float* colourOut,
                         float* focalOut) {
    const b2r_frame_params* fp = c->fp;
    ovec3 P = odivs(p->pos3d, p->zinv);                      /* :557 */
    P = ovec_mat(P, &c->Rinv);                               /* :559 */
    P = oadd(P, c->cam);                                     /* :560 */
    ovec3 result = ov(0, 0, 0);
    float distance = odistance(P, c->cam);                   /* :564 */
    *focalOut = distance - fp->dofFocalLength;               /* :565 */
    for (int i = 0; i < fp->numLights; ++i) {                /* :567-584 */
        ovec3 L = ld3(fp->lights[i].position);
        float r = odistance(P, L);
        float A = (float)(4 * M_PI * (double)(r * r));       /* :576 */
        ovec3 lightColor = oscale(ld3(fp->lights[i].color), fp->lights[i].intensity);
        ovec3 rDir = onormalize(osub(L, P));
        ovec3 B = odivs(lightColor, A);
        ovec3 D = oscale(B, omaxf(odot(rDir, normal), 0.0f)); /* normal NOT re-normalised, :579 */
        result = oadd(result, D);
    }
    /* :587  currentReflectance * (result + indirect) * color, left to right */
    ovec3 out = omul(omul(ld3(fp->currentReflectance), oadd(result, ld3(fp->indirectLight))), color);
    colourOut[0] = out.x;
    colourOut[1] = out.y;
    colourOut[2] = out.z;
}

typedef struct {
    float* depth;
    float* colours;
    float* focal;
    int32_t* winner;
    unsigned long long rows, tests, passes, tris;
} oframe;

/* One row: DrawLineSDL (:592-612) around Bresenham (:639-672) with dy == 0. */
static void draw_row(const ocam* c, oframe* f, const opixel* a, const opixel* b, ovec3 color, ovec3 normal,
                     int tri) {
    int dx = b->x - a->x;                                    /* :598 */
    float zstep = (b->zinv - a->zinv) / (float)dx;           /* :648 */
    ovec3 pstep = odivs(osub(b->pos3d, a->pos3d), (float)dx);/* :649 */
    int y = a->y; /* dy2 == 0 and d == -dx < 0 for every step, so y never advances (:654-662) */
    for (int i = 0; i < dx; ++i) {
        int x = a->x + 1 + i;                                /* :653 */
        if (!(x >= 0 && x < c->W)) continue;                 /* :663 (+P3) */
        if (!(y < c->H && y >= 0)) continue;                 /* :606 */
        opixel q;
        q.x = x;
        q.y = y;
        q.zinv = a->zinv + zstep * (float)i;                 /* :667 */
        q.pos3d = oadd(a->pos3d, oscale(pstep, (float)i));   /* :668 */
        size_t idx = (size_t)y * (size_t)c->W + (size_t)x;
        ++f->tests;
        if (q.zinv > f->depth[idx]) {                        /* :606 strict */
            f->depth[idx] = q.zinv;                          /* :608 */
            f->winner[idx] = tri;                            /* P4 */
            ++f->passes;
            pixel_shader(c, &q, color, normal, f->colours + 3 * idx, f->focal + idx); /* :609 */
        }
    }
}

/* ---- exported sub-stage entry points ------------------------------------- */
ORACLE_API void oracle_ras_vertex_shader(const b2r_frame_params* fp, int W, int H, const float v[3],
                                         void* outPixel24) {
    ocam c;
    cam_init(&c, fp, W, H);
    opixel p = vertex_shader(&c, ld3(v));
    memcpy(outPixel24, &p, sizeof p);
}

ORACLE_API void oracle_ras_interpolate(const void* a24, const void* b24, void* out, int n) {
    opixel a, b;
    memcpy(&a, a24, sizeof a);
    memcpy(&b, b24, sizeof b);
    interpolate(a, b, (opixel*)out, n);
}

ORACLE_API int oracle_ras_compute_polygon_rows(const void* vp3x24, void* outLeft, void* outRight, int maxRows) {
    opixel vp[3];
    memcpy(vp, vp3x24, sizeof vp);
    orows r = {0, 0, 0, 0};
    int rows = compute_polygon_rows(vp, &r);
    if (rows > 0 && rows <= maxRows) {
        memcpy(outLeft, r.left, sizeof(opixel) * rows);
        memcpy(outRight, r.right, sizeof(opixel) * rows);
    }
    free(r.left);
    free(r.right);
    free(r.edge);
    return rows;
}

ORACLE_API void oracle_ras_pixel_shader(const b2r_frame_params* fp, int W, int H, const void* p24,
                                        const float color[3], const float normal[3], float outColour[3],
                                        float* outFocal) {
    ocam c;
    cam_init(&c, fp, W, H);
    opixel p;
    memcpy(&p, p24, sizeof p);
    pixel_shader(&c, &p, ld3(color), ld3(normal), outColour, outFocal);
}

/* ---- culling block of Update() (:385-447) -------------------------------- */
ORACLE_API int oracle_ras_cull(const float* tris15, int T, const b2r_frame_params* fp, int W, int H,
                               uint8_t* culled) {
    ocam c;
    cam_init(&c, fp, W, H);
    ovec3 fVec = onormalize(ovec_mat(ov(0, 0, 1.0f), &c.R));               /* :385 */
    float nearZ = c.cam.z + fVec.z * 0.1f, farZ = c.cam.z + fVec.z * 15.0f; /* :386 */
    float w = (float)W, h = (float)H;
    ovec3 t = ov(0.0f, -h / 2.0f, c.focal), b = ov(0.0f, h / 2.0f, c.focal); /* :392-393 */
    float cy = odot(t, b) / (olength(t) * olength(b));                     /* :394 */
    float rfovy = acosf(cy);                                               /* :395 */
    float aspect = w / h;
    float m00 = (1.0f / tanf(rfovy / 2.0f)) / aspect;                      /* :398 */
    float m11 = (1.0f / tanf(rfovy / 2.0f));                               /* :399 */
    float m22 = farZ / (farZ - nearZ);                                     /* :400 */
    float m32 = 1.0f;                                                      /* :401-402 */
    for (int i = 0; i < T; ++i) {
        const float* tr = tris15 + 15 * i;
        int cull = 0;
        if (fp->backfaceCulling) {                                         /* :408-414 */
            if (odot(osub(ld3(tr), c.cam), ld3(tr + 9)) > 0.0f) cull = 1;
        }
        if (fp->frustumCulling && !cull) {                                 /* :416-446 */
            int inside[3];
            for (int k = 0; k < 3; ++k) {
                ovec3 v = ovec_mat(osub(ld3(tr + 3 * k), c.cam), &c.R);    /* :423-425 */
                /* vec4(v,1) * transform: glm/detail/type_mat4x4.inl:664-675, all 16 terms */
                float v3 = 1.0f;
                float X = m00 * v.x + 0.0f * v.y + 0.0f * v.z + 0.0f * v3;
                float Y = 0.0f * v.x + m11 * v.y + 0.0f * v.z + 0.0f * v3;
                float Z = 0.0f * v.x + 0.0f * v.y + m22 * v.z + 0.0f * v3;
                float Wc = 0.0f * v.x + 0.0f * v.y + m32 * v.z + 0.0f * v3;
                X = X / Wc;                                                /* :435-437 */
                Y = Y / Wc;
                Z = Z / Wc;
                inside[k] = (X >= -1.0f && X <= 1.0f && Y >= -1.0f && Y <= 1.0f && Z >= 0.0f && Z <= 1.0f); /* :453 */
            }
            if (!inside[0] && !inside[1] && !inside[2]) cull = 1;          /* :444-445 */
        }
        culled[i] = (uint8_t)cull;
    }
    return 0;
}

/* ---- Draw ------------------------------------------------------------------
 * Whole frame (the reference draws triangles serially in index order, :468).
 * culled may be NULL (nothing culled).  Outputs are full-frame, may be NULL.
 * counters (may be NULL): [0] triangles drawn, [1] rows, [2] on-screen depth
 * tests, [3] depth passes. */
ORACLE_API int oracle_ras_draw(const float* tris15, const uint8_t* culled, int T, const b2r_frame_params* fp,
                               int W, int H, float* depthBuffer, float* pixelColours, float* focalDistances,
                               int32_t* winnerIndex, unsigned long long* counters) {
    if (!tris15 || !fp || W <= 0 || H <= 0) return -1;
    size_t n = (size_t)W * (size_t)H;
    oframe f;
    memset(&f, 0, sizeof f);
    f.depth = (float*)calloc(n, sizeof(float));              /* clears: :183-192 */
    f.colours = (float*)calloc(n * 3, sizeof(float));
    f.focal = (float*)calloc(n, sizeof(float));
    f.winner = (int32_t*)malloc(n * sizeof(int32_t));
    if (!f.depth || !f.colours || !f.focal || !f.winner) return -2;
    for (size_t i = 0; i < n; ++i) f.winner[i] = -1;
    ocam c;
    cam_init(&c, fp, W, H);
    orows r = {0, 0, 0, 0};
    int rc = 0;
    for (int i = 0; i < T && rc == 0; ++i) {                 /* :468 */
        if (culled && culled[i]) continue;                   /* :470 */
        const float* tr = tris15 + 15 * i;
        opixel vp[3];
        for (int k = 0; k < 3; ++k) vp[k] = vertex_shader(&c, ld3(tr + 3 * k)); /* :760-761 */
        int ROWS = compute_polygon_rows(vp, &r);             /* :766 */
        if (ROWS < 0) {
            rc = -3;
            break;
        }
        ++f.tris;
        f.rows += (unsigned long long)ROWS;
        ovec3 color = ld3(tr + 12), normal = ld3(tr + 9);
        for (int k = 0; k < ROWS; ++k) {                     /* DrawRows :740-752 */
            const opixel* a = &r.left[k];
            const opixel* b = &r.right[k];
            if ((a->y >= H && b->y >= H) || (a->y < 0 && b->y < 0)) continue; /* :743 */
            draw_row(&c, &f, a, b, color, normal, i);
        }
    }
    if (rc == 0) {
        if (depthBuffer) memcpy(depthBuffer, f.depth, n * sizeof(float));
        if (pixelColours) memcpy(pixelColours, f.colours, n * 3 * sizeof(float));
        if (focalDistances) memcpy(focalDistances, f.focal, n * sizeof(float));
        if (winnerIndex) memcpy(winnerIndex, f.winner, n * sizeof(int32_t));
        if (counters) {
            counters[0] = f.tris;
            counters[1] = f.rows;
            counters[2] = f.tests;
            counters[3] = f.passes;
        }
    }
    free(f.depth);
    free(f.colours);
    free(f.focal);
    free(f.winner);
    free(r.left);
    free(r.right);
    free(r.edge);
    return rc;
}
