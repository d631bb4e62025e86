/* oracle_resolve.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the frame tail shared by both reference programs:
 *   CalculateDOF   raytracer.cpp:608-656 == rasteriser.cpp:484-529
 *   PutPixelSDL    SDLauxiliary.h:70-81
 * plus the 24-bit bottom-up BMP payload SDL_SaveBMP produces (raytracer.cpp:175).
 *
 * Defined policy for the reference's out-of-range reads in the DOF branch
 * (raytracer.cpp:637 has no bounds checks): the flattened index
 * (y+z)*W+(x+z2) is used as-is when it lies inside [0, W*H) -- so columns wrap
 * into the neighbouring row exactly like the reference -- and contributes 0
 * when it falls outside the array (where the reference reads unrelated
 * memory: undefined behaviour).
 *
 * Build: gcc -O2 -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/b2r.h"

#define ORACLE_API __attribute__((visibility("default")))

static inline float clamp255(float c) {
    float v = 255 * c;
    v = (v < 0.f) ? 0.f : v;     /* glm::max(x, 0) */
    v = (255.f < v) ? 255.f : v; /* glm::min(.., 255) */
    return v;
}

static inline uint32_t put_pixel(float r, float g, float b) { /* SDLauxiliary.h:75-80 */
    uint8_t R = (uint8_t)clamp255(r), G = (uint8_t)clamp255(g), B = (uint8_t)clamp255(b);
    return ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B;
}

ORACLE_API int oracle_resolve_surface(const float* pixelColours, const float* focalDistances, int W, int H,
                                      int dofEnabled, int dofKernelSize, uint32_t* surface) {
    if (!pixelColours || !surface || W <= 0 || H <= 0) return -1;
    memset(surface, 0, (size_t)W * H * 4); /* border is never written: stays black */
    const float totalPixels = (float)(dofKernelSize * dofKernelSize);
    const int zlo = (int)ceilf(dofKernelSize / -2.0f), zhi = (int)ceilf(dofKernelSize / 2.0f);
    const long long n = (long long)W * H;
    for (int y = 1; y < H - 1; ++y) {
        for (int x = 1; x < W - 1; ++x) {
            float fr = 0.0f, fg = 0.0f, fb = 0.0f;
            const long long c = (long long)y * W + x;
            if (dofEnabled) {
                for (int z = zlo; z < zhi; ++z) {
                    for (int z2 = zlo; z2 < zhi; ++z2) {
                        float a = fabsf(focalDistances[c]);
                        float m = (1.0f < a) ? 1.0f : a; /* std::min(abs(fd), 1.0f) */
                        float wgt;
                        if (z == 0 && z2 == 0)
                            wgt = 1 - (m * ((totalPixels - 1) / totalPixels));
                        else
                            wgt = m * (1.0f / totalPixels);
                        long long q = (long long)(y + z) * W + (x + z2);
                        if (q < 0 || q >= n) continue; /* policy: see header */
                        fr += pixelColours[3 * q] * wgt;
                        fg += pixelColours[3 * q + 1] * wgt;
                        fb += pixelColours[3 * q + 2] * wgt;
                    }
                }
            } else {
                fr = pixelColours[3 * c];
                fg = pixelColours[3 * c + 1];
                fb = pixelColours[3 * c + 2];
            }
            surface[c] = put_pixel(fr, fg, fb);
        }
    }
    return 0;
}

ORACLE_API size_t oracle_bmp_payload_bytes(int W, int H) { return (size_t)((W * 3 + 3) & ~3) * (size_t)H; }

/* XRGB surface -> bottom-up BGR rows, each padded to 4 bytes. */
ORACLE_API int oracle_surface_to_bgr8(const uint32_t* surface, int W, int H, uint8_t* bgr) {
    const size_t pitch = (size_t)((W * 3 + 3) & ~3);
    memset(bgr, 0, pitch * H);
    for (int y = 0; y < H; ++y) {
        uint8_t* row = bgr + pitch * (size_t)(H - 1 - y);
        for (int x = 0; x < W; ++x) {
            uint32_t p = surface[(size_t)y * W + x];
            row[3 * x] = (uint8_t)(p & 0xFF);
            row[3 * x + 1] = (uint8_t)((p >> 8) & 0xFF);
            row[3 * x + 2] = (uint8_t)((p >> 16) & 0xFF);
        }
    }
    return 0;
}
