// pixel_pack.cuh -- PutPixelSDL (SDLauxiliary.h:70-81): Uint8(glm::clamp(255*c, 0, 255)) per channel, truncation,
// packed as XRGB8888.  Shared by the resolve kernels and the raytracer's fused trace+resolve path.
#pragma once
#include <stdint.h>

#include "exact.cuh"

namespace b2r {

__device__ __forceinline__ uint32_t quantise(float c) {
    float v = xmul(255.0f, c);        // 255*color.r
    v = (v < 0.f) ? 0.f : v;          // glm::clamp = min(max(x, 0), 255)
    v = (255.f < v) ? 255.f : v;
    return __float2uint_rz(v) & 0xFFu;  // Uint8(...)
}

// SDL_MapRGB on the XRGB8888 surface
__device__ __forceinline__ uint32_t pack_xrgb(float r, float g, float b) {
    return (quantise(r) << 16) | (quantise(g) << 8) | quantise(b);
}

// CalculateDOF's loops run over [1, W-1) x [1, H-1) (raytracer.cpp:616-617): the 1-pixel border is never written.
__device__ __forceinline__ bool inside_border(int x, int y, int W, int H) {
    return x >= 1 && x < W - 1 && y >= 1 && y < H - 1;
}

}  // namespace b2r
