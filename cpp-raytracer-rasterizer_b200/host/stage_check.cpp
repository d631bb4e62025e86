// stage_check.cpp -- the reference's Draw() loops re-assembled from the stage functions of the drop-ins
// (ClosestIntersection + DirectLight; VertexShader + ComputePolygonRows + Interpolate + PixelShader), compared with
// the fused Draw() on a tiny screen.  Shows that code written against the reference's stage signatures keeps
// working and agrees bit for bit with the frame kernels.  Exit code 0 = all equal.
#include <cstdio>
#include <cstring>

#include "rasteriser_dropin.h"
#include "raytracer_dropin.h"

static int check_raytracer() {
    using namespace rtref;
    const int W = 24, H = 16;
    if (Initialize(W, H, 0)) return 1;
    Update();
    Draw();
    std::vector<vec3> fused = pixelColours;
    std::vector<Intersection> fusedClosest = closestIntersections;
    int bad = 0;
    for (int y = 0; y < H; ++y)          // raytracer.cpp:558-603 with realSamples == 1
        for (int x = 0; x < W; ++x) {
            Intersection c;
            c.position = vec3(0, 0, 0);
            c.distance = 3.402823466e+38f;
            c.triangleIndex = -1;
            vec3 d(x - (float)W / 2.0f, y - (float)H / 2.0f, focalLength);
            vec3 dir(cameraRot[0][0] * d.x + cameraRot[1][0] * d.y + cameraRot[2][0] * d.z,
                     cameraRot[0][1] * d.x + cameraRot[1][1] * d.y + cameraRot[2][1] * d.z,
                     cameraRot[0][2] * d.x + cameraRot[1][2] * d.y + cameraRot[2][2] * d.z);
            vec3 colour(0, 0, 0);
            if (ClosestIntersection(cameraPos, dir, triangles, c, false, x, y)) {
                vec3 D = DirectLight(c);
                vec3 p = triangles[c.triangleIndex].color;
                colour = vec3(p.x * (D.x + indirectLight.x), p.y * (D.y + indirectLight.y), p.z * (D.z + indirectLight.z));
            }
            const size_t i = (size_t)y * W + x;
            if (std::memcmp(&colour, &fused[i], 12) != 0 || std::memcmp(&c, &fusedClosest[i], sizeof c) != 0) ++bad;
        }
    Shutdown();
    std::printf("raytracer stages vs Draw(): %d of %d pixels differ\n", bad, W * H);
    return bad;
}

static int check_rasteriser() {
    using namespace raref;
    const int W = 32, H = 24;
    if (Initialize(W, H, 0)) return 1;
    Update();
    Draw();
    std::vector<vec3> fused = pixelColours;
    std::vector<float> fusedDepth = depthBuffer, fusedFocal = focalDistances;
    // rasteriser.cpp:461-479 + DrawPolygon/DrawRows/DrawLineSDL/Bresenham on the host, stages on the GPU
    std::vector<float> depth((size_t)W * H, 0.0f);
    pixelColours.assign((size_t)W * H, vec3(0, 0, 0));
    focalDistances.assign((size_t)W * H, 0.0f);
    for (size_t t = 0; t < triangles.size(); ++t) {
        if (triangles[t].isCulled) continue;
        std::vector<Pixel> vp(3), left, right;
        Vertex v;
        v.position = triangles[t].v0; VertexShader(v, vp[0]);
        v.position = triangles[t].v1; VertexShader(v, vp[1]);
        v.position = triangles[t].v2; VertexShader(v, vp[2]);
        ComputePolygonRows(vp, left, right);
        for (size_t r = 0; r < left.size(); ++r) {
            const Pixel a = left[r], b = right[r];
            if ((a.y >= H && b.y >= H) || (a.y < 0 && b.y < 0)) continue;
            const int dx = b.x - a.x;
            for (int i = 0; i < dx; ++i) {  // Bresenham with dy == 0 (:651-670)
                Pixel q;
                q.x = a.x + 1 + i;
                q.y = a.y;
                if (q.x < 0 || q.x >= W || q.y < 0 || q.y >= H) continue;
                const float zs = (b.zinv - a.zinv) / float(dx);
                q.zinv = a.zinv + zs * float(i);
                q.pos3d = vec3(a.pos3d.x + ((b.pos3d.x - a.pos3d.x) / float(dx)) * float(i),
                               a.pos3d.y + ((b.pos3d.y - a.pos3d.y) / float(dx)) * float(i),
                               a.pos3d.z + ((b.pos3d.z - a.pos3d.z) / float(dx)) * float(i));
                if (q.zinv > depth[(size_t)q.y * W + q.x]) {
                    depth[(size_t)q.y * W + q.x] = q.zinv;
                    PixelShader(q, triangles[t].color, triangles[t].normal);
                }
            }
        }
    }
    int bad = 0;
    for (size_t i = 0; i < (size_t)W * H; ++i)
        if (std::memcmp(&pixelColours[i], &fused[i], 12) != 0 || std::memcmp(&depth[i], &fusedDepth[i], 4) != 0 ||
            std::memcmp(&focalDistances[i], &fusedFocal[i], 4) != 0)
            ++bad;
    Shutdown();
    std::printf("rasteriser stages vs Draw(): %d of %d pixels differ\n", bad, W * H);
    return bad;
}

int main() {
    int bad = check_raytracer();
    bad += check_rasteriser();
    return bad ? 1 : 0;
}
