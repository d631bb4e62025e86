// scenes.cpp -- host-side scene/camera helpers of the C ABI (no GPU involved).
//
// The Cornell box is the input fixture of every BASELINE config; it is the
// reference's LoadTestModel (raytracer TestModel.h:51-192, identical in
// rasteriser TestModel.h:151-292) restated as a vertex/index table.  All
// arithmetic is the reference's, in its order (float scale 2/L, subtract 1,
// negate x and y, normal = normalize(cross(e2, e1))), so the triangle bytes
// are identical (tests pin the FNV-1a32 fingerprint b715a8a2 from SURVEY.md).
// Compiled with -ffp-contract=off.
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b2r.h"

namespace {

struct F3 {
    float x, y, z;
};
inline F3 sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 cross(F3 a, F3 b) {  // glm::cross, func_geometric.inl:133-142
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline F3 normalize(F3 v) {  // v * (1/sqrt(dot(v,v)))
    float d = v.x * v.x + v.y * v.y + v.z * v.z;
    float inv = 1.0f / sqrtf(d);
    return {v.x * inv, v.y * inv, v.z * inv};
}

// Triangle ctor + ComputeNormal (TestModel.h:20-31): normal from e2 x e1.
void write_triangle(unsigned char* dst, int stride, F3 v0, F3 v1, F3 v2, F3 color) {
    F3 e1 = sub(v1, v0), e2 = sub(v2, v0);
    F3 n = normalize(cross(e2, e1));
    float rec[15] = {v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z, n.x, n.y, n.z, color.x, color.y, color.z};
    memcpy(dst, rec, sizeof rec);
    if (stride == 64) memset(dst + 60, 0, 4);  // isCulled = false (+ padding)
}

const float kL = 555.0f;  // TestModel.h:72
// corner order A..H
const float kRoom[8][3] = {{kL, 0, 0}, {0, 0, 0}, {kL, 0, kL}, {0, 0, kL}, {kL, kL, 0}, {0, kL, 0}, {kL, kL, kL}, {0, kL, kL}};
const float kShort[8][3] = {{290, 0, 114}, {130, 0, 65}, {240, 0, 272}, {82, 0, 225},
                            {290, 165, 114}, {130, 165, 65}, {240, 165, 272}, {82, 165, 225}};
const float kTall[8][3] = {{423, 0, 247}, {265, 0, 296}, {472, 0, 406}, {314, 0, 456},
                           {423, 330, 247}, {265, 330, 296}, {472, 330, 406}, {314, 330, 456}};
enum { A, B, C, D, E, F, G, H };
// room faces (TestModel.h:84-102) with their colours
const int kRoomFaces[10][3] = {{C, B, A}, {C, D, B}, {A, E, C}, {C, E, G}, {F, B, D},
                               {H, F, D}, {E, F, G}, {F, H, G}, {G, D, C}, {G, H, D}};
const float kRoomColour[5][3] = {{0.15f, 0.75f, 0.15f} /*green floor*/,  {0.75f, 0.15f, 0.75f} /*purple*/,
                                 {0.75f, 0.75f, 0.15f} /*yellow*/,       {0.15f, 0.75f, 0.75f} /*cyan ceiling*/,
                                 {0.75f, 0.75f, 0.75f} /*white back*/};
// block faces (TestModel.h:117-135, 152-170)
const int kBlockFaces[10][3] = {{E, B, A}, {E, F, B}, {F, D, B}, {F, H, D}, {H, C, D},
                                {H, G, C}, {G, E, C}, {E, A, C}, {G, F, E}, {G, H, F}};
const float kRed[3] = {0.75f, 0.15f, 0.15f}, kBlue[3] = {0.15f, 0.15f, 0.75f};

F3 to_unit_box(const float* p) {
    // *= 2/L ; -= (1,1,1) ; x *= -1 ; y *= -1   (TestModel.h:172-188)
    const float s = 2 / kL;
    F3 v = {p[0] * s, p[1] * s, p[2] * s};
    v.x -= 1.0f; v.y -= 1.0f; v.z -= 1.0f;
    v.x *= -1.0f;
    v.y *= -1.0f;
    return v;
}

}  // namespace

extern "C" {

int b2r_scene_cornell_box(void* out, int capacity, int stride_bytes) {
    if (!out || capacity < 30 || (stride_bytes != 60 && stride_bytes != 64)) return B2R_E_INVALID;
    unsigned char* dst = (unsigned char*)out;
    int n = 0;
    for (int f = 0; f < 10; ++f, ++n) {
        const float* c = kRoomColour[f / 2];
        write_triangle(dst + (size_t)n * stride_bytes, stride_bytes, to_unit_box(kRoom[kRoomFaces[f][0]]),
                       to_unit_box(kRoom[kRoomFaces[f][1]]), to_unit_box(kRoom[kRoomFaces[f][2]]), {c[0], c[1], c[2]});
    }
    for (int blk = 0; blk < 2; ++blk) {
        const float(*corner)[3] = blk ? kTall : kShort;
        const float* c = blk ? kBlue : kRed;
        for (int f = 0; f < 10; ++f, ++n)
            write_triangle(dst + (size_t)n * stride_bytes, stride_bytes, to_unit_box(corner[kBlockFaces[f][0]]),
                           to_unit_box(corner[kBlockFaces[f][1]]), to_unit_box(corner[kBlockFaces[f][2]]),
                           {c[0], c[1], c[2]});
    }
    return n;
}

long long b2r_scene_tessellate(const void* in, int count, int in_stride, int k, void* out, int out_stride) {
    if (!in || count < 0 || k < 1 || (in_stride != 60 && in_stride != 64)) return B2R_E_INVALID;
    const long long total = (long long)count * k * k;
    if (!out) return total;
    if (out_stride != 60 && out_stride != 64) return B2R_E_INVALID;
    const unsigned char* src = (const unsigned char*)in;
    unsigned char* dst = (unsigned char*)out;
    long long n = 0;
    const float fk = (float)k;
    for (int t = 0; t < count; ++t) {
        float rec[15];
        memcpy(rec, src + (size_t)t * in_stride, sizeof rec);
        const F3 a = {rec[0], rec[1], rec[2]}, b = {rec[3], rec[4], rec[5]}, c = {rec[6], rec[7], rec[8]};
        const F3 col = {rec[12], rec[13], rec[14]};
        const F3 ab = sub(b, a), ac = sub(c, a);
        auto P = [&](int i, int j) -> F3 {  // A + (i/k)(B-A) + (j/k)(C-A), float, left to right
            const float fi = (float)i / fk, fj = (float)j / fk;
            return {a.x + fi * ab.x + fj * ac.x, a.y + fi * ab.y + fj * ac.y, a.z + fi * ab.z + fj * ac.z};
        };
        for (int j = 0; j < k; ++j)
            for (int i = 0; i + j < k; ++i) {
                write_triangle(dst + (size_t)n++ * out_stride, out_stride, P(i, j), P(i + 1, j), P(i, j + 1), col);
                if (i + j < k - 1)
                    write_triangle(dst + (size_t)n++ * out_stride, out_stride, P(i + 1, j), P(i + 1, j + 1),
                                   P(i, j + 1), col);
            }
    }
    return n;
}

// ASCII STL ingestion with the reference loader's exact semantics (rasteriser/Source/LoadSTL.cpp:17-81): every line
// containing "outer" is followed by three vertex lines, split on single spaces with empty tokens and the word
// "vertex" dropped, each coordinate (float)atof(token); every triangle gets colour (0.5,0.5,0.5); afterwards all
// coordinates are multiplied by -0.05f and the normal recomputed like the Triangle ctor.
long long b2r_scene_load_stl(const char* path, void* out, long long capacity, int stride_bytes) {
    if (!path || (out && stride_bytes != 60 && stride_bytes != 64)) return B2R_E_INVALID;
    FILE* fp = fopen(path, "rb");
    if (!fp) return B2R_E_IO;
    long long n = 0;
    char* line = nullptr;
    size_t cap = 0;
    const float scale = 0.05f;
    auto read_vertex = [&](F3* v) -> bool {
        if (getline(&line, &cap, fp) < 0) return false;
        float c[3] = {0.f, 0.f, 0.f};
        int got = 0;
        for (char* tok = line; *tok && got < 3;) {
            while (*tok == ' ') ++tok;          // delimiters (getline(ss, tok, ' ') yields empty tokens, which are dropped)
            if (!*tok || *tok == '\n') break;
            char* end = tok;
            while (*end && *end != ' ' && *end != '\n') ++end;
            const bool isWord = (end - tok == 6) && strncmp(tok, "vertex", 6) == 0;
            if (!isWord) {
                char save = *end;
                *end = 0;
                c[got++] = (float)atof(tok);
                *end = save;
            }
            tok = end;
        }
        *v = {c[0], c[1], c[2]};
        return true;
    };
    while (getline(&line, &cap, fp) >= 0) {
        if (!strstr(line, "outer")) continue;
        F3 v[3];
        bool ok = true;
        for (int k = 0; k < 3 && ok; ++k) ok = read_vertex(&v[k]);
        if (!ok) break;
        if (out) {
            if (n >= capacity) {
                free(line);
                fclose(fp);
                return B2R_E_CAPACITY;
            }
            for (int k = 0; k < 3; ++k) {  // x, z, y each *= -scale (LoadSTL.cpp:61-73)
                v[k].x *= -scale;
                v[k].z *= -scale;
                v[k].y *= -scale;
            }
            write_triangle((unsigned char*)out + (size_t)n * stride_bytes, stride_bytes, v[0], v[1], v[2], {0.5f, 0.5f, 0.5f});
        }
        ++n;
    }
    free(line);
    fclose(fp);
    return n;
}

int b2r_camera_rot_from_yaw(float yaw, float rot11, float* r) {
    if (!r) return B2R_E_INVALID;
    for (int i = 0; i < 9; ++i) r[i] = 0.0f;  // mat3(0.0f), raytracer.cpp:73
    r[4] = rot11;                              // [1][1]
    const float c = cosf(yaw), s = sinf(yaw);  // raytracer.cpp:377-378
    r[0] = c;                                  // [0][0]
    r[2] = s;                                  // [0][2]
    r[6] = -s;                                 // [2][0]
    r[8] = c;                                  // [2][2]
    return B2R_OK;
}

int b2r_orbit_camera(int frame, int nframes, float radius, float* pos, float* rot9) {
    if (!pos || !rot9 || nframes <= 0) return B2R_E_INVALID;
    const float yaw = (float)frame * (2.0f * (float)M_PI / (float)nframes);
    b2r_camera_rot_from_yaw(yaw, 1.0f, rot9);
    // cameraPos = -radius * forward, forward = column 2 (raytracer.cpp:348)
    pos[0] = -radius * rot9[6];
    pos[1] = -radius * rot9[7];
    pos[2] = -radius * rot9[8];
    return B2R_OK;
}

int b2r_jitter_table(unsigned seed, const float* lp, float* out768) {
    if (!lp || !out768) return B2R_E_INVALID;
    memset(out768, 0, sizeof(float) * 768);
    srand(seed);
    for (int i = 0; i < 16; ++i) {
        // RandomNumber() = ((double)rand()/RAND_MAX) - 0.5f (raytracer.cpp:260-263).  The three calls
        // are constructor arguments in the reference (raytracer.cpp:188): g++ evaluates them last
        // to first, so z is drawn first.
        float rz = (float)(((double)rand() / (RAND_MAX)) - 0.5f);
        float ry = (float)(((double)rand() / (RAND_MAX)) - 0.5f);
        float rx = (float)(((double)rand() / (RAND_MAX)) - 0.5f);
        out768[3 * i] = lp[0] + (rx * 0.08f);
        out768[3 * i + 1] = lp[1] + (ry * 0.08f);
        out768[3 * i + 2] = lp[2] + (rz * 0.08f);
    }
    return B2R_OK;
}

int b2r_default_frame_params(b2r_frame_params* p, int which, int width, int height) {
    if (!p || (which != 0 && which != 1) || width <= 0 || height <= 0) return B2R_E_INVALID;
    memset(p, 0, sizeof *p);
    p->numLights = 1;  // AddLight(vec3(0,-0.5f,-0.7f), vec3(1,1,1), 14): raytracer.cpp:116, rasteriser.cpp:104
    p->lights[0].position[0] = 0.0f;
    p->lights[0].position[1] = -0.5f;
    p->lights[0].position[2] = -0.7f;
    p->lights[0].color[0] = p->lights[0].color[1] = p->lights[0].color[2] = 1.0f;
    p->lights[0].intensity = 14.0f;
    p->aaSamples = 3;            // raytracer.cpp:38
    p->softShadowsSamples = 16;  // raytracer.cpp:41
    p->dofKernelSize = 8;        // raytracer.cpp:44
    for (int i = 0; i < 3; ++i) {
        p->indirectLight[i] = 0.2f * 1.0f;  // 0.2f*vec3(1,1,1): raytracer.cpp:81, rasteriser.cpp:47
        p->currentReflectance[i] = 1.0f;    // rasteriser.cpp:466
    }
    p->cameraPos[0] = p->cameraPos[1] = 0.0f;
    if (which == 0) {
        p->cameraPos[2] = -2.0f;                       // raytracer.cpp:70
        p->focalLength = (float)height / 2.0f;         // 250 at 500x500 (raytracer.cpp:69)
        b2r_camera_rot_from_yaw(0.0f, 1.0f, p->cameraRot);   // raytracer.cpp:162
        p->dofFocalLength = 1.3f;                      // raytracer.cpp:45
    } else {
        p->cameraPos[2] = -3.0f;                       // rasteriser.cpp:39
        p->focalLength = (float)height;                // 500 at 500x500 (rasteriser.cpp:41)
        b2r_camera_rot_from_yaw(0.0f, 1.01f, p->cameraRot);  // rasteriser.cpp:115 (sic)
        p->dofFocalLength = 1.9f;                      // rasteriser.cpp:31
        p->backfaceCulling = 1;                        // rasteriser.cpp:26
        p->frustumCulling = 1;                         // rasteriser.cpp:27
    }
    return B2R_OK;
}

}  // extern "C"
