// reference_types.h -- the reference's scene types, re-declared with identical memory layout.
//
// The reference keeps its hot-path state in file-scope globals of these types
// (raytracer.cpp:28,47-48,69-98; rasteriser.cpp:39-70).  The drop-in Draw()
// implementations in this directory read and write globals of the same names
// and types, so code written against the reference (Update(), key handling,
// scene loaders) keeps working.  Layouts (checked by static_assert and by
// tests/test_host_dropin.py against the reference's own sizeof values):
//   Triangle  60 bytes (raytracer TestModel.h:11-32) / 64 bytes with isCulled (rasteriser TestModel.h:10-32)
//   Light     28 bytes (TestModel.h:35-45)      Intersection 20 bytes (raytracer.cpp:91-96)
//   Pixel     24 bytes (rasteriser TestModel.h:34-53)   Vertex 12 bytes (:135-145)
//
// When the reference's own GLM is on the include path, define B2R_WITH_GLM and
// the types below are built from glm::vec3/glm::mat3 exactly like the
// reference's; otherwise a minimal layout-compatible vec3/mat3 is used.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#ifdef B2R_WITH_GLM
#include <glm/glm.hpp>
namespace b2rhost {
using glm::mat3;
using glm::vec3;
inline vec3 normalized(const vec3& v) { return glm::normalize(v); }
inline vec3 crossed(const vec3& a, const vec3& b) { return glm::cross(a, b); }
}  // namespace b2rhost
#else
namespace b2rhost {
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, vec3 b) { a = a - b; return a; }
// column-major 3x3, m[col][row], like glm::mat3
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline vec3 crossed(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline vec3 normalized(const vec3& v) {
    const float inv = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return v * inv;
}
}  // namespace b2rhost
#endif

namespace b2rhost {

// raytracer flavour: no isCulled (raytracer TestModel.h:11-32)
struct TriangleRT {
    vec3 v0, v1, v2, normal, color;
    TriangleRT(vec3 a, vec3 b, vec3 c, vec3 col) : v0(a), v1(b), v2(c), color(col) { ComputeNormal(); }
    void ComputeNormal() { normal = normalized(crossed(v2 - v0, v1 - v0)); }
};
// rasteriser flavour (rasteriser TestModel.h:10-32)
struct TriangleRA {
    vec3 v0, v1, v2, normal, color;
    bool isCulled;
    TriangleRA(vec3 a, vec3 b, vec3 c, vec3 col) : v0(a), v1(b), v2(c), color(col), isCulled(false) { ComputeNormal(); }
    void ComputeNormal() { normal = normalized(crossed(v2 - v0, v1 - v0)); }
};
struct Light {
    vec3 position, color;
    float intensity;
    Light() : intensity(0) {}
    Light(vec3 p, vec3 c, float i) : position(p), color(c), intensity(i) {}
};
struct Intersection {
    vec3 position;
    float distance;
    int triangleIndex;
};
struct Pixel {
    int x, y;
    float zinv;
    vec3 pos3d;
};
struct Vertex {
    vec3 position;
};

static_assert(sizeof(vec3) == 12 && sizeof(mat3) == 36, "vec3/mat3 must match glm's layout");
static_assert(sizeof(TriangleRT) == 60 && sizeof(TriangleRA) == 64, "Triangle layout");
static_assert(sizeof(Light) == 28 && sizeof(Intersection) == 20 && sizeof(Pixel) == 24 && sizeof(Vertex) == 12, "POD layouts");

}  // namespace b2rhost
