// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Harness around the UNMODIFIED-but-mechanically-patched reference raytracer
// (raytracer/Source/raytracer.cpp).  oracle/build_ref.py pipes the reference
// source through the patches P1/P2, writes the result
// to a temporary file named REF_PATCHED_SOURCE, and compiles this translation
// unit, which #includes it.  Including the file (rather than linking it) gives
// the harness direct access to the reference's file-scope globals
// (raytracer.cpp:28-98) and to Update()/Draw() (raytracer.cpp:329,547).
//
// P5 (triangleIndex = -1 for never-hit pixels) and P6 (main renamed) are done
// here, not by editing the reference.  P7 (row sampling, timing only) replaces
// the bounds of Draw()'s row loop (raytracer.cpp:558) by harness variables.
//
// Exported C symbols: ref_rt_*.  One shared object per compile-time screen
// size (-DREF_W=.. -DREF_H=..), because the reference sizes its static arrays
// with SCREEN_WIDTH/SCREEN_HEIGHT (raytracer.cpp:87-89).

#include <chrono>
#include <cstring>
#include <iostream>
#include <streambuf>

// P7 state (row sampling for bounded timing runs), referenced by the patched Draw() loop.
static int ref_y0 = 0, ref_y1 = REF_H, ref_ystep = 1;

#define main ref_reference_main  // P6
#include REF_PATCHED_SOURCE
#undef main

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;
bool g_inited = false;

void ensure_init() {
    if (g_inited) return;
    g_inited = true;
    screen = InitializeSDL(SCREEN_WIDTH, SCREEN_HEIGHT);  // stub surface, raytracer.cpp:115
    closestIntersections.resize((size_t)SCREEN_WIDTH * SCREEN_HEIGHT);
    cameraRot = mat3(0.0f);
    cameraRot[1][1] = 1.0f;  // raytracer.cpp:162
    MULTITHREADING_ENABLED = true;
}

// The per-frame precondition of Draw(): raytracer.cpp:335-339 (plus P5).
void reset_intersections() {
    const float m = std::numeric_limits<float>::max();
    for (size_t i = 0; i < closestIntersections.size(); ++i) {
        closestIntersections[i].position = vec3(0.0f);
        closestIntersections[i].distance = m;
        closestIntersections[i].triangleIndex = -1;
    }
}
}  // namespace

REF_API int ref_rt_width() { return SCREEN_WIDTH; }
REF_API int ref_rt_height() { return SCREEN_HEIGHT; }
REF_API int ref_rt_sizeof_triangle() { return (int)sizeof(Triangle); }
REF_API int ref_rt_sizeof_intersection() { return (int)sizeof(Intersection); }
REF_API int ref_rt_sizeof_light() { return (int)sizeof(Light); }

// Scene: the reference's own Cornell box (TestModel.h:51-192).
REF_API int ref_rt_load_test_model() {
    ensure_init();
    LoadTestModel(triangles);
    return (int)triangles.size();
}

// Scene: caller-provided triangles as raw reference `Triangle` records
// (v0,v1,v2,normal,color = 15 floats, TestModel.h:11-32).  The normal is
// taken as given (the Triangle ctor would recompute it; callers pass what
// they want the reference to see).
REF_API void ref_rt_set_triangles(const float* t15, int n) {
    ensure_init();
    triangles.clear();
    triangles.reserve(n);
    for (int i = 0; i < n; ++i) {
        const float* p = t15 + 15 * i;
        Triangle t(vec3(p[0], p[1], p[2]), vec3(p[3], p[4], p[5]), vec3(p[6], p[7], p[8]),
                   vec3(p[12], p[13], p[14]));
        t.normal = vec3(p[9], p[10], p[11]);
        triangles.push_back(t);
    }
}

REF_API int ref_rt_num_triangles() { return (int)triangles.size(); }

REF_API void ref_rt_get_triangles(float* out15) {
    for (size_t i = 0; i < triangles.size(); ++i)
        std::memcpy(out15 + 15 * i, &triangles[i], 15 * sizeof(float));
}

// Camera given directly (column-major 3x3 like glm::mat3).
REF_API void ref_rt_set_camera(const float pos[3], const float rot_colmajor[9], float focal) {
    ensure_init();
    cameraPos = vec3(pos[0], pos[1], pos[2]);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) cameraRot[c][r] = rot_colmajor[3 * c + r];
    focalLength = focal;
}

// Camera given as the reference's own state (yaw); the rotation matrix is
// then built by the reference's Update() (raytracer.cpp:377-382).
REF_API void ref_rt_set_camera_yaw(const float pos[3], float yaw_, float focal) {
    ensure_init();
    cameraPos = vec3(pos[0], pos[1], pos[2]);
    yaw = yaw_;
    focalLength = focal;
    std::streambuf* old = std::cout.rdbuf(&g_nullbuf);
    Update();
    std::cout.rdbuf(old);
}

REF_API void ref_rt_get_camera_rot(float out9[9]) {
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) out9[3 * c + r] = cameraRot[c][r];
}

// lights7: n x {position[3], color[3], intensity} (TestModel.h:35-45).
// random768: randomPositions table (raytracer.cpp:84), 256 x vec3, or NULL.
REF_API void ref_rt_set_lights(int n, const float* lights7, const float* random768) {
    ensure_init();
    NUM_LIGHTS = n;
    for (int i = 0; i < n; ++i) {
        const float* p = lights7 + 7 * i;
        lights[i].position = vec3(p[0], p[1], p[2]);
        lights[i].color = vec3(p[3], p[4], p[5]);
        lights[i].intensity = p[6];
    }
    if (random768)
        for (int i = 0; i < 256; ++i)
            randomPositions[i] = vec3(random768[3 * i], random768[3 * i + 1], random768[3 * i + 2]);
}

// Runs the reference's own AddLight (raytracer.cpp:180-193) after srand(seed)
// so tests can obtain the glibc rand() jitter table the reference would use.
REF_API void ref_rt_add_light_reference(unsigned seed, int reset, const float pos[3],
                                        const float color[3], float intensity) {
    ensure_init();
    if (reset) NUM_LIGHTS = 0;
    srand(seed);
    AddLight(vec3(pos[0], pos[1], pos[2]), vec3(color[0], color[1], color[2]), intensity);
}

REF_API void ref_rt_get_random_positions(float* out768) {
    std::memcpy(out768, randomPositions, sizeof(float) * 768);
}

REF_API void ref_rt_set_flags(int aa, int aaSamples, int soft, int softSamples, int dof,
                              float dofFocalLength, int threads) {
    ensure_init();
    AA_ENABLED = aa != 0;
    AA_SAMPLES = aaSamples;
    SOFT_SHADOWS_ENABLED = soft != 0;
    SOFT_SHADOWS_SAMPLES = softSamples;
    DOF_ENABLED = dof != 0;
    FOCAL_LENGTH = dofFocalLength;
    if (threads > 0) omp_set_num_threads(threads);
}

// P7: Draw() visits rows y0, y0+step, ... < y1 (default: every row).
REF_API void ref_rt_set_rows(int y0, int y1, int step) {
    ref_y0 = y0;
    ref_y1 = y1;
    ref_ystep = step > 0 ? step : 1;
}

REF_API void ref_rt_set_indirect(const float v[3]) { indirectLight = vec3(v[0], v[1], v[2]); }

// One frame: the Update() resets + Draw() (raytracer.cpp:547-606), then copy
// out whatever the caller asked for.  Returns seconds spent inside Draw().
REF_API double ref_rt_draw(float* outPixelColours, float* outFocalDistances, void* outClosest20,
                           uint32_t* outSurface) {
    ensure_init();
    reset_intersections();
    std::memset(screen->pixels, 0, (size_t)SCREEN_WIDTH * SCREEN_HEIGHT * 4);
    // focalDistances is a zero-initialised static the reference never clears (pixels that
    // miss keep the previous frame's value); zero it so frames are independent.
    std::memset(focalDistances, 0, sizeof(focalDistances));
    auto t0 = std::chrono::steady_clock::now();
    Draw();
    auto t1 = std::chrono::steady_clock::now();
    const size_t n = (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
    if (outPixelColours) std::memcpy(outPixelColours, pixelColours, n * 3 * sizeof(float));
    if (outFocalDistances) std::memcpy(outFocalDistances, focalDistances, n * sizeof(float));
    if (outClosest20) std::memcpy(outClosest20, closestIntersections.data(), n * sizeof(Intersection));
    if (outSurface) std::memcpy(outSurface, screen->pixels, n * 4);
    return std::chrono::duration<double>(t1 - t0).count();
}

// Sub-stage entry points with the reference signatures (raytracer.cpp:105-107),
// flattened to C so the sub-stage parity tests can drive them.
REF_API int ref_rt_closest_intersection(const float start[3], const float dir[3], float* ioClosest5,
                                        int isLight) {
    Intersection c;
    std::memcpy(&c, ioClosest5, sizeof c);
    bool hit = ClosestIntersection(vec3(start[0], start[1], start[2]), vec3(dir[0], dir[1], dir[2]),
                                   triangles, c, isLight != 0, 0, 0);
    std::memcpy(ioClosest5, &c, sizeof c);
    return hit ? 1 : 0;
}

REF_API void ref_rt_direct_light(const float* closest5, float out3[3]) {
    Intersection c;
    std::memcpy(&c, closest5, sizeof c);
    vec3 r = DirectLight(c);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
