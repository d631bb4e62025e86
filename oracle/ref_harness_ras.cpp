// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Harness around the mechanically patched reference rasteriser
// (rasteriser/Source/rasteriser.cpp, which itself #includes LoadSTL.cpp).
// oracle/build_ref.py applies P1/P2/P3/P4 with sed to a
// temporary copy named REF_PATCHED_SOURCE and compiles this translation unit,
// which #includes it to reach the file-scope globals (rasteriser.cpp:22-80)
// and Update()/Draw() (rasteriser.cpp:174,461).
//
//   P3  DrawLineSDL: entries Bresenham() leaves unwritten (rasteriser.cpp:663)
//       are given an off-screen sentinel instead of being read uninitialised.
//   P4  the index of the triangle being drawn is recorded next to every depth
//       buffer write (rasteriser.cpp:608) so the depth *winner* can be compared.
//
// Exported C symbols: ref_ras_*.  One shared object per screen size.

#include <chrono>
#include <cstring>
#include <iostream>
#include <streambuf>

// P4 state, referenced by the patched source.
static int ref_cur_tri = -1;
static int* ref_winner = nullptr;
static long long ref_depth_tests = 0;   // counts loop iterations of rasteriser.cpp:603-611 that are on screen
static long long ref_depth_passes = 0;  // counts writes at rasteriser.cpp:608

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
    screen = InitializeSDL(SCREEN_WIDTH, SCREEN_HEIGHT);
    ref_winner = new int[(size_t)SCREEN_WIDTH * SCREEN_HEIGHT];
    cameraRot = mat3(0.0f);
    cameraRot[1][1] = 1.01f;  // rasteriser.cpp:115 (sic)
    MULTITHREADING_ENABLED = false;
    omp_set_num_threads(1);  // rasteriser.cpp:132-133: the reference default; its OpenMP mode races
}

// The per-frame clears of Update() (rasteriser.cpp:183-192); focalDistances is
// a zero-initialised static the reference never clears, so zero it here to
// keep frames independent.
void clear_frame() {
    const size_t n = (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
    std::memset(depthBuffer, 0, n * sizeof(float));
    std::memset((void*)pixelColours, 0, n * sizeof(vec3));
    std::memset(focalDistances, 0, n * sizeof(float));
    std::memset(screen->pixels, 0, n * 4);
    for (size_t i = 0; i < n; ++i) ref_winner[i] = -1;
    ref_depth_tests = ref_depth_passes = 0;
}
}  // namespace

REF_API int ref_ras_width() { return SCREEN_WIDTH; }
REF_API int ref_ras_height() { return SCREEN_HEIGHT; }
REF_API int ref_ras_sizeof_triangle() { return (int)sizeof(Triangle); }
REF_API int ref_ras_sizeof_pixel() { return (int)sizeof(Pixel); }

REF_API int ref_ras_load_test_model() {
    ensure_init();
    LoadTestModel(triangles);
    return (int)triangles.size();
}

// The reference's STL path (LoadSTL.cpp:17-81) opens "Source/enemy1.stl"
// relative to the working directory, so the caller chdir()s first.
REF_API int ref_ras_load_stl_cwd() {
    ensure_init();
    LoadSTL loader;
    loader.LoadSTLFile(triangles);
    return (int)triangles.size();
}

REF_API void ref_ras_set_triangles(const float* t15, int n) {
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

REF_API int ref_ras_num_triangles() { return (int)triangles.size(); }

REF_API void ref_ras_get_triangles(float* out15) {
    for (size_t i = 0; i < triangles.size(); ++i)
        std::memcpy(out15 + 15 * i, &triangles[i], 15 * sizeof(float));
}

REF_API void ref_ras_set_culled(const unsigned char* c) {
    for (size_t i = 0; i < triangles.size(); ++i) triangles[i].isCulled = c[i] != 0;
}
REF_API void ref_ras_get_culled(unsigned char* c) {
    for (size_t i = 0; i < triangles.size(); ++i) c[i] = triangles[i].isCulled ? 1 : 0;
}

REF_API void ref_ras_set_camera(const float pos[3], const float rot_colmajor[9], float focal) {
    ensure_init();
    cameraPos = vec3(pos[0], pos[1], pos[2]);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) cameraRot[c][r] = rot_colmajor[3 * c + r];
    focalLength = focal;
}

REF_API void ref_ras_get_camera_rot(float out9[9]) {
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) out9[3 * c + r] = cameraRot[c][r];
}

REF_API void ref_ras_set_lights(int n, const float* lights7) {
    ensure_init();
    NUM_LIGHTS = n;
    for (int i = 0; i < n; ++i) {
        const float* p = lights7 + 7 * i;
        lights[i].position = vec3(p[0], p[1], p[2]);
        lights[i].color = vec3(p[3], p[4], p[5]);
        lights[i].intensity = p[6];
    }
}

REF_API void ref_ras_set_flags(int backface, int frustum, int dof, float dofFocalLength) {
    ensure_init();
    BACKFACE_CULLING_ENABLED = backface != 0;
    FRUSTUM_CULLING_ENABLED = frustum != 0;
    DOF_ENABLED = dof != 0;
    FOCAL_LENGTH = dofFocalLength;
}

// Runs the reference's own Update() (rasteriser.cpp:174-449) with no key
// pressed and isUpdated forced on: builds cameraRot from yaw, clears, and
// computes Triangle::isCulled (rasteriser.cpp:375-448).
REF_API double ref_ras_update_yaw(const float pos[3], float yaw_, float focal) {
    ensure_init();
    cameraPos = vec3(pos[0], pos[1], pos[2]);
    yaw = yaw_;
    focalLength = focal;
    isUpdated = true;
    std::streambuf* old = std::cout.rdbuf(&g_nullbuf);
    auto t0 = std::chrono::steady_clock::now();
    Update();
    auto t1 = std::chrono::steady_clock::now();
    std::cout.rdbuf(old);
    return std::chrono::duration<double>(t1 - t0).count();
}

// One frame: clears + Draw() (rasteriser.cpp:461-482).  Returns seconds in Draw().
REF_API double ref_ras_draw(float* outDepth, float* outPixelColours, float* outFocalDistances,
                            int* outWinner, uint32_t* outSurface) {
    ensure_init();
    clear_frame();
    auto t0 = std::chrono::steady_clock::now();
    Draw();
    auto t1 = std::chrono::steady_clock::now();
    const size_t n = (size_t)SCREEN_WIDTH * SCREEN_HEIGHT;
    if (outDepth) std::memcpy(outDepth, depthBuffer, n * sizeof(float));
    if (outPixelColours) std::memcpy(outPixelColours, pixelColours, n * 3 * sizeof(float));
    if (outFocalDistances) std::memcpy(outFocalDistances, focalDistances, n * sizeof(float));
    if (outWinner) std::memcpy(outWinner, ref_winner, n * sizeof(int));
    if (outSurface) std::memcpy(outSurface, screen->pixels, n * 4);
    return std::chrono::duration<double>(t1 - t0).count();
}

REF_API long long ref_ras_depth_tests() { return ref_depth_tests; }
REF_API long long ref_ras_depth_passes() { return ref_depth_passes; }

// ---- sub-stage entry points (reference signatures, flattened to C) ----------
// Pixel is passed as 6 x 4 bytes: int x, int y, float zinv, float pos3d[3]
// (TestModel.h:34-53, 24 bytes).

REF_API void ref_ras_vertex_shader(const float v[3], void* outPixel24) {
    Vertex vert;
    vert.position = vec3(v[0], v[1], v[2]);
    Pixel p;
    VertexShader(vert, p);  // rasteriser.cpp:532
    std::memcpy(outPixel24, &p, sizeof p);
}

// Returns ROWS; left/right must hold maxRows Pixels each.
REF_API int ref_ras_compute_polygon_rows(const void* vertexPixels3x24, void* outLeft, void* outRight,
                                         int maxRows) {
    vector<Pixel> vp(3), l, r;
    std::memcpy(vp.data(), vertexPixels3x24, 3 * sizeof(Pixel));
    ComputePolygonRows(vp, l, r);  // rasteriser.cpp:674
    int rows = (int)l.size();
    if (rows <= maxRows) {
        std::memcpy(outLeft, l.data(), rows * sizeof(Pixel));
        std::memcpy(outRight, r.data(), rows * sizeof(Pixel));
    }
    return rows;
}

REF_API void ref_ras_interpolate(const void* a24, const void* b24, void* out, int n) {
    Pixel a, b;
    std::memcpy(&a, a24, sizeof a);
    std::memcpy(&b, b24, sizeof b);
    vector<Pixel> res(n);
    Interpolate(a, b, res);  // rasteriser.cpp:615
    std::memcpy(out, res.data(), n * sizeof(Pixel));
}

// PixelShader writes pixelColours/focalDistances at (p.x,p.y); read them back.
REF_API void ref_ras_pixel_shader(const void* p24, const float color[3], const float normal[3],
                                  float outColour[3], float* outFocal) {
    ensure_init();
    Pixel p;
    std::memcpy(&p, p24, sizeof p);
    currentReflectance = vec3(1.0f, 1.0f, 1.0f);  // rasteriser.cpp:466
    PixelShader(p, vec3(color[0], color[1], color[2]), vec3(normal[0], normal[1], normal[2]));
    vec3 c = pixelColours[p.y * SCREEN_WIDTH + p.x];
    outColour[0] = c.x; outColour[1] = c.y; outColour[2] = c.z;
    *outFocal = focalDistances[p.y * SCREEN_WIDTH + p.x];
}
