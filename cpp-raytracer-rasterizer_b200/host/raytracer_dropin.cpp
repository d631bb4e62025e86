// raytracer_dropin.cpp -- see raytracer_dropin.h.  Host C++ only; all rendering is in libb2r.so.
#include "raytracer_dropin.h"

#include <cstdlib>
#include <cstring>

#include "../../include/b2r.h"

namespace rtref {

std::vector<Triangle> triangles;
bool AA_ENABLED = false;
int AA_SAMPLES = 3;
bool SOFT_SHADOWS_ENABLED = false;
int SOFT_SHADOWS_SAMPLES = 16;
bool DOF_ENABLED = false;
int DOF_KERNEL_SIZE = 8;
float FOCAL_LENGTH = 1.3f;
int NUM_LIGHTS = 0;
Light lights[32];
int SCREEN_WIDTH = 500, SCREEN_HEIGHT = 500;
float focalLength = 250.0f;
vec3 cameraPos(0.0f, 0.0f, -2.0f);
mat3 cameraRot = mat3(0.0f);
float yaw = 0.0f;
bool isUpdated = true;
vec3 indirectLight = 0.2f * vec3(1, 1, 1);
vec3 randomPositions[256];
std::vector<float> focalDistances;
std::vector<vec3> pixelColours;
std::vector<Intersection> closestIntersections;
std::vector<uint32_t> screenPixels;

namespace {
b2r_ctx* g_ctx = nullptr;
b2r_group* g_group = nullptr;  // more than one device: g_ctx is its first member
const Triangle* g_uploaded = nullptr;
size_t g_uploadedCount = 0;
int g_rc = 0;
float RandomNumber() { return ((double)rand() / (RAND_MAX)) - 0.5f; }  // :260-263
}  // namespace

const char* LastError() { return g_group ? b2r_group_last_error(g_group) : b2r_last_error(g_ctx); }

int Initialize(int width, int height, int device) {
    Shutdown();
    SCREEN_WIDTH = width;
    SCREEN_HEIGHT = height;
    focalLength = (float)height / 2.0f;  // 250 at the reference's 500x500 (:69)
    const size_t n = (size_t)width * height;
    focalDistances.assign(n, 0.0f);
    pixelColours.assign(n, vec3());
    screenPixels.assign(n, 0u);
    Intersection init;
    init.position = vec3();
    init.distance = 3.402823466e+38f;  // numeric_limits<float>::max(), :153-160
    init.triangleIndex = -1;
    closestIntersections.assign(n, init);
    cameraRot = mat3(0.0f);
    cameraRot[1][1] = 1.0f;  // :162
    NUM_LIGHTS = 0;
    AddLight(vec3(0, -0.5f, -0.7f), vec3(1, 1, 1), 14);  // :116
    LoadTestModel(triangles);                            // :149
    g_uploaded = nullptr;
    isUpdated = true;
    g_rc = b2r_create(&g_ctx, device, width, height);
    if (g_rc == 0) {  // the frame arrays receive every Draw(): page-lock them so the copies overlap the tracing
        b2r_pin_host_buffer(g_ctx, screenPixels.data(), n * sizeof(uint32_t));
        b2r_pin_host_buffer(g_ctx, pixelColours.data(), n * sizeof(vec3));
        b2r_pin_host_buffer(g_ctx, closestIntersections.data(), n * sizeof(Intersection));
        b2r_pin_host_buffer(g_ctx, focalDistances.data(), n * sizeof(float));
    }
    return g_rc;
}

int InitializeDevices(int width, int height, const int* devices, int n) {
    if (n <= 1) return Initialize(width, height, n == 1 ? devices[0] : 0);
    const int rc = Initialize(width, height, devices[0]);  // the globals; its single context is replaced by the group
    if (rc) return rc;
    Shutdown();
    g_rc = b2r_group_create(&g_group, devices, n, width, height);
    if (g_rc == 0) g_ctx = b2r_group_ctx(g_group, 0);
    return g_rc;
}

void Shutdown() {
    if (g_group) {
        b2r_group_destroy(g_group);
        g_group = nullptr;
        g_ctx = nullptr;
        return;
    }
    if (g_ctx) {
        b2r_unpin_host_buffer(g_ctx, screenPixels.data());
        b2r_unpin_host_buffer(g_ctx, pixelColours.data());
        b2r_unpin_host_buffer(g_ctx, closestIntersections.data());
        b2r_unpin_host_buffer(g_ctx, focalDistances.data());
        b2r_destroy(g_ctx);
    }
    g_ctx = nullptr;
}

void LoadTestModel(std::vector<Triangle>& out) {
    static_assert(sizeof(Triangle) == 60, "raytracer Triangle");
    out.clear();
    out.reserve(30);
    unsigned char raw[30 * 60];
    const int n = b2r_scene_cornell_box(raw, 30, 60);
    for (int i = 0; i < n; ++i) {
        Triangle t(vec3(0, 0, 0), vec3(0, 0, 0), vec3(0, 0, 0), vec3(0, 0, 0));
        std::memcpy(&t, raw + 60 * i, 60);
        out.push_back(t);
    }
}

void AddLight(vec3 position, vec3 color, float intensity) {
    lights[NUM_LIGHTS].position = position;
    lights[NUM_LIGHTS].color = color;
    lights[NUM_LIGHTS].intensity = intensity;
    for (int i = 0; i < SOFT_SHADOWS_SAMPLES; i++) {
        // constructor arguments of :188; g++ evaluates them last to first
        const float rz = RandomNumber(), ry = RandomNumber(), rx = RandomNumber();
        randomPositions[(NUM_LIGHTS * SOFT_SHADOWS_SAMPLES) + i] =
            vec3(position.x + (rx * 0.08f), position.y + (ry * 0.08f), position.z + (rz * 0.08f));
    }
    NUM_LIGHTS++;
}

void DeleteLight() {
    if (NUM_LIGHTS > 0) NUM_LIGHTS--;
}

void Update() {
    // :335-339 resets every Intersection to distance FLT_MAX: b2r_rt_frame does that on the device.
    const float c = std::cos(yaw), s = std::sin(yaw);  // :377-382
    cameraRot[0][0] = c;
    cameraRot[0][2] = s;
    cameraRot[2][0] = -s;
    cameraRot[2][2] = c;
}

namespace {
// scene + frame globals -> device; the scene is (re)uploaded only when the vector changed
int push_state(const std::vector<Triangle>& tris) {
    if (!g_ctx) return B2R_E_NO_SCENE;
    if (g_uploaded != tris.data() || g_uploadedCount != tris.size()) {
        int rc = g_group ? b2r_group_set_triangles(g_group, tris.data(), (int)tris.size(), (int)sizeof(Triangle))
                         : b2r_set_triangles(g_ctx, tris.data(), (int)tris.size(), (int)sizeof(Triangle));
        if (rc) return rc;
        g_uploaded = tris.data();
        g_uploadedCount = tris.size();
    }
    b2r_frame_params p;
    std::memset(&p, 0, sizeof p);
    std::memcpy(p.cameraPos, &cameraPos, 12);
    std::memcpy(p.cameraRot, &cameraRot, 36);
    p.focalLength = focalLength;
    p.numLights = NUM_LIGHTS;
    std::memcpy(p.lights, lights, sizeof(Light) * 32);
    std::memcpy(p.randomPositions, randomPositions, sizeof randomPositions);
    p.aaEnabled = AA_ENABLED;
    p.aaSamples = AA_SAMPLES;
    p.softShadowsEnabled = SOFT_SHADOWS_ENABLED;
    p.softShadowsSamples = SOFT_SHADOWS_SAMPLES;
    std::memcpy(p.indirectLight, &indirectLight, 12);
    p.dofFocalLength = FOCAL_LENGTH;
    p.dofEnabled = DOF_ENABLED;
    p.dofKernelSize = DOF_KERNEL_SIZE;
    p.currentReflectance[0] = p.currentReflectance[1] = p.currentReflectance[2] = 1.0f;
    return g_group ? b2r_group_set_frame(g_group, &p) : b2r_set_frame(g_ctx, &p);
}
}  // namespace

void Draw() {
    if ((g_rc = push_state(triangles)) != 0) return;
    if (g_group) {  // rows are the parallel axis (:557-558): here over GPUs
        g_rc = b2r_group_rt_frame(g_group, screenPixels.data());
        return;
    }
    g_rc = b2r_rt_frame(g_ctx, screenPixels.data(), reinterpret_cast<float*>(pixelColours.data()),
                        reinterpret_cast<b2r_intersection*>(closestIntersections.data()), focalDistances.data());
}

bool ClosestIntersection(vec3 start, vec3 dir, const std::vector<Triangle>& tris, Intersection& closestIntersection,
                         bool isLight, int x, int y) {
    if ((g_rc = push_state(tris)) != 0) return false;
    int32_t light = isLight ? 1 : 0, hit = 0;
    float focal = 0.0f;
    g_rc = b2r_rt_closest_intersection_batch(g_ctx, 1, &start.x, &dir.x, &light,
                                             reinterpret_cast<b2r_intersection*>(&closestIntersection), &hit, &focal);
    // :248-249: the primary-ray call also records the focal distance of an updated hit
    if (g_rc == 0 && !isLight && focal != 0.0f && x >= 0 && y >= 0 && x < SCREEN_WIDTH && y < SCREEN_HEIGHT)
        focalDistances[(size_t)y * SCREEN_WIDTH + x] = focal;
    return g_rc == 0 && hit != 0;
}

vec3 DirectLight(const Intersection& i) {
    vec3 out(0, 0, 0);
    if ((g_rc = push_state(triangles)) != 0) return out;
    g_rc = b2r_rt_direct_light_batch(g_ctx, 1, reinterpret_cast<const b2r_intersection*>(&i), &out.x);
    return out;
}

int SaveBMP(const char* path) {
    if (!g_ctx) return B2R_E_NO_SCENE;
    std::vector<uint8_t> bgr(b2r_bmp_payload_bytes(SCREEN_WIDTH, SCREEN_HEIGHT));
    int rc = 0;
    if (g_group) {  // the frame is in screenPixels (XRGB): 24-bit bottom-up rows padded to 4 bytes, like SDL_SaveBMP
        const size_t rowBytes = ((size_t)SCREEN_WIDTH * 3 + 3) & ~(size_t)3;
        for (int y = 0; y < SCREEN_HEIGHT; ++y) {
            uint8_t* dst = bgr.data() + (size_t)(SCREEN_HEIGHT - 1 - y) * rowBytes;
            const uint32_t* src = screenPixels.data() + (size_t)y * SCREEN_WIDTH;
            for (int x = 0; x < SCREEN_WIDTH; ++x) {
                dst[3 * x] = (uint8_t)(src[x] & 0xFF);
                dst[3 * x + 1] = (uint8_t)((src[x] >> 8) & 0xFF);
                dst[3 * x + 2] = (uint8_t)((src[x] >> 16) & 0xFF);
            }
        }
    } else {
        rc = b2r_resolve_bgr8(g_ctx, bgr.data());
    }
    if (rc) return rc;
    return b2r_write_bmp(path, bgr.data(), SCREEN_WIDTH, SCREEN_HEIGHT);
}

}  // namespace rtref
