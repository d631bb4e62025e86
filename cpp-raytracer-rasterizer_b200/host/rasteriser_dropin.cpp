// rasteriser_dropin.cpp -- see rasteriser_dropin.h.  Host C++ only; all rendering is in libb2r.so.
#include "rasteriser_dropin.h"

#include <cstring>

#include "../../include/b2r.h"

namespace raref {

bool BACKFACE_CULLING_ENABLED = true;
bool FRUSTUM_CULLING_ENABLED = true;
bool DOF_ENABLED = false;
int DOF_KERNEL_SIZE = 8;
float FOCAL_LENGTH = 1.9f;
int SCREEN_WIDTH = 500, SCREEN_HEIGHT = 500;
vec3 cameraPos(0, 0, -3.0f);
mat3 cameraRot = mat3(0.0f);
float focalLength = 500.0f;
float yaw = 0;
vec3 currentReflectance;
vec3 indirectLightPowerPerArea = 0.2f * vec3(1, 1, 1);
int NUM_LIGHTS = 0;
Light lights[32];
std::vector<float> depthBuffer;
std::vector<Triangle> triangles;
std::vector<float> focalDistances;
std::vector<vec3> pixelColours;
bool isUpdated = true;
std::vector<uint32_t> screenPixels;

namespace {
b2r_ctx* g_ctx = nullptr;
const Triangle* g_uploaded = nullptr;
size_t g_uploadedCount = 0;
int g_rc = 0;

void fill_params(b2r_frame_params& p) {
    std::memset(&p, 0, sizeof p);
    std::memcpy(p.cameraPos, &cameraPos, 12);
    std::memcpy(p.cameraRot, &cameraRot, 36);
    p.focalLength = focalLength;
    p.numLights = NUM_LIGHTS;
    std::memcpy(p.lights, lights, sizeof(Light) * 32);
    std::memcpy(p.indirectLight, &indirectLightPowerPerArea, 12);
    std::memcpy(p.currentReflectance, &currentReflectance, 12);
    p.aaSamples = 1;
    p.softShadowsSamples = 1;
    p.dofFocalLength = FOCAL_LENGTH;
    p.dofEnabled = DOF_ENABLED;
    p.dofKernelSize = DOF_KERNEL_SIZE;
    p.backfaceCulling = BACKFACE_CULLING_ENABLED;
    p.frustumCulling = FRUSTUM_CULLING_ENABLED;
}

int upload_scene_if_changed() {
    if (g_uploaded == triangles.data() && g_uploadedCount == triangles.size()) return 0;
    int rc = b2r_set_triangles(g_ctx, triangles.data(), (int)triangles.size(), (int)sizeof(Triangle));
    if (rc == 0) {
        g_uploaded = triangles.data();
        g_uploadedCount = triangles.size();
    }
    return rc;
}
}  // namespace

const char* LastError() { return b2r_last_error(g_ctx); }

int Initialize(int width, int height, int device) {
    Shutdown();
    SCREEN_WIDTH = width;
    SCREEN_HEIGHT = height;
    focalLength = (float)height;  // 500 at the reference's 500x500 (:41)
    const size_t n = (size_t)width * height;
    depthBuffer.assign(n, 0.0f);
    focalDistances.assign(n, 0.0f);
    pixelColours.assign(n, vec3());
    screenPixels.assign(n, 0u);
    cameraRot = mat3(0.0f);
    cameraRot[1][1] = 1.01f;  // :115 (sic)
    NUM_LIGHTS = 0;
    AddLight(vec3(0, -0.5f, -0.7f), vec3(1, 1, 1), 14);  // :104
    LoadTestModel(triangles);                            // :112
    g_uploaded = nullptr;
    isUpdated = true;
    g_rc = b2r_create(&g_ctx, device, width, height);
    if (g_rc == 0) {  // the frame arrays receive every Draw(): page-lock them so the copies run at full PCIe speed
        b2r_pin_host_buffer(g_ctx, screenPixels.data(), n * sizeof(uint32_t));
        b2r_pin_host_buffer(g_ctx, pixelColours.data(), n * sizeof(vec3));
        b2r_pin_host_buffer(g_ctx, depthBuffer.data(), n * sizeof(float));
        b2r_pin_host_buffer(g_ctx, focalDistances.data(), n * sizeof(float));
    }
    return g_rc;
}

void Shutdown() {
    if (g_ctx) {
        b2r_unpin_host_buffer(g_ctx, screenPixels.data());
        b2r_unpin_host_buffer(g_ctx, pixelColours.data());
        b2r_unpin_host_buffer(g_ctx, depthBuffer.data());
        b2r_unpin_host_buffer(g_ctx, focalDistances.data());
        b2r_destroy(g_ctx);
    }
    g_ctx = nullptr;
}

void LoadTestModel(std::vector<Triangle>& out) {
    static_assert(sizeof(Triangle) == 64, "rasteriser Triangle");
    out.clear();
    out.reserve(30);
    unsigned char raw[30 * 64];
    const int n = b2r_scene_cornell_box(raw, 30, 64);
    for (int i = 0; i < n; ++i) {
        Triangle t(vec3(0, 0, 0), vec3(0, 0, 0), vec3(0, 0, 0), vec3(0, 0, 0));
        std::memcpy(static_cast<void*>(&t), raw + 64 * i, 60);  // v0, v1, v2, normal, color: 15 floats
        t.isCulled = false;
        out.push_back(t);
    }
}

void AddLight(vec3 position, vec3 color, float intensity) {
    lights[NUM_LIGHTS].position = position;
    lights[NUM_LIGHTS].color = color;
    lights[NUM_LIGHTS].intensity = intensity;
    NUM_LIGHTS++;
}

void DeleteLight() {
    if (NUM_LIGHTS > 0) NUM_LIGHTS--;
}

void Update() {
    // :183-192 clears depthBuffer/pixelColours/screen: b2r_ras_frame starts from cleared device buffers.
    if (!isUpdated || !g_ctx) return;
    const float c = std::cos(yaw), s = std::sin(yaw);  // :378-383
    cameraRot[0][0] = c;
    cameraRot[0][2] = s;
    cameraRot[2][0] = -s;
    cameraRot[2][2] = c;
    // :385-447: Triangle::isCulled for this camera, computed on the GPU and written back
    if ((g_rc = upload_scene_if_changed()) != 0) return;
    b2r_frame_params p;
    fill_params(p);
    if ((g_rc = b2r_set_frame(g_ctx, &p)) != 0) return;
    std::vector<uint8_t> culled(triangles.size());
    if ((g_rc = b2r_ras_cull(g_ctx, culled.data())) != 0) return;
    for (size_t i = 0; i < triangles.size(); ++i) triangles[i].isCulled = culled[i] != 0;
}

void Draw() {
    if (!g_ctx) {
        g_rc = B2R_E_NO_SCENE;
        return;
    }
    currentReflectance = vec3(1.0f, 1.0f, 1.0f);  // :466
    if ((g_rc = upload_scene_if_changed()) != 0) return;
    // isCulled may have been edited by the caller since Update(): send the flags the reference would read (:470)
    std::vector<uint8_t> culled(triangles.size());
    for (size_t i = 0; i < triangles.size(); ++i) culled[i] = triangles[i].isCulled ? 1 : 0;
    if ((g_rc = b2r_set_culled(g_ctx, culled.data(), (int)culled.size())) != 0) return;
    b2r_frame_params p;
    fill_params(p);
    if ((g_rc = b2r_set_frame(g_ctx, &p)) != 0) return;
    g_rc = b2r_ras_frame(g_ctx, screenPixels.data(), depthBuffer.data(), reinterpret_cast<float*>(pixelColours.data()),
                         focalDistances.data(), nullptr);
}

namespace {
int push_frame() {
    if (!g_ctx) return B2R_E_NO_SCENE;
    b2r_frame_params p;
    fill_params(p);
    return b2r_set_frame(g_ctx, &p);
}
}  // namespace

void VertexShader(const Vertex& v, Pixel& p) {
    if ((g_rc = push_frame()) != 0) return;
    g_rc = b2r_ras_vertex_shader_batch(g_ctx, 1, &v.position.x, &p);
}

void Interpolate(Pixel a, Pixel b, std::vector<Pixel>& result) {
    if (!g_ctx) {
        g_rc = B2R_E_NO_SCENE;
        return;
    }
    g_rc = b2r_ras_interpolate(g_ctx, &a, &b, (int)result.size(), result.data());
}

void ComputePolygonRows(const std::vector<Pixel>& vertexPixels, std::vector<Pixel>& leftPixels,
                        std::vector<Pixel>& rightPixels) {
    if (!g_ctx || vertexPixels.size() < 3) {
        g_rc = B2R_E_NO_SCENE;
        return;
    }
    int maxY = vertexPixels[0].y, minY = vertexPixels[0].y;
    for (int k = 1; k < 3; ++k) {
        if (vertexPixels[k].y > maxY) maxY = vertexPixels[k].y;
        if (vertexPixels[k].y < minY) minY = vertexPixels[k].y;
    }
    const long long rows = (long long)maxY - minY + 1;  // :682
    if (rows > (1 << 22)) {
        g_rc = B2R_E_CAPACITY;
        return;
    }
    leftPixels.resize((size_t)rows);   // :687-688
    rightPixels.resize((size_t)rows);
    int got = 0;
    g_rc = b2r_ras_compute_polygon_rows(g_ctx, vertexPixels.data(), leftPixels.data(), rightPixels.data(), (int)rows, &got);
}

void PixelShader(const Pixel& p, vec3 color, vec3 normal) {
    if ((g_rc = push_frame()) != 0) return;
    vec3 out(0, 0, 0);
    float focal = 0.0f;
    g_rc = b2r_ras_pixel_shader_batch(g_ctx, 1, &p, &color.x, &normal.x, &out.x, &focal);
    if (g_rc == 0 && p.x >= 0 && p.y >= 0 && p.x < SCREEN_WIDTH && p.y < SCREEN_HEIGHT) {
        focalDistances[(size_t)p.y * SCREEN_WIDTH + p.x] = focal;   // :565
        pixelColours[(size_t)p.y * SCREEN_WIDTH + p.x] = out;       // :588
    }
}

int SaveBMP(const char* path) {
    if (!g_ctx) return B2R_E_NO_SCENE;
    std::vector<uint8_t> bgr(b2r_bmp_payload_bytes(SCREEN_WIDTH, SCREEN_HEIGHT));
    int rc = b2r_resolve_bgr8(g_ctx, bgr.data());
    if (rc) return rc;
    return b2r_write_bmp(path, bgr.data(), SCREEN_WIDTH, SCREEN_HEIGHT);
}

}  // namespace raref
