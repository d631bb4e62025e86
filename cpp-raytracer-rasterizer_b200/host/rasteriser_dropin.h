// rasteriser_dropin.h -- the reference rasteriser's globals and entry points, B200-backed.
//
// Mirrors rasteriser/Source/rasteriser.cpp: same global names and meanings,
// `void Update()` reduced to what feeds Draw() (clears :183-192 happen on the
// device; cameraRot :378-383; the culling block :385-447 runs on the GPU and
// writes Triangle::isCulled back), and `void Draw()` (:86,461) forwarding to
// libb2r.so.  SCREEN_WIDTH/HEIGHT are run-time values here (:35-36).
#pragma once
#include <cstdint>
#include <vector>

#include "reference_types.h"

namespace raref {
using namespace b2rhost;
typedef TriangleRA Triangle;

extern bool BACKFACE_CULLING_ENABLED;            // rasteriser.cpp:26
extern bool FRUSTUM_CULLING_ENABLED;             // :27
extern bool DOF_ENABLED;                         // :29
extern int DOF_KERNEL_SIZE;                      // :30
extern float FOCAL_LENGTH;                       // :31
extern int SCREEN_WIDTH, SCREEN_HEIGHT;          // :35-36 (run-time here)
extern vec3 cameraPos;                           // :39
extern mat3 cameraRot;                           // :40
extern float focalLength;                        // :41
extern float yaw;                                // :42
extern vec3 currentReflectance;                  // :46
extern vec3 indirectLightPowerPerArea;           // :47
extern int NUM_LIGHTS;                           // :49
extern Light lights[32];                         // :50
extern std::vector<float> depthBuffer;           // :52  [y*SCREEN_WIDTH + x]
extern std::vector<Triangle> triangles;          // :64
extern std::vector<float> focalDistances;        // :68
extern std::vector<vec3> pixelColours;           // :69
extern bool isUpdated;                           // :80
extern std::vector<uint32_t> screenPixels;       // screen->pixels (:33), XRGB8888

int Initialize(int width, int height, int device);
void Shutdown();
void LoadTestModel(std::vector<Triangle>& out);               // TestModel.h:151-292
void AddLight(vec3 position, vec3 color, float intensity);    // :158-165
void DeleteLight();                                           // :168-172
void Update();   // :174-449 without SDL
void Draw();     // :461-482 -- the hot path, on the GPU
// The stages Draw() is made of, with the reference's signatures (rasteriser.cpp:532,549,615,674); each call runs
// that stage on the GPU for the current globals (one item per call: for testing and porting, not for speed).
void VertexShader(const Vertex& v, Pixel& p);                                               // :532-546
void Interpolate(Pixel a, Pixel b, std::vector<Pixel>& result);                             // :615-637
void ComputePolygonRows(const std::vector<Pixel>& vertexPixels, std::vector<Pixel>& leftPixels,
                        std::vector<Pixel>& rightPixels);                                   // :674-735
void PixelShader(const Pixel& p, vec3 color, vec3 normal);                                  // :549-589
int SaveBMP(const char* path);
const char* LastError();
}  // namespace raref
