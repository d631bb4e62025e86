// raytracer_dropin.h -- the reference raytracer's globals and entry points, B200-backed.
//
// Mirrors raytracer/Source/raytracer.cpp: same global names and meanings
// (cited per declaration), `void Update()` reduced to its effect on Draw()'s
// inputs (no SDL), and `void Draw()` with the reference signature
// (raytracer.cpp:104,547), which forwards to libb2r.so.  SCREEN_WIDTH/HEIGHT are
// run-time values here (the reference fixes them at compile time, :67-68).
#pragma once
#include <cstdint>
#include <vector>

#include "reference_types.h"

namespace rtref {
using namespace b2rhost;
typedef TriangleRT Triangle;

extern std::vector<Triangle> triangles;          // raytracer.cpp:28
extern bool AA_ENABLED;                          // :37
extern int AA_SAMPLES;                           // :38
extern bool SOFT_SHADOWS_ENABLED;                // :40
extern int SOFT_SHADOWS_SAMPLES;                 // :41
extern bool DOF_ENABLED;                         // :43
extern int DOF_KERNEL_SIZE;                      // :44
extern float FOCAL_LENGTH;                       // :45
extern int NUM_LIGHTS;                           // :47
extern Light lights[32];                         // :48
extern int SCREEN_WIDTH, SCREEN_HEIGHT;          // :67-68 (run-time here)
extern float focalLength;                        // :69
extern vec3 cameraPos;                           // :70
extern mat3 cameraRot;                           // :73
extern float yaw;                                // :74
extern bool isUpdated;                           // :78
extern vec3 indirectLight;                       // :81
extern vec3 randomPositions[256];                // :84
extern std::vector<float> focalDistances;        // :87
extern std::vector<vec3> pixelColours;           // :88
extern std::vector<Intersection> closestIntersections;  // :98
extern std::vector<uint32_t> screenPixels;       // screen->pixels of the SDL surface (:76), XRGB8888

// Opens the GPU context for a W x H screen and applies the reference's start-up state
// (main(): :115-116,149-162).  device = CUDA ordinal.  Returns 0 or a B2R_E_* code.
int Initialize(int width, int height, int device);
// Several GPUs of one box behind the same Draw(): device i traces tile rows i, i+n, ... of every frame (b2r_group_rt_frame).
// Draw() then fills screenPixels only -- what the reference shows and saves; the per-pixel arrays stay on the GPUs.
int InitializeDevices(int width, int height, const int* devices, int n);
void Shutdown();
void LoadTestModel(std::vector<Triangle>& out);               // TestModel.h:51-192
void AddLight(vec3 position, vec3 color, float intensity);    // :180-193 (jitter table from glibc rand())
void DeleteLight();                                           // :195-199
void Update();   // :329-545 without SDL: the per-frame reset (:335-339, done on the GPU) and cameraRot from yaw (:377-382)
void Draw();     // :547-606 -- the hot path, on the GPU; fills pixelColours/focalDistances/closestIntersections/screenPixels
// The stages Draw() is made of, with the reference's signatures (raytracer.cpp:105-107); each call runs that
// stage on the GPU for the current globals (one item per call: for testing and porting, not for speed).
bool ClosestIntersection(vec3 start, vec3 dir, const std::vector<Triangle>& triangles,
                         Intersection& closestIntersection, bool isLight, int x, int y);   // :202-257
vec3 DirectLight(const Intersection& i);                                                   // :265-327
int SaveBMP(const char* path);                                // SDL_SaveBMP(screen, path), :175
const char* LastError();
}  // namespace rtref
