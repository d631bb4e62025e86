// headless_main.cpp -- the reference's main() loops without SDL: Update(); Draw(); ... SaveBMP.
//
//   b2r_headless raytracer  [W H] [--aa N] [--soft] [--dof] [--frames F] [--out prefix] [--gpus N]
//   b2r_headless rasteriser [W H] [--dof] [--frames F] [--out prefix]
// With --frames F > 1 the camera orbits the box (yaw = f*2pi/F) and one BMP
// per frame is written; otherwise a single screenshot like the reference's Esc key (raytracer.cpp:175).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rasteriser_dropin.h"
#include "raytracer_dropin.h"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s raytracer|rasteriser [W H] [--aa N] [--soft] [--dof] [--frames F] [--out prefix]\n", argv[0]);
        return 2;
    }
    const bool rt = std::strcmp(argv[1], "raytracer") == 0;
    int W = 500, H = 500, frames = 1, aa = 0, soft = 0, dof = 0, a = 2, gpus = 1;
    std::string out = rt ? "raytracer" : "rasteriser";
    if (argc >= 4 && argv[2][0] != '-') {
        W = std::atoi(argv[2]);
        H = std::atoi(argv[3]);
        a = 4;
    }
    for (; a < argc; ++a) {
        if (!std::strcmp(argv[a], "--aa") && a + 1 < argc) aa = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--soft")) soft = 1;
        else if (!std::strcmp(argv[a], "--dof")) dof = 1;
        else if (!std::strcmp(argv[a], "--frames") && a + 1 < argc) frames = std::atoi(argv[++a]);
        else if (!std::strcmp(argv[a], "--out") && a + 1 < argc) out = argv[++a];
        else if (!std::strcmp(argv[a], "--gpus") && a + 1 < argc) gpus = std::atoi(argv[++a]);
    }
    int devices[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    if (gpus < 1 || gpus > 8) gpus = 1;
    int rc = rt ? rtref::InitializeDevices(W, H, devices, gpus) : raref::Initialize(W, H, 0);
    if (rc) {
        std::fprintf(stderr, "init failed (%d): %s\n", rc, rt ? rtref::LastError() : raref::LastError());
        return 1;
    }
    if (rt) {
        rtref::AA_ENABLED = aa > 1;
        if (aa > 1) rtref::AA_SAMPLES = aa;
        rtref::SOFT_SHADOWS_ENABLED = soft;
        rtref::DOF_ENABLED = dof;
    } else {
        raref::DOF_ENABLED = dof;
    }
    double total = 0;
    for (int f = 0; f < frames; ++f) {
        const float yaw = (float)f * (2.0f * 3.14159265358979323846f / (float)frames);
        auto t0 = std::chrono::steady_clock::now();
        if (rt) {
            rtref::yaw = yaw;
            rtref::Update();
            if (frames > 1) rtref::cameraPos = b2rhost::vec3(2.0f * std::sin(yaw), 0.0f, -2.0f * std::cos(yaw));
            rtref::Draw();
        } else {
            raref::yaw = frames > 1 ? 0.3f * std::sin(yaw) : 0.0f;
            raref::isUpdated = true;
            raref::Update();
            raref::Draw();
        }
        total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        char name[512];
        if (frames > 1) std::snprintf(name, sizeof name, "%s_%04d.bmp", out.c_str(), f);
        else std::snprintf(name, sizeof name, "%s.bmp", out.c_str());
        rc = rt ? rtref::SaveBMP(name) : raref::SaveBMP(name);
        if (rc) {
            std::fprintf(stderr, "frame %d failed (%d): %s\n", f, rc, rt ? rtref::LastError() : raref::LastError());
            return 1;
        }
    }
    std::printf("%s %dx%d: %d frame(s), %.3f ms per Update()+Draw() incl. host copies\n", argv[1], W, H, frames, 1e3 * total / frames);
    if (rt) rtref::Shutdown(); else raref::Shutdown();
    return 0;
}
