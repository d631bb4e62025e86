/* Headless stand-in for <SDL.h> (SDL 1.2), TEST INFRASTRUCTURE ONLY.
 *
 * The reference programs include <SDL.h> for a window, a key-state array, a
 * millisecond tick and a 32-bit surface (SDLauxiliary.h:31-81,
 * raytracer.cpp:115,146,175,332,349,610-654, rasteriser.cpp:103,135,147,177,
 * 198,463-527).  SDL is not installed in this image, and the hot path does not
 * need it, so the oracle build (oracle/build_ref.py) points the compiler at
 * this header instead.  It provides exactly the names those call sites use:
 * an in-memory XRGB8888 surface, no keys pressed, a tick counter, no events.
 * Nothing here is shipped in the product library.
 */
#ifndef B2R_ORACLE_SDL_STUB_H
#define B2R_ORACLE_SDL_STUB_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

typedef uint8_t Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;

struct SDL_PixelFormat {
    Uint8 BitsPerPixel;
    Uint8 BytesPerPixel;
    Uint32 Rmask, Gmask, Bmask, Amask;
};

struct SDL_Surface {
    Uint32 flags;
    SDL_PixelFormat* format;
    int w, h;
    Uint16 pitch;  // bytes per row
    void* pixels;
};

struct SDL_keysym { int sym; };
struct SDL_KeyboardEvent { SDL_keysym keysym; };
struct SDL_Event {
    Uint8 type;
    SDL_KeyboardEvent key;
};

enum { SDL_INIT_TIMER = 0x1, SDL_INIT_VIDEO = 0x20 };
enum { SDL_SWSURFACE = 0, SDL_FULLSCREEN = 0x80000000u };
enum { SDL_KEYDOWN = 2, SDL_QUIT = 12 };

/* key symbols used by Update(); values only need to be distinct indices */
enum {
    SDLK_ESCAPE = 27,
    SDLK_2 = 50, SDLK_3, SDLK_4, SDLK_5, SDLK_6, SDLK_7, SDLK_8, SDLK_9,
    SDLK_LEFTBRACKET = 91, SDLK_RIGHTBRACKET = 93,
    SDLK_a = 97, SDLK_d = 100, SDLK_s = 115, SDLK_w = 119,
    SDLK_UP = 273, SDLK_DOWN, SDLK_RIGHT, SDLK_LEFT,
    SDLK_LAST = 323
};

static inline int SDL_Init(Uint32) { return 0; }
static inline void SDL_Quit(void) {}
static inline const char* SDL_GetError(void) { return "headless SDL stub"; }

static inline SDL_Surface* SDL_SetVideoMode(int w, int h, int bpp, Uint32 flags) {
    static SDL_PixelFormat fmt = {32, 4, 0x00FF0000u, 0x0000FF00u, 0x000000FFu, 0u};
    (void)bpp;
    SDL_Surface* s = (SDL_Surface*)std::calloc(1, sizeof(SDL_Surface));
    s->flags = flags;
    s->format = &fmt;
    s->w = w;
    s->h = h;
    s->pitch = (Uint16)(w * 4);  /* callers index with y*pitch/4; see B2R_STUB_PITCH below */
    s->pixels = std::calloc((size_t)w * (size_t)h, 4);
    return s;
}

/* No events, except that with B2R_STUB_FRAMES=n in the environment the n-th event-pump round reports
 * SDL_QUIT, so the reference's own `while( NoQuitMessageSDL() )` main loop ends after n iterations
 * (used by the integration binaries, oracle/build_ref.py --integration). */
static inline int SDL_PollEvent(SDL_Event* e) {
    static long rounds = 0;
    static long limit = -2;
    if (limit == -2) {
        const char* v = std::getenv("B2R_STUB_FRAMES");
        limit = v ? std::atol(v) : -1;
    }
    if (limit >= 0 && ++rounds > limit) {
        e->type = SDL_QUIT;
        return 1;
    }
    return 0;
}
static inline Uint32 SDL_MapRGB(const SDL_PixelFormat*, Uint8 r, Uint8 g, Uint8 b) {
    return ((Uint32)r << 16) | ((Uint32)g << 8) | (Uint32)b;
}
static inline Uint32 SDL_GetTicks(void) {
    static Uint32 t = 0;
    return t += 16;
}
static inline Uint8* SDL_GetKeyState(int*) {
    static Uint8 keys[SDLK_LAST];
    return keys; /* nothing is ever pressed */
}
#define SDL_MUSTLOCK(s) (0)
static inline int SDL_LockSurface(SDL_Surface*) { return 0; }
static inline void SDL_UnlockSurface(SDL_Surface*) {}
static inline void SDL_UpdateRect(SDL_Surface*, int, int, Uint32, Uint32) {}
/* 24-bit bottom-up BMP of the XRGB surface, rows padded to 4 bytes, 54-byte header: what SDL 1.2 writes. */
static inline int SDL_SaveBMP(SDL_Surface* s, const char* path) {
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return -1;
    const Uint32 pitch = (Uint32)((s->w * 3 + 3) & ~3), size = pitch * (Uint32)s->h;
    unsigned char h[54];
    std::memset(h, 0, sizeof h);
    h[0] = 'B'; h[1] = 'M';
    const Uint32 fields[][2] = {{2, 54u + size}, {10, 54u}, {14, 40u}, {18, (Uint32)s->w}, {22, (Uint32)s->h}, {34, size}};
    for (const auto& fd : fields)
        for (int b = 0; b < 4; ++b) h[fd[0] + b] = (unsigned char)(fd[1] >> (8 * b));
    h[26] = 1; h[28] = 24;
    std::fwrite(h, 1, 54, f);
    unsigned char* row = (unsigned char*)std::calloc(pitch, 1);
    for (int y = s->h - 1; y >= 0; --y) {
        const Uint32* px = (const Uint32*)s->pixels + (size_t)y * (s->pitch / 4);
        for (int x = 0; x < s->w; ++x) {
            row[3 * x] = (unsigned char)(px[x] & 0xFF);
            row[3 * x + 1] = (unsigned char)((px[x] >> 8) & 0xFF);
            row[3 * x + 2] = (unsigned char)((px[x] >> 16) & 0xFF);
        }
        std::fwrite(row, 1, pitch, f);
    }
    std::free(row);
    return std::fclose(f);
}
static inline int SDL_FillRect(SDL_Surface* s, void*, Uint32 c) {
    Uint32* p = (Uint32*)s->pixels;
    for (size_t i = 0, n = (size_t)s->w * (size_t)s->h; i < n; ++i) p[i] = c;
    return 0;
}

#endif
