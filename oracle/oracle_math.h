/* oracle_math.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference).
 *
 * Scalar FP32 helpers that restate the GLM 0.9.7.2 operations the reference's
 * hot path uses, in GLM's evaluation order:
 *   dot        glm/detail/func_geometric.inl:65-72   (x*x + y*y) + z*z
 *   cross      glm/detail/func_geometric.inl:133-142
 *   length     glm/detail/func_geometric.inl:94-100  sqrt(dot(v,v))
 *   distance   glm/detail/func_geometric.inl:111-115 length(p1 - p0)
 *   normalize  glm/detail/func_geometric.inl:153-159 v * (1/sqrt(dot(v,v)))
 *   mat3*vec3  glm/detail/type_mat3x3.inl:506-513
 *   vec3*mat3  glm/detail/type_mat3x3.inl:515-522
 *   inverse    glm/detail/type_mat3x3.inl:36-57
 * Must be compiled with -ffp-contract=off and without -ffast-math / -march
 * flags that enable FMA, exactly like the reference (raytracer/Makefile:13).
 */
#ifndef B2R_ORACLE_MATH_H
#define B2R_ORACLE_MATH_H

#include <math.h>

typedef struct { float x, y, z; } ovec3;
/* column-major like glm::mat3: c[col][row] */
typedef struct { float c[3][3]; } omat3;

static inline ovec3 ov(float x, float y, float z) { ovec3 r = {x, y, z}; return r; }
static inline ovec3 oadd(ovec3 a, ovec3 b) { return ov(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline ovec3 osub(ovec3 a, ovec3 b) { return ov(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline ovec3 omul(ovec3 a, ovec3 b) { return ov(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline ovec3 oscale(ovec3 a, float s) { return ov(a.x * s, a.y * s, a.z * s); }
static inline ovec3 odivs(ovec3 a, float s) { return ov(a.x / s, a.y / s, a.z / s); }
static inline ovec3 oneg(ovec3 a) { return ov(-a.x, -a.y, -a.z); }

static inline float odot(ovec3 a, ovec3 b) {
    ovec3 t = omul(a, b);
    return t.x + t.y + t.z;
}
static inline ovec3 ocross(ovec3 a, ovec3 b) {
    return ov(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline float olength(ovec3 v) { return sqrtf(odot(v, v)); }
static inline float odistance(ovec3 p0, ovec3 p1) { return olength(osub(p1, p0)); }
static inline ovec3 onormalize(ovec3 v) { return oscale(v, 1.0f / sqrtf(odot(v, v))); }

static inline ovec3 omat_vec(const omat3* m, ovec3 v) { /* m * v */
    return ov(m->c[0][0] * v.x + m->c[1][0] * v.y + m->c[2][0] * v.z,
              m->c[0][1] * v.x + m->c[1][1] * v.y + m->c[2][1] * v.z,
              m->c[0][2] * v.x + m->c[1][2] * v.y + m->c[2][2] * v.z);
}
static inline ovec3 ovec_mat(ovec3 v, const omat3* m) { /* v * m */
    return ov(m->c[0][0] * v.x + m->c[0][1] * v.y + m->c[0][2] * v.z,
              m->c[1][0] * v.x + m->c[1][1] * v.y + m->c[1][2] * v.z,
              m->c[2][0] * v.x + m->c[2][1] * v.y + m->c[2][2] * v.z);
}
static inline omat3 oinverse(const omat3* m) {
    const float (*a)[3] = m->c;
    float ood = 1.0f / (+a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2])
                        - a[1][0] * (a[0][1] * a[2][2] - a[2][1] * a[0][2])
                        + a[2][0] * (a[0][1] * a[1][2] - a[1][1] * a[0][2]));
    omat3 r;
    r.c[0][0] = +(a[1][1] * a[2][2] - a[2][1] * a[1][2]) * ood;
    r.c[1][0] = -(a[1][0] * a[2][2] - a[2][0] * a[1][2]) * ood;
    r.c[2][0] = +(a[1][0] * a[2][1] - a[2][0] * a[1][1]) * ood;
    r.c[0][1] = -(a[0][1] * a[2][2] - a[2][1] * a[0][2]) * ood;
    r.c[1][1] = +(a[0][0] * a[2][2] - a[2][0] * a[0][2]) * ood;
    r.c[2][1] = -(a[0][0] * a[2][1] - a[2][0] * a[0][1]) * ood;
    r.c[0][2] = +(a[0][1] * a[1][2] - a[1][1] * a[0][2]) * ood;
    r.c[1][2] = -(a[0][0] * a[1][2] - a[1][0] * a[0][2]) * ood;
    r.c[2][2] = +(a[0][0] * a[1][1] - a[1][0] * a[0][1]) * ood;
    return r;
}

/* std::max<float>(a,b) as used at raytracer.cpp:304 / rasteriser.cpp:582: (a<b)?b:a */
static inline float omaxf(float a, float b) { return (a < b) ? b : a; }

#endif
