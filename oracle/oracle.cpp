// oracle.cpp — CPU restatement of nr-ray-tracer's per-pixel path-tracing hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (nr_ray_tracer_b200/) may
// include, link or call this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker or as
// the timed CPU baseline, never as the thing shipped.
//
// PARITY UNPINNED: the reference (Rust) cannot be compiled in this environment
// (no cargo/rustc) and ships no tests, golden vectors or fixtures.  Every
// function below restates the cited reference lines (paths relative to
// /root/reference/packages/ray-tracer-lib/src/) including the reference's
// quirks; third-party arithmetic (glam 0.30.9, noise 0.9.0) is restated from
// the published algorithms and is equally unpinned.  RNG: the reference's
// ChaCha8 streams are NOT reproduced (north_star replaces them by Philox);
// the draw *structure* (what is sampled, rejection loops, p/|p|^2) is.
//
// Everything is f64 in the reference's operation order; build with
// -ffp-contract=off so no FMA is contracted (rustc never contracts).
//
// Object model deliberately mirrors the reference: a recursive BVH enum of
// heap nodes with virtual dispatch at leaves, both children always visited
// with the original interval (objects/object.rs:89-121).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/nrrt.h"  // data-description structs only (graph/camera/hit)

namespace oracle {

constexpr double INF = std::numeric_limits<double>::infinity();
constexpr double PI = 3.14159265358979323846264338327950288;

// ---------------------------------------------------------------- glam DVec3
struct V3 {
    double x, y, z;
};
static inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
// glam: (x*x') + (y*y') + (z*z'), left to right
static inline double dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
static inline double length_squared(V3 a) { return dot(a, a); }
// glam normalize: self * (1/length)
static inline V3 normalize(V3 a) { return a * (1.0 / std::sqrt(dot(a, a))); }
// glam reflect: self - 2.0*self.dot(n)*n
static inline V3 reflect(V3 i, V3 n) { return i - (2.0 * dot(i, n)) * n; }
// glam refract
static inline V3 refract(V3 i, V3 n, double eta) {
    double n_dot_i = dot(n, i);
    double k = 1.0 - eta * eta * (1.0 - n_dot_i * n_dot_i);
    if (k >= 0.0) return eta * i - (eta * n_dot_i + std::sqrt(k)) * n;
    return {0, 0, 0};
}
// Rust f64::min/max: NaN-ignoring
static inline double rmin(double a, double b) { return std::fmin(a, b); }
static inline double rmax(double a, double b) { return std::fmax(a, b); }
// Rust f64::signum: +-0 -> +-1, NaN -> NaN
static inline double signum(double x) {
    if (x != x) return x;
    return std::signbit(x) ? -1.0 : 1.0;
}

// DMat3 (column major) — glam from_axis_angle / mul_vec3
struct M3 {
    V3 c0, c1, c2;
};
static M3 mat3_from_axis_angle(V3 axis, double angle) {
    double s = std::sin(angle), c = std::cos(angle);
    double xsin = axis.x * s, ysin = axis.y * s, zsin = axis.z * s;
    double x = axis.x, y = axis.y, z = axis.z;
    double x2 = axis.x * axis.x, y2 = axis.y * axis.y, z2 = axis.z * axis.z;
    double omc = 1.0 - c;
    double xyomc = x * y * omc, xzomc = x * z * omc, yzomc = y * z * omc;
    return M3{v3(x2 * omc + c, xyomc + zsin, xzomc - ysin), v3(xyomc - zsin, y2 * omc + c, yzomc + xsin),
              v3(xzomc + ysin, yzomc - xsin, z2 * omc + c)};
}
static inline V3 mul(const M3& m, V3 v) {
    V3 r = m.c0 * v.x;
    r = r + m.c1 * v.y;
    r = r + m.c2 * v.z;
    return r;
}

// DMat4 (column major): only what Scale needs — from_scale, inverse, transform_point3/vector3
struct M4 {
    double m[4][4];  // m[col][row]
};
static M4 mat4_from_scale(V3 s) {
    M4 r;
    std::memset(&r, 0, sizeof r);
    r.m[0][0] = s.x;
    r.m[1][1] = s.y;
    r.m[2][2] = s.z;
    r.m[3][3] = 1.0;
    return r;
}
// glam DMat4::inverse (scalar path; the GLM cofactor formulation)
static M4 mat4_inverse(const M4& a) {
    double m00 = a.m[0][0], m01 = a.m[0][1], m02 = a.m[0][2], m03 = a.m[0][3];
    double m10 = a.m[1][0], m11 = a.m[1][1], m12 = a.m[1][2], m13 = a.m[1][3];
    double m20 = a.m[2][0], m21 = a.m[2][1], m22 = a.m[2][2], m23 = a.m[2][3];
    double m30 = a.m[3][0], m31 = a.m[3][1], m32 = a.m[3][2], m33 = a.m[3][3];
    double coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    double coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    double coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    double coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    double coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    double coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;
    double fac0[4] = {coef00, coef00, coef02, coef03}, fac1[4] = {coef04, coef04, coef06, coef07};
    double fac2[4] = {coef08, coef08, coef10, coef11}, fac3[4] = {coef12, coef12, coef14, coef15};
    double fac4[4] = {coef16, coef16, coef18, coef19}, fac5[4] = {coef20, coef20, coef22, coef23};
    double vec0[4] = {m10, m00, m00, m00}, vec1[4] = {m11, m01, m01, m01};
    double vec2[4] = {m12, m02, m02, m02}, vec3_[4] = {m13, m03, m03, m03};
    const double sign_a[4] = {1.0, -1.0, 1.0, -1.0}, sign_b[4] = {-1.0, 1.0, -1.0, 1.0};
    M4 inv;
    for (int i = 0; i < 4; ++i) {
        double inv0 = (vec1[i] * fac0[i] - vec2[i] * fac1[i]) + vec3_[i] * fac2[i];
        double inv1 = (vec0[i] * fac0[i] - vec2[i] * fac3[i]) + vec3_[i] * fac4[i];
        double inv2 = (vec0[i] * fac1[i] - vec1[i] * fac3[i]) + vec3_[i] * fac5[i];
        double inv3 = (vec0[i] * fac2[i] - vec1[i] * fac4[i]) + vec2[i] * fac5[i];
        inv.m[0][i] = inv0 * sign_a[i];
        inv.m[1][i] = inv1 * sign_b[i];
        inv.m[2][i] = inv2 * sign_a[i];
        inv.m[3][i] = inv3 * sign_b[i];
    }
    double d0 = a.m[0][0] * inv.m[0][0], d1 = a.m[0][1] * inv.m[1][0];
    double d2 = a.m[0][2] * inv.m[2][0], d3 = a.m[0][3] * inv.m[3][0];
    double det = d0 + d1 + d2 + d3;
    double rcp = 1.0 / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) inv.m[c][r] = inv.m[c][r] * rcp;
    return inv;
}
static inline V3 transform_point3(const M4& a, V3 p) {
    double r[3];
    for (int i = 0; i < 3; ++i) {
        double v = a.m[0][i] * p.x;
        v = a.m[1][i] * p.y + v;
        v = a.m[2][i] * p.z + v;
        v = a.m[3][i] + v;
        r[i] = v;
    }
    return {r[0], r[1], r[2]};
}
static inline V3 transform_vector3(const M4& a, V3 p) {
    double r[3];
    for (int i = 0; i < 3; ++i) {
        double v = a.m[0][i] * p.x;
        v = a.m[1][i] * p.y + v;
        v = a.m[2][i] * p.z + v;
        r[i] = v;
    }
    return {r[0], r[1], r[2]};
}

// ------------------------------------------------------------- interval.rs
struct Interval {
    double min, max;
    static Interval ensure(double a, double b) {  // :22-28
        if (a < b) return {a, b};
        return {b, a};
    }
    Interval unite(const Interval& o) const { return {rmin(min, o.min), rmax(max, o.max)}; }          // :30-35
    Interval intersection(const Interval& o) const { return {rmax(min, o.min), rmin(max, o.max)}; }  // :37-42
    bool is_empty() const { return min > max; }                                                      // :44-46
    Interval pad(double p) const { return {min - p, max + p}; }                                      // :48-53
    double size() const { return max - min; }
    bool contains(double v) const { return min <= v && v <= max; }   // :59-61
    bool surrounds(double v) const { return min < v && v < max; }    // :63-65
};
static const Interval IV_EMPTY{INF, -INF};

// ------------------------------------------------------------------ ray.rs
struct Ray {
    V3 origin, direction;
    size_t bounce;
    double time;
    V3 at(double t) const { return origin + t * direction; }  // :40-42
};

// ----------------------------------------------------------------- aabb.rs
struct Counters {
    uint64_t aabb_tests = 0, prim_tests = 0, segments = 0, paths = 0;
};

struct AABB {
    Interval x, y, z;
    static constexpr double EPSILON = 0.0001;  // :14
    static AABB padded(Interval x, Interval y, Interval z) {  // pad_to_minimums :16-30 via new :34-40
        if (x.size() < EPSILON) x = x.pad((EPSILON - x.size()) / 2.);
        if (y.size() < EPSILON) y = y.pad((EPSILON - y.size()) / 2.);
        if (z.size() < EPSILON) z = z.pad((EPSILON - z.size()) / 2.);
        return AABB{x, y, z};
    }
    AABB unite(const AABB& o) const { return padded(x.unite(o.x), y.unite(o.y), z.unite(o.z)); }  // :42-51
    static AABB from_points(V3 a, V3 b) {  // :53-76
        Interval x = a.x < b.x ? Interval{a.x, b.x} : Interval{b.x, a.x};
        Interval y = a.y < b.y ? Interval{a.y, b.y} : Interval{b.y, a.y};
        Interval z = a.z < b.z ? Interval{a.z, b.z} : Interval{b.z, a.z};
        return padded(x, y, z);
    }
    const Interval& axis_interval(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    int longest_axis() const {  // :101-108 — max_by keeps the LAST maximum
        double s[3] = {x.size(), y.size(), z.size()};
        int best = 0;
        for (int i = 1; i < 3; ++i) {
            // total_cmp(a,b) != Greater  -> replace
            if (!(total_cmp(s[best], s[i]) > 0)) best = i;
        }
        return best;
    }
    static int total_cmp(double a, double b) {  // f64::total_cmp
        int64_t l, r;
        std::memcpy(&l, &a, 8);
        std::memcpy(&r, &b, 8);
        l ^= (int64_t)((uint64_t)(l >> 63) >> 1);
        r ^= (int64_t)((uint64_t)(r >> 63) >> 1);
        return l < r ? -1 : (l > r ? 1 : 0);
    }
    bool hit(const Ray& ray, Interval hit_range) const {  // :110-132
        Interval interval = hit_range;
        const Interval* ax[3] = {&x, &y, &z};
        const double o[3] = {ray.origin.x, ray.origin.y, ray.origin.z};
        const double d[3] = {ray.direction.x, ray.direction.y, ray.direction.z};
        for (int i = 0; i < 3; ++i) {
            interval = interval.intersection(Interval::ensure((ax[i]->min - o[i]) / d[i], (ax[i]->max - o[i]) / d[i]));
            if (interval.is_empty()) return false;
        }
        return true;
    }
    AABB translated(V3 off) const {  // :136-154 (no re-padding)
        return AABB{{x.min + off.x, x.max + off.x}, {y.min + off.y, y.max + off.y}, {z.min + off.z, z.max + off.z}};
    }
};
static const AABB AABB_EMPTY{IV_EMPTY, IV_EMPTY, IV_EMPTY};

// -------------------------------------------------------------- hitable.rs
struct Material;
struct HitRecord {
    bool front_face;
    const Material* material;
    V3 normal, point;
    double t;
    double u, v;
    uint32_t object;  // graph object index of the primitive (test bookkeeping, not in the reference)
};
struct OptHit {
    bool some = false;
    HitRecord h;
};
static HitRecord new_with_uv(const Ray& ray, const Material* m, V3 point, V3 outward, double u, double v, double t,
                             uint32_t object) {  // hitable.rs:38-59
    double sign = signum(dot(ray.direction, outward));
    HitRecord h;
    h.front_face = sign < 0.0;
    h.normal = (-sign) * outward;
    h.material = m;
    h.point = point;
    h.t = t;
    h.u = u;
    h.v = v;
    h.object = object;
    return h;
}

struct Hitable {
    virtual ~Hitable() {}
    virtual AABB bbox() const = 0;
    virtual OptHit hit(const Ray& ray, Interval range, Counters& c) const = 0;
};
using HitPtr = std::shared_ptr<Hitable>;

// ----------------------------------------------------- objects/sphere.rs
struct Sphere : Hitable {
    V3 center, speed;  // speed: Option<DVec3>; None and Some(ZERO) are numerically the same (:76, :110)
    double radius;
    const Material* material;
    AABB box;
    uint32_t object;
    Sphere(V3 c, V3 spd, double r, const Material* m, uint32_t obj)
        : center(c), speed(spd), radius(r), material(m), object(obj) {
        V3 rvec = v3(r, r, r);  // :72
        V3 c0 = center, c1 = center + speed;  // :75-76
        AABB b0 = AABB::from_points(c0 - rvec, c0 + rvec);
        AABB b1 = AABB::from_points(c1 - rvec, c1 + rvec);
        box = b0.unite(b1);
    }
    AABB bbox() const override { return box; }
    OptHit hit(const Ray& ray, Interval range, Counters& cn) const override {  // :105-163
        cn.prim_tests++;
        V3 ctr = center + ray.time * speed;  // Ray::new(center, speed).at(time) (:110-111, ray.rs at())
        V3 dir = ray.direction, eye = ray.origin;
        V3 ec = ctr - eye;
        double a = length_squared(dir);
        double h = dot(ec, dir);
        double c = length_squared(ec) - radius * radius;
        double disc = h * h - a * c;
        OptHit out;
        if (disc < 0.0) return out;
        double sqrtd = std::sqrt(disc);
        double t = (h - sqrtd) / a;
        if (!range.surrounds(t)) {
            t = (h + sqrtd) / a;
            if (!range.surrounds(t)) return out;
        }
        V3 point = ray.at(t);
        V3 normal = normalize(point - ctr);
        double theta = std::acos(-normal.y);
        double phi = std::atan2(-normal.z, normal.x) + PI;
        out.some = true;
        out.h = new_with_uv(ray, material, point, normal, phi / (2.0 * PI), theta / PI, t, object);
        return out;
    }
};

// ------------------------------------------------------ objects/plane.rs
struct Plane : Hitable {
    V3 p, u, v, normal, w;
    double d;
    bool triangle;
    const Material* material;
    AABB box;
    uint32_t object;
    Plane(V3 p_, V3 u_, V3 v_, bool tri, const Material* m, uint32_t obj)
        : p(p_), u(u_), v(v_), triangle(tri), material(m), object(obj) {  // :95-127
        AABB b0 = AABB::from_points(p, p + u + v);
        AABB b1 = AABB::from_points(p + u, p + v);
        box = b0.unite(b1);
        V3 n = cross(u, v);
        normal = normalize(n);
        d = dot(normal, p);
        w = n / dot(n, n);
    }
    AABB bbox() const override { return box; }
    OptHit hit(const Ray& ray, Interval range, Counters& cn) const override {  // :141-174
        cn.prim_tests++;
        OptHit out;
        double denom = dot(normal, ray.direction);
        if (std::fabs(denom) < 1e-8) return out;
        double t = (d - dot(normal, ray.origin)) / denom;
        if (!range.contains(t)) return out;
        V3 point = ray.at(t);
        V3 q = point - p;
        double alpha = dot(w, cross(q, v));
        double beta = dot(w, cross(u, q));
        bool interior;
        if (triangle)
            interior = alpha > 0.0 && beta > 0.0 && (alpha + beta) < 1.0;  // :28-30
        else
            interior = (0.0 <= alpha && alpha <= 1.0) && (0.0 <= beta && beta <= 1.0);  // :23-26
        if (!interior) return out;
        out.some = true;
        out.h = new_with_uv(ray, material, point, normal, alpha, beta, t, object);
        return out;
    }
};

// ----------------------------------------------------- objects/object.rs
struct BVH : Hitable {
    // Leaf(None): leaf && !object ; Leaf(Some): leaf && object ; Node: !leaf
    bool leaf = true;
    HitPtr object;
    AABB box = AABB_EMPTY;
    std::shared_ptr<BVH> left, right;

    static std::shared_ptr<BVH> from(HitPtr* objs, size_t n) {  // :41-73
        auto b = std::make_shared<BVH>();
        if (n == 0) return b;
        if (n == 1) {
            b->object = objs[0];
            return b;
        }
        if (n == 2) {
            b->leaf = false;
            b->left = std::make_shared<BVH>();
            b->left->object = objs[0];
            b->right = std::make_shared<BVH>();
            b->right->object = objs[1];
            b->box = objs[0]->bbox().unite(objs[1]->bbox());
            return b;
        }
        AABB box = AABB_EMPTY;
        for (size_t i = 0; i < n; ++i) box = box.unite(objs[i]->bbox());
        int axis = box.longest_axis();
        std::stable_sort(objs, objs + n, [axis](const HitPtr& a, const HitPtr& c) {
            return AABB::total_cmp(a->bbox().axis_interval(axis).min, c->bbox().axis_interval(axis).min) < 0;
        });
        size_t mid = n / 2;
        b->leaf = false;
        b->box = box;
        b->left = from(objs, mid);
        b->right = from(objs + mid, n - mid);
        return b;
    }
    AABB bbox() const override {  // :77-87
        if (!leaf) return box;
        if (object) return object->bbox();
        return AABB_EMPTY;
    }
    OptHit hit(const Ray& ray, Interval range, Counters& c) const override {  // :89-121
        if (leaf) {
            if (object) return object->hit(ray, range, c);
            return OptHit{};
        }
        c.aabb_tests++;
        if (!box.hit(ray, range)) return OptHit{};
        OptHit l = left->hit(ray, range, c);
        OptHit r = right->hit(ray, range, c);
        if (l.some && !r.some) return l;
        if (!l.some && r.some) return r;
        if (l.some && r.some) return (l.h.t < r.h.t) ? l : r;  // ties -> right
        return OptHit{};
    }
};

// ------------------------------------ objects/translate.rs, rotate.rs, scale.rs
struct Translate : Hitable {
    HitPtr object;
    V3 offset;
    AABB box;
    Translate(HitPtr o, V3 off) : object(o), offset(off) { box = object->bbox().translated(off); }  // :18-29
    AABB bbox() const override { return box; }
    OptHit hit(const Ray& ray, Interval range, Counters& c) const override {  // :37-49
        Ray tr{ray.origin - offset, ray.direction, 0, ray.time};
        OptHit h = object->hit(tr, range, c);
        if (h.some) h.h.point = h.h.point + offset;
        return h;
    }
};

struct Rotate : Hitable {
    HitPtr object;
    M3 rot, rot_inv;
    AABB box;
    static AABB rotate_bbox(const AABB& b, const M3& m) {  // :13-36
        V3 mn = v3(INF, INF, INF), mx = v3(-INF, -INF, -INF);
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) {
                    double x = (double)i * b.x.max + (1.0 - (double)i) * b.x.min;
                    double y = (double)j * b.y.max + (1.0 - (double)j) * b.y.min;
                    double z = (double)k * b.z.max + (1.0 - (double)k) * b.z.min;
                    V3 t = mul(m, v3(x, y, z));
                    mn = v3(rmin(mn.x, t.x), rmin(mn.y, t.y), rmin(mn.z, t.z));
                    mx = v3(rmax(mx.x, t.x), rmax(mx.y, t.y), rmax(mx.z, t.z));
                }
        return AABB::from_points(mn, mx);
    }
    Rotate(HitPtr o, V3 axis, double angle) : object(o) {  // :47-62
        rot = mat3_from_axis_angle(axis, -angle);
        rot_inv = mat3_from_axis_angle(axis, angle);
        box = rotate_bbox(object->bbox(), rot_inv);
    }
    AABB bbox() const override { return box; }
    OptHit hit(const Ray& ray, Interval range, Counters& c) const override {  // :91-106
        Ray rr{mul(rot, ray.origin), mul(rot, ray.direction), 0, ray.time};
        OptHit h = object->hit(rr, range, c);
        if (h.some) {
            h.h.point = mul(rot_inv, h.h.point);
            h.h.normal = mul(rot_inv, h.h.normal);
        }
        return h;
    }
};

struct Scale : Hitable {
    HitPtr object;
    M4 mat, mat_inv;
    AABB box;
    static AABB scale_bbox(const AABB& b, const M4& m) {  // :10-33
        V3 mn = v3(INF, INF, INF), mx = v3(-INF, -INF, -INF);
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) {
                    double x = (double)i * b.x.max + (1.0 - (double)i) * b.x.min;
                    double y = (double)j * b.y.max + (1.0 - (double)j) * b.y.min;
                    double z = (double)k * b.z.max + (1.0 - (double)k) * b.z.min;
                    V3 t = transform_point3(m, v3(x, y, z));
                    mn = v3(rmin(mn.x, t.x), rmin(mn.y, t.y), rmin(mn.z, t.z));
                    mx = v3(rmax(mx.x, t.x), rmax(mx.y, t.y), rmax(mx.z, t.z));
                }
        return AABB::from_points(mn, mx);
    }
    Scale(HitPtr o, V3 s) : object(o) {  // :44-58
        mat = mat4_from_scale(s);
        mat_inv = mat4_inverse(mat);
        box = scale_bbox(object->bbox(), mat);
    }
    AABB bbox() const override { return box; }
    OptHit hit(const Ray& ray, Interval range, Counters& c) const override {  // :73-86
        Ray sr{transform_point3(mat_inv, ray.origin), transform_vector3(mat_inv, ray.direction), 0, ray.time};
        OptHit h = object->hit(sr, range, c);
        if (h.some) h.h.point = transform_point3(mat, h.h.point);  // normal untouched (quirk Q4)
        return h;
    }
};

// --------------------------------------------------------- noise 0.9.0 (restated)
struct Perm {
    uint8_t v[256];
    explicit Perm(uint32_t seed) {
        // PermutationTable::new: XorShiftRng from 16 seed bytes [1,0,0,0, s,s,s], Fisher-Yates (rand 0.8.5)
        uint32_t x = 1, y = seed, z = seed, w = seed;
        auto next = [&]() {
            uint32_t t = x ^ (x << 11);
            x = y;
            y = z;
            z = w;
            w = w ^ (w >> 19) ^ (t ^ (t >> 8));
            return w;
        };
        for (int i = 0; i < 256; ++i) v[i] = (uint8_t)i;
        for (uint32_t i = 255; i >= 1; --i) {
            uint32_t range = i + 1;  // gen_range(0..i+1), rand 0.8.5 UniformInt::sample_single
            int lz = __builtin_clz(range);
            uint32_t zone = (range << lz) - 1u;
            uint32_t idx;
            for (;;) {
                uint32_t r = next();
                uint64_t m = (uint64_t)r * (uint64_t)range;
                uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
                if (lo <= zone) {
                    idx = hi;
                    break;
                }
            }
            std::swap(v[i], v[idx]);
        }
    }
    size_t hash(int64_t a, int64_t b, int64_t c) const {
        size_t i = (size_t)(a & 0xff);
        i = (size_t)v[i] ^ (size_t)(b & 0xff);
        i = (size_t)v[i] ^ (size_t)(c & 0xff);
        return v[i];
    }
};

static inline double grad3(size_t h, double x, double y, double z) {
    switch (h & 15) {
        case 0: case 12: return x + y;
        case 1: case 13: return -x + y;
        case 2: return x - y;
        case 3: return -x - y;
        case 4: return x + z;
        case 5: return -x + z;
        case 6: return x - z;
        case 7: return -x - z;
        case 8: return y + z;
        case 9: case 14: return -y + z;
        case 10: return y - z;
        default: return -y - z;  // 11 | 15
    }
}
static inline double quintic(double t) { return t * t * t * (t * (t * 6.0 - 15.0) + 10.0); }

static double perlin3(const Perm& pm, double px, double py, double pz) {
    const double SCALE = 1.1547005383792515;  // 2/sqrt(3)
    double fx = std::floor(px), fy = std::floor(py), fz = std::floor(pz);
    int64_t cx = (int64_t)fx, cy = (int64_t)fy, cz = (int64_t)fz;
    double dx = px - fx, dy = py - fy, dz = pz - fz;
    auto g = [&](int ox, int oy, int oz) {
        return grad3(pm.hash(cx + ox, cy + oy, cz + oz), dx - (double)ox, dy - (double)oy, dz - (double)oz);
    };
    double g000 = g(0, 0, 0), g100 = g(1, 0, 0), g010 = g(0, 1, 0), g110 = g(1, 1, 0);
    double g001 = g(0, 0, 1), g101 = g(1, 0, 1), g011 = g(0, 1, 1), g111 = g(1, 1, 1);
    double u = quintic(dx), v = quintic(dy), w = quintic(dz);
    double k0 = g000;
    double k1 = g100 - g000;
    double k2 = g010 - g000;
    double k3 = g001 - g000;
    double k4 = g000 + g110 - g100 - g010;
    double k5 = g000 + g101 - g100 - g001;
    double k6 = g000 + g011 - g010 - g001;
    double k7 = g100 + g010 + g001 + g111 - g000 - g110 - g101 - g011;
    double r = k0 + k1 * u + k2 * v + k3 * w + k4 * u * v + k5 * u * w + k6 * v * w + k7 * u * v * w;
    r = r * SCALE;
    return r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
}

struct Fbm {
    std::vector<Perm> sources;
    uint32_t octaves;
    double frequency, lacunarity, persistence, scale_factor;
    Fbm(uint32_t seed, uint32_t oct, double freq, double lac, double pers)
        : frequency(freq), lacunarity(lac), persistence(pers) {
        octaves = std::min<uint32_t>(std::max<uint32_t>(oct, 1u), 32u);  // set_octaves clamps to 1..=32
        for (uint32_t i = 0; i < octaves; ++i) sources.emplace_back(seed + i);
        double denom = 0.0;
        double pw = 1.0;
        for (uint32_t i = 1; i <= octaves; ++i) {
            // persistence.powi(i): repeated multiplication
            pw = 1.0;
            {
                double base = persistence;
                uint32_t n = i;
                for (;;) {
                    if (n & 1) pw *= base;
                    n >>= 1;
                    if (!n) break;
                    base *= base;
                }
            }
            denom = denom + pw;
        }
        scale_factor = 1.0 / denom;
    }
    double get(V3 p) const {
        double x = p.x * frequency, y = p.y * frequency, z = p.z * frequency;
        double result = 0.0, att = persistence;
        for (uint32_t i = 0; i < octaves; ++i) {
            double s = perlin3(sources[i], x, y, z);
            s *= att;
            att *= persistence;
            result += s;
            x *= lacunarity;
            y *= lacunarity;
            z *= lacunarity;
        }
        return result * scale_factor;
    }
};

// ------------------------------------------------------------- textures/*
struct ImageData {
    uint32_t w, h;
    std::vector<float> rgb;  // into_rgb32f(): u8 as f32 / 255
};
struct Texture {
    uint32_t kind;
    V3 color{0, 0, 0};
    const Texture *even = nullptr, *odd = nullptr;
    double scale = 0.5;
    const ImageData* image = nullptr;
    std::unique_ptr<Fbm> fbm;
    double frequency = 1.0;

    V3 get_color(double u, double v, V3 point) const {
        switch (kind) {
            case NRRT_TEX_SOLID: return color;  // solid_color.rs:35-43
            case NRRT_TEX_CHECKER: {            // checker.rs:77-89
                auto as_u64 = [](double x) -> uint64_t {  // Rust saturating `as u64`
                    if (!(x == x)) return 0;
                    if (x <= 0.0) return 0;
                    if (x >= 18446744073709551615.0) return UINT64_MAX;
                    return (uint64_t)x;
                };
                uint64_t s = as_u64(u * scale) + as_u64(v * scale);  // dot(ONE): x*1 + y*1 (wrapping)
                return (s % 2 == 0) ? even->get_color(u, v, point) : odd->get_color(u, v, point);
            }
            case NRRT_TEX_IMAGE: {  // image.rs:30-40
                auto clamp01 = [](double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); };
                auto as_u32 = [](double x) -> uint32_t {
                    if (!(x == x) || x <= 0.0) return 0;
                    if (x >= 4294967295.0) return UINT32_MAX;
                    return (uint32_t)x;
                };
                uint32_t x = as_u32(clamp01(u) * (double)image->w);
                uint32_t y = as_u32((1.0 - clamp01(v)) * (double)image->h);
                // the reference panics for x == w / y == h (u == 1 or v == 0); clamp instead
                if (x >= image->w) x = image->w - 1;
                if (y >= image->h) y = image->h - 1;
                const float* px = &image->rgb[((size_t)y * image->w + x) * 3];
                return v3((double)px[0], (double)px[1], (double)px[2]);
            }
            case NRRT_TEX_NOISE: {  // noise.rs:135-145 — Abs(Fbm)
                double n = std::fabs(fbm->get(point));
                return n * v3(1, 1, 1);
            }
            default: {  // NRRT_TEX_MARBLE  marble.rs:86-97
                double n = std::fabs(fbm->get(point));
                double val = (1. + std::sin(frequency * point.z + 10. * n)) / 2.;
                return val * v3(1, 1, 1);
            }
        }
    }
};

// ------------------------------------------------------------ RNG (Philox4x32-10)
// Counter layout shared with the CUDA kernels (DESIGN.md "Sampling"):
//   key = (seed lo, seed hi); counter = (pixel, sample, stage<<12 | iter, 0)
//   stage 0 = camera ray (iter 0: jitter x,y ; iter 1+j: lens rejection j)
//   stage k+1 = scatter at bounce k (iter j: unit-ball rejection j / dielectric draw)
struct Philox {
    static inline void round(uint32_t c[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
    }
    static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        uint32_t c[4] = {c0, c1, c2, c3};
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c[0];
        out[1] = c[1];
        out[2] = c[2];
        out[3] = c[3];
    }
};

struct Sampler {
    uint64_t seed;
    uint32_t pixel, sample;
    void draw(uint32_t stage, uint32_t iter, uint32_t out[4]) const {
        Philox::gen(seed, pixel, sample, (stage << 12) | (iter & 0xFFFu), 0u, out);
    }
};
static inline double u_m1_1(uint32_t r) { return (double)r * (1.0 / 2147483648.0) - 1.0; }   // random_range(-1.0..1.0)
static inline double u_mh_h(uint32_t r) { return (double)r * (1.0 / 4294967296.0) - 0.5; }   // random_range(-0.5..=0.5)
static inline double u_0_1(uint32_t r) { return (double)r * (1.0 / 4294967296.0); }          // random_range(0.0..1.0)

// vector.rs:61-70 — returns p/|p|^2 (quirk Q1)
static V3 random_in_unit_sphere(const Sampler& s, uint32_t stage) {
    for (uint32_t it = 0;; ++it) {
        uint32_t r[4];
        s.draw(stage, it, r);
        V3 p = v3(u_m1_1(r[0]), u_m1_1(r[1]), u_m1_1(r[2]));
        double l2 = length_squared(p);
        if ((1e-160 < l2 && l2 <= 1.0) || it == 0xFFFu) return p / l2;
    }
}
// vector.rs:72-81 — returns p/|p|^2 (quirk Q2)
static V3 random_in_unit_disk(const Sampler& s) {
    for (uint32_t it = 1;; ++it) {
        uint32_t r[4];
        s.draw(0, it, r);
        V3 p = v3(u_m1_1(r[0]), u_m1_1(r[1]), 0.0);
        double l2 = length_squared(p);
        if (l2 < 1.0 || it == 0xFFFu) return p / l2;
    }
}

// ------------------------------------------------------------- materials/*
struct Material {
    uint32_t kind;
    const Texture* texture;
    double param;
    uint32_t index;

    V3 emit(const Ray& ray, const HitRecord& hit) const {  // diffuse_light.rs:63-75, material.rs:20-26
        if (kind != NRRT_MAT_DIFFUSE_LIGHT) return v3(0, 0, 0);
        double k = ray.bounce > 0 ? param : 1.0;  // quirk Q3
        return k * texture->get_color(hit.u, hit.v, hit.point);
    }
    bool scatter(const Ray& ray, const HitRecord& hit, const Sampler& s, uint32_t stage, Ray& out, V3& color) const {
        switch (kind) {
            case NRRT_MAT_LAMBERTIAN: {  // lambertian.rs:39-55
                V3 dir = hit.normal + random_in_unit_sphere(s, stage);
                if (std::fabs(dir.x) < 1e-8 && std::fabs(dir.y) < 1e-8 && std::fabs(dir.z) < 1e-8) dir = hit.normal;
                out = Ray{hit.point, dir, 0, ray.time};
                color = texture->get_color(hit.u, hit.v, hit.point);
                return true;
            }
            case NRRT_MAT_METAL: {  // metal.rs:73-91
                V3 dir = normalize(reflect(ray.direction, hit.normal)) + param * random_in_unit_sphere(s, stage);
                if (dot(dir, hit.normal) > 0.0) {
                    out = Ray{hit.point, dir, 0, ray.time};
                    color = texture->get_color(hit.u, hit.v, hit.point);
                    return true;
                }
                return false;
            }
            case NRRT_MAT_DIELECTRIC: {  // dielectric.rs:39-67
                double ri = hit.front_face ? 1.0 / param : param;
                V3 unit = normalize(ray.direction);
                double cos_theta = rmin(dot(-unit, hit.normal), 1.0);
                double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
                bool refl = ri * sin_theta > 1.0;
                if (!refl) {
                    double r0 = (1.0 - ri) / (1.0 + ri);  // reflectance :13-19
                    r0 = r0 * r0;
                    double x = 1.0 - cos_theta;
                    double x2 = x * x;
                    double x5 = x * (x2 * x2);  // powi(5): x * (x^2)^2
                    double refl_p = r0 + (1.0 - r0) * x5;
                    uint32_t r[4];
                    s.draw(stage, 0, r);
                    refl = refl_p > u_0_1(r[0]);
                }
                V3 dir = refl ? reflect(unit, hit.normal) : refract(unit, hit.normal, ri);
                out = Ray{hit.point, dir, 0, ray.time};
                color = v3(1, 1, 1);
                return true;
            }
            default: return false;  // DiffuseLight: default scatter = None (material.rs:11-18)
        }
    }
};

// ---------------------------------------------------------------- scene build
struct Scene {
    std::vector<std::unique_ptr<ImageData>> images;
    std::vector<std::unique_ptr<Texture>> textures;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<HitPtr> objects;  // memo per graph object (Ref shares the Arc)
    std::shared_ptr<BVH> root;
    std::string error;
};

static HitPtr build_object(Scene& sc, const nrrt_graph_desc& g, uint32_t idx, int depth) {
    if (idx >= g.n_objects || depth > 64) {
        sc.error = "bad object index / nesting";
        return nullptr;
    }
    if (sc.objects[idx]) return sc.objects[idx];
    const nrrt_object& o = g.objects[idx];
    HitPtr out;
    auto child = [&](uint32_t k) -> HitPtr {
        if (o.first_child + k >= g.n_child_ids) {
            sc.error = "bad child range";
            return nullptr;
        }
        return build_object(sc, g, g.child_ids[o.first_child + k], depth + 1);
    };
    auto mat = [&]() -> const Material* {
        if (o.material >= g.n_materials) {
            sc.error = "bad material index";
            return nullptr;
        }
        return sc.materials[o.material].get();
    };
    switch (o.kind) {
        case NRRT_OBJ_SPHERE: {
            const Material* m = mat();
            if (!m) return nullptr;
            out = std::make_shared<Sphere>(v3(o.v[0], o.v[1], o.v[2]), v3(o.v[4], o.v[5], o.v[6]), o.v[3], m, idx);
            break;
        }
        case NRRT_OBJ_QUAD:
        case NRRT_OBJ_TRIANGLE: {
            const Material* m = mat();
            if (!m) return nullptr;
            out = std::make_shared<Plane>(v3(o.v[0], o.v[1], o.v[2]), v3(o.v[3], o.v[4], o.v[5]),
                                          v3(o.v[6], o.v[7], o.v[8]), o.kind == NRRT_OBJ_TRIANGLE, m, idx);
            break;
        }
        case NRRT_OBJ_GROUP: {
            std::vector<HitPtr> kids;
            for (uint32_t k = 0; k < o.n_children; ++k) {
                HitPtr c = child(k);
                if (!c) return nullptr;
                kids.push_back(c);
            }
            out = BVH::from(kids.data(), kids.size());
            break;
        }
        case NRRT_OBJ_TRANSLATE: {
            HitPtr c = child(0);
            if (!c) return nullptr;
            out = std::make_shared<Translate>(c, v3(o.v[0], o.v[1], o.v[2]));
            break;
        }
        case NRRT_OBJ_ROTATE_X:
        case NRRT_OBJ_ROTATE_Y:
        case NRRT_OBJ_ROTATE_Z: {
            HitPtr c = child(0);
            if (!c) return nullptr;
            V3 axis = o.kind == NRRT_OBJ_ROTATE_X ? v3(1, 0, 0) : (o.kind == NRRT_OBJ_ROTATE_Y ? v3(0, 1, 0) : v3(0, 0, 1));
            out = std::make_shared<Rotate>(c, axis, o.v[0]);
            break;
        }
        case NRRT_OBJ_SCALE: {
            HitPtr c = child(0);
            if (!c) return nullptr;
            out = std::make_shared<Scale>(c, v3(o.v[0], o.v[1], o.v[2]));
            break;
        }
        default: sc.error = "bad object kind"; return nullptr;
    }
    sc.objects[idx] = out;
    return out;
}

static Scene* build_scene(const nrrt_graph_desc& g) {
    auto sc = std::make_unique<Scene>();
    for (uint32_t i = 0; i < g.n_images; ++i) {
        auto im = std::make_unique<ImageData>();
        im->w = g.images[i].width;
        im->h = g.images[i].height;
        size_t n = (size_t)im->w * im->h * 3;
        im->rgb.resize(n);
        for (size_t k = 0; k < n; ++k) im->rgb[k] = (float)g.images[i].rgb[k] / 255.0f;
        sc->images.push_back(std::move(im));
    }
    for (uint32_t i = 0; i < g.n_textures; ++i) sc->textures.push_back(std::make_unique<Texture>());
    for (uint32_t i = 0; i < g.n_textures; ++i) {
        const nrrt_texture& t = g.textures[i];
        Texture& o = *sc->textures[i];
        o.kind = t.kind;
        o.color = v3(t.color[0], t.color[1], t.color[2]);
        if (t.kind == NRRT_TEX_CHECKER) {
            if (t.a >= g.n_textures || t.b >= g.n_textures) return nullptr;
            o.even = sc->textures[t.a].get();
            o.odd = sc->textures[t.b].get();
            o.scale = t.f0;
        } else if (t.kind == NRRT_TEX_IMAGE) {
            if (t.a >= g.n_images) return nullptr;
            o.image = sc->images[t.a].get();
        } else if (t.kind == NRRT_TEX_NOISE) {
            o.fbm = std::make_unique<Fbm>(t.seed, t.octaves, t.f0, t.f1, t.f2);
        } else if (t.kind == NRRT_TEX_MARBLE) {
            // Fbm::new(seed).set_octaves(7).set_frequency(f): default lacunarity 2pi/3, persistence 0.5
            o.fbm = std::make_unique<Fbm>(t.seed, 7, t.f0, PI * 2.0 / 3.0, 0.5);
            o.frequency = t.f0;
        }
    }
    for (uint32_t i = 0; i < g.n_materials; ++i) {
        auto m = std::make_unique<Material>();
        m->kind = g.materials[i].kind;
        m->param = g.materials[i].param;
        m->index = i;
        uint32_t ti = g.materials[i].texture;
        m->texture = (ti < g.n_textures) ? sc->textures[ti].get() : nullptr;
        if (!m->texture && m->kind != NRRT_MAT_DIELECTRIC) return nullptr;
        sc->materials.push_back(std::move(m));
    }
    sc->objects.resize(g.n_objects);
    HitPtr root = build_object(*sc, g, g.root, 0);
    if (!root) return nullptr;
    sc->root = std::dynamic_pointer_cast<BVH>(root);
    if (!sc->root) return nullptr;
    return sc.release();
}

// ------------------------------------------------------------------ camera.rs
static void camera_build(const nrrt_camera_config& c, nrrt_camera& out) {  // :94-159
    std::memset(&out, 0, sizeof out);
    out.width = c.width;
    out.height = c.height;
    out.ray_max_bounces = c.ray_max_bounces;
    out.samples_per_pixel = c.samples_per_pixel < 1 ? 1 : c.samples_per_pixel;
    double defocus_angle = c.defocus_angle < 0. ? 0. : (c.defocus_angle > PI ? PI : c.defocus_angle);
    double focus_dist = c.focus_dist;
    double h = std::tan(c.field_of_view / 2.);
    double viewport_height = focus_dist * h * 2.0;
    double aspect = (double)c.width / (double)c.height;
    double viewport_width = viewport_height * aspect;
    V3 look_from = v3(c.look_from[0], c.look_from[1], c.look_from[2]);
    V3 look_at = v3(c.look_at[0], c.look_at[1], c.look_at[2]);
    V3 view_up = v3(c.view_up[0], c.view_up[1], c.view_up[2]);
    V3 w = normalize(look_from - look_at);
    V3 u = normalize(cross(view_up, w));
    V3 v = normalize(cross(w, u));
    V3 viewport_u = u * viewport_width;
    V3 viewport_v = (-v) * viewport_height;
    V3 du = viewport_u / (double)c.width;
    V3 dv = viewport_v / (double)c.height;
    V3 top_left = look_from - w * focus_dist - viewport_u / 2.0 - viewport_v / 2.0 + (du + dv) / 2.0;
    double defocus_radius = focus_dist * std::tan(defocus_angle / 2.0);
    V3 ddu = u * defocus_radius, ddv = v * defocus_radius;
    auto put = [](double* d, V3 s) {
        d[0] = s.x;
        d[1] = s.y;
        d[2] = s.z;
    };
    out.background[0] = c.background[0];
    out.background[1] = c.background[1];
    out.background[2] = c.background[2];
    put(out.look_from, look_from);
    put(out.defocus_disk_u, ddu);
    put(out.defocus_disk_v, ddv);
    put(out.pixel_delta_u, du);
    put(out.pixel_delta_v, dv);
    put(out.viewport_top_left, top_left);
}

static inline V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }

static Ray get_ray(const nrrt_camera& cam, uint32_t x, uint32_t y, const Sampler& s) {  // :244-267
    double ox = 0.0, oy = 0.0;
    if (cam.samples_per_pixel > 1) {
        uint32_t r[4];
        s.draw(0, 0, r);
        ox = u_mh_h(r[0]);
        oy = u_mh_h(r[1]);
    }
    V3 point = ld3(cam.viewport_top_left) + ((double)x + ox) * ld3(cam.pixel_delta_u) +
               ((double)y + oy) * ld3(cam.pixel_delta_v);
    V3 ddu = ld3(cam.defocus_disk_u), ddv = ld3(cam.defocus_disk_v);
    V3 origin;
    bool no_lens = ddu.x == 0.0 && ddu.y == 0.0 && ddu.z == 0.0 && ddv.x == 0.0 && ddv.y == 0.0 && ddv.z == 0.0;
    if (no_lens) {
        // reference still draws p (:236-242) but p*0 adds exactly zero; the draw is counter-addressed so skipping is exact
        origin = ld3(cam.look_from) + v3(0, 0, 0) + v3(0, 0, 0);
    } else {
        V3 p = random_in_unit_disk(s);
        origin = ld3(cam.look_from) + p.x * ddu + p.y * ddv;
    }
    // time = random_range(0.0..1.0) (:264): third word of the stage-0 / iteration-0 draw.  It only feeds moving
    // spheres (SphereBuilder::with_speed), which no scene file can create but the library API can.
    uint32_t rt[4];
    s.draw(0, 0, rt);
    return Ray{origin, point - origin, 0, u_0_1(rt[2])};
}

static V3 get_ray_color(const nrrt_camera& cam, const Scene& sc, const Ray& ray, size_t bounce, const Sampler& s,
                        Counters& c) {  // :269-300
    if (bounce >= cam.ray_max_bounces) return v3(0, 0, 0);
    c.segments++;
    OptHit h = sc.root->hit(ray, Interval{0.001, INF}, c);
    if (!h.some) return ld3(cam.background);
    const Material* m = h.h.material;
    V3 emitted = m->emit(ray, h.h);
    Ray scattered;
    V3 color;
    if (m->scatter(ray, h.h, s, (uint32_t)bounce + 1, scattered, color)) {
        scattered.bounce += 1;  // :287
        return emitted + color * get_ray_color(cam, sc, scattered, bounce + 1, s, c);
    }
    return emitted;
}

}  // namespace oracle

// =============================================================== C interface
using namespace oracle;

extern "C" {

struct oracle_scene {
    Scene* sc;
};

oracle_scene* oracle_build(const nrrt_graph_desc* g) {
    if (!g) return nullptr;
    Scene* s = build_scene(*g);
    if (!s) return nullptr;
    return new oracle_scene{s};
}
void oracle_free(oracle_scene* s) {
    if (s) {
        delete s->sc;
        delete s;
    }
}

int oracle_camera_build(const nrrt_camera_config* cfg, nrrt_camera* out) {
    if (!cfg || !out || cfg->width == 0 || cfg->height == 0) return -1;
    camera_build(*cfg, *out);
    return 0;
}

// Ray::time of the fixed-ray queries below (default 0): lets the tests probe moving spheres at any shutter time.
static double g_trace_time = 0.0;
void oracle_set_trace_time(double t) { g_trace_time = t; }

// BVH::hit for n rays (bounce flag 0, time = oracle_set_trace_time).  counters: [aabb_tests, prim_tests]
int oracle_trace_rays(const oracle_scene* s, const double* rays, uint64_t n, double tmin, double tmax, nrrt_hit* out,
                      uint64_t* counters, int n_threads) {
    if (!s || !rays || !out) return -1;
    uint64_t aabb = 0, prim = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : aabb, prim)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        Counters c;
        Ray r{ld3(rays + 6 * i), ld3(rays + 6 * i + 3), 0, g_trace_time};
        OptHit h = s->sc->root->hit(r, Interval{tmin, tmax}, c);
        nrrt_hit& o = out[i];
        std::memset(&o, 0, sizeof o);
        if (h.some) {
            o.t = h.h.t;
            o.point[0] = h.h.point.x, o.point[1] = h.h.point.y, o.point[2] = h.h.point.z;
            o.normal[0] = h.h.normal.x, o.normal[1] = h.h.normal.y, o.normal[2] = h.h.normal.z;
            o.uv[0] = h.h.u, o.uv[1] = h.h.v;
            o.object = h.h.object;
            o.prim = 0;
            o.material = h.h.material->index;
            o.front_face = h.h.front_face ? 1u : 0u;
        } else {
            o.t = INF;
            o.object = 0xFFFFFFFFu;
            o.prim = NRRT_REF_NONE;
            o.material = 0xFFFFFFFFu;
        }
        aabb += c.aabb_tests;
        prim += c.prim_tests;
    }
    if (counters) {
        counters[0] = aabb;
        counters[1] = prim;
    }
    return 0;
}

// Camera::render (camera.rs:302-343) for pixels [pixel_begin, pixel_end) (row-major index n, x=n%W, y=n/W),
// samples [sample_begin, sample_end) of each pixel; out_rgb indexed by full-image pixel (W*H*3 f32).
// The value written is sum/(sample_end-sample_begin) cast to f32, as the reference does for the full range.
// counters: [paths, segments, aabb_tests, prim_tests]
int oracle_render(const oracle_scene* s, const nrrt_camera* cam, uint64_t seed, uint32_t pixel_begin, uint32_t pixel_end,
                  uint32_t sample_begin, uint32_t sample_end, float* out_rgb, uint64_t* counters, int n_threads) {
    if (!s || !cam || !out_rgb || sample_end <= sample_begin) return -1;
    uint64_t paths = 0, segs = 0, aabb = 0, prim = 0;
    const uint32_t W = cam->width;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : paths, segs, aabb, prim)
    for (int64_t n = pixel_begin; n < (int64_t)pixel_end; ++n) {
        Counters c;
        uint32_t x = (uint32_t)(n % W), y = (uint32_t)(n / W);
        V3 sum = v3(0, 0, 0);
        for (uint32_t k = sample_begin; k < sample_end; ++k) {
            Sampler smp{seed, (uint32_t)n, k};
            Ray ray = get_ray(*cam, x, y, smp);
            sum = sum + get_ray_color(*cam, *s->sc, ray, 0, smp, c);
            c.paths++;
        }
        V3 color = sum / (double)(sample_end - sample_begin);
        out_rgb[3 * n + 0] = (float)color.x;
        out_rgb[3 * n + 1] = (float)color.y;
        out_rgb[3 * n + 2] = (float)color.z;
        paths += c.paths;
        segs += c.segments;
        aabb += c.aabb_tests;
        prim += c.prim_tests;
    }
    if (counters) {
        counters[0] = paths;
        counters[1] = segs;
        counters[2] = aabb;
        counters[3] = prim;
    }
    return 0;
}

// The same render with the per-pixel sum taken in a GIVEN association: samples [starts[c], starts[c+1]) are summed
// in order into a partial that starts at zero, and the partials are added in order (c = 0 .. n_chunks-1), then
// divided by the sample count.  With one chunk this is oracle_render (the reference's plain `.sum::<DVec3>()`,
// camera.rs:325-329).  A device that splits a pixel's samples into work items sums like this; f64 addition is not
// associative, so a bit-for-bit comparison of a many-sample render needs the association stated (tests only).
int oracle_render_chunked(const oracle_scene* s, const nrrt_camera* cam, uint64_t seed, uint32_t pixel_begin,
                          uint32_t pixel_end, const uint32_t* starts, uint32_t n_chunks, float* out_rgb,
                          uint64_t* counters, int n_threads) {
    if (!s || !cam || !out_rgb || !starts || n_chunks == 0) return -1;
    for (uint32_t c = 0; c < n_chunks; ++c)
        if (starts[c + 1] <= starts[c]) return -1;
    uint64_t paths = 0, segs = 0, aabb = 0, prim = 0;
    const uint32_t W = cam->width;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : paths, segs, aabb, prim)
    for (int64_t n = pixel_begin; n < (int64_t)pixel_end; ++n) {
        Counters c;
        uint32_t x = (uint32_t)(n % W), y = (uint32_t)(n / W);
        V3 sum = v3(0, 0, 0);
        for (uint32_t ch = 0; ch < n_chunks; ++ch) {
            V3 part = v3(0, 0, 0);
            for (uint32_t k = starts[ch]; k < starts[ch + 1]; ++k) {
                Sampler smp{seed, (uint32_t)n, k};
                Ray ray = get_ray(*cam, x, y, smp);
                part = part + get_ray_color(*cam, *s->sc, ray, 0, smp, c);
                c.paths++;
            }
            sum = sum + part;
        }
        V3 color = sum / (double)(starts[n_chunks] - starts[0]);
        out_rgb[3 * n + 0] = (float)color.x;
        out_rgb[3 * n + 1] = (float)color.y;
        out_rgb[3 * n + 2] = (float)color.z;
        paths += c.paths;
        segs += c.segments;
        aabb += c.aabb_tests;
        prim += c.prim_tests;
    }
    if (counters) counters[0] = paths, counters[1] = segs, counters[2] = aabb, counters[3] = prim;
    return 0;
}

// Texture::get_color for n (u,v,point) tuples: in = n x {u,v,px,py,pz}, out = n x rgb
int oracle_texture_eval(const oracle_scene* s, uint32_t texture, const double* in, uint64_t n, double* out) {
    if (!s || texture >= s->sc->textures.size()) return -1;
    const Texture* t = s->sc->textures[texture].get();
    for (uint64_t i = 0; i < n; ++i) {
        V3 c = t->get_color(in[5 * i], in[5 * i + 1], v3(in[5 * i + 2], in[5 * i + 3], in[5 * i + 4]));
        out[3 * i] = c.x, out[3 * i + 1] = c.y, out[3 * i + 2] = c.z;
    }
    return 0;
}

// Philox4x32-10 known-answer hook
void oracle_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out4) {
    Philox::gen(seed, c0, c1, c2, c3, out4);
}

// noise-rs permutation table for a seed (256 bytes)
void oracle_perm_table(uint32_t seed, uint8_t* out256) {
    Perm p(seed);
    std::memcpy(out256, p.v, 256);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
