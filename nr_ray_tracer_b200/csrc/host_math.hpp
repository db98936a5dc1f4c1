// host_math.hpp — f64 vector / matrix / box helpers for the host scene layer.
//
// These restate the handful of glam 0.30.9 and aabb.rs operations whose exact operation order
// decides box coordinates and plane constants (SURVEY.md note A).  No FMA: build with
// -ffp-contract=off (g++) or -fmad=false (nvcc host pass uses the host compiler's default, which
// for x86-64 without -march flags has no FMA to contract into).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace nrrt_host {

struct D3 {
    double x, y, z;
    D3 operator+(D3 b) const { return {x + b.x, y + b.y, z + b.z}; }
    D3 operator-(D3 b) const { return {x - b.x, y - b.y, z - b.z}; }
    D3 operator-() const { return {-x, -y, -z}; }
    D3 operator*(double s) const { return {x * s, y * s, z * s}; }
    D3 operator/(double s) const { return {x / s, y / s, z / s}; }
    double dot(D3 b) const { return (x * b.x) + (y * b.y) + (z * b.z); }  // glam: left-to-right sum
    D3 cross(D3 b) const { return {y * b.z - b.y * z, z * b.x - b.z * x, x * b.y - b.x * y}; }
    D3 normalize() const { return *this * (1.0 / std::sqrt(dot(*this))); }  // glam: v * length_recip()
};

// f64::total_cmp
inline int total_cmp(double a, double b) {
    int64_t l, r;
    std::memcpy(&l, &a, 8);
    std::memcpy(&r, &b, 8);
    l ^= (int64_t)((uint64_t)(l >> 63) >> 1);
    r ^= (int64_t)((uint64_t)(r >> 63) >> 1);
    return (l > r) - (l < r);
}

// aabb.rs:6-108 as a min/max pair per axis
struct Aabb {
    double lo[3], hi[3];

    static Aabb empty() {  // AABB::EMPTY :78-82
        const double inf = std::numeric_limits<double>::infinity();
        return Aabb{{inf, inf, inf}, {-inf, -inf, -inf}};
    }
    // AABB::new -> pad_to_minimums (:16-40): every axis at least EPSILON = 1e-4 wide
    static Aabb padded(Aabb b) {
        const double EPSILON = 0.0001;
        for (int a = 0; a < 3; ++a) {
            double size = b.hi[a] - b.lo[a];
            if (size < EPSILON) {
                double pad = (EPSILON - size) / 2.;
                b.lo[a] = b.lo[a] - pad;
                b.hi[a] = b.hi[a] + pad;
            }
        }
        return b;
    }
    Aabb unite(const Aabb& o) const {  // :42-51 (Interval::union uses NaN-ignoring f64::min/max)
        Aabb r;
        for (int a = 0; a < 3; ++a) {
            r.lo[a] = std::fmin(lo[a], o.lo[a]);
            r.hi[a] = std::fmax(hi[a], o.hi[a]);
        }
        return padded(r);
    }
    static Aabb from_points(D3 a, D3 b) {  // :53-76
        const double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
        Aabb r;
        for (int k = 0; k < 3; ++k) {
            if (av[k] < bv[k])
                r.lo[k] = av[k], r.hi[k] = bv[k];
            else
                r.lo[k] = bv[k], r.hi[k] = av[k];
        }
        return padded(r);
    }
    Aabb translated(D3 off) const {  // :136-154 — no re-padding
        const double o[3] = {off.x, off.y, off.z};
        Aabb r;
        for (int a = 0; a < 3; ++a) r.lo[a] = lo[a] + o[a], r.hi[a] = hi[a] + o[a];
        return r;
    }
    int longest_axis() const {  // :101-108 — Iterator::max_by returns the LAST maximum
        int best = 0;
        for (int a = 1; a < 3; ++a)
            if (total_cmp(hi[best] - lo[best], hi[a] - lo[a]) <= 0) best = a;
        return best;
    }
};

// glam DMat3 (column major)
struct Mat3 {
    D3 c0, c1, c2;
    static Mat3 from_axis_angle(D3 axis, double angle) {
        double s = std::sin(angle), c = std::cos(angle);
        double xs = axis.x * s, ys = axis.y * s, zs = axis.z * s;
        double x2 = axis.x * axis.x, y2 = axis.y * axis.y, z2 = axis.z * axis.z;
        double omc = 1.0 - c;
        double xyomc = axis.x * axis.y * omc, xzomc = axis.x * axis.z * omc, yzomc = axis.y * axis.z * omc;
        return Mat3{D3{x2 * omc + c, xyomc + zs, xzomc - ys}, D3{xyomc - zs, y2 * omc + c, yzomc + xs},
                    D3{xzomc + ys, yzomc - xs, z2 * omc + c}};
    }
    D3 mul(D3 v) const { return (c0 * v.x + c1 * v.y) + c2 * v.z; }
};

// glam DMat4 (column major, m[col][row]) — from_scale / inverse / transform_point3
struct Mat4 {
    double m[4][4];
    static Mat4 from_scale(D3 s) {
        Mat4 r;
        std::memset(&r, 0, sizeof r);
        r.m[0][0] = s.x, r.m[1][1] = s.y, r.m[2][2] = s.z, r.m[3][3] = 1.0;
        return r;
    }
    D3 transform_point3(D3 p) const {
        double o[3];
        for (int i = 0; i < 3; ++i) o[i] = m[3][i] + (m[2][i] * p.z + (m[1][i] * p.y + m[0][i] * p.x));
        return {o[0], o[1], o[2]};
    }
    // General 4x4 inverse by 2x2 sub-determinant cofactors scaled by 1/det (glam's scalar path).
    Mat4 inverse() const {
        const double(*a)[4] = m;
        // 2x2 minors of rows {2,3} / {1,3} / {1,2} for column pairs
        auto sub = [&](int c0, int r0, int c1, int r1) { return a[c0][r0] * a[c1][r1] - a[c1][r0] * a[c0][r1]; };
        double s00 = sub(2, 2, 3, 3), s02 = sub(1, 2, 3, 3), s03 = sub(1, 2, 2, 3);
        double s04 = sub(2, 1, 3, 3), s06 = sub(1, 1, 3, 3), s07 = sub(1, 1, 2, 3);
        double s08 = sub(2, 1, 3, 2), s10 = sub(1, 1, 3, 2), s11 = sub(1, 1, 2, 2);
        double s12 = sub(2, 0, 3, 3), s14 = sub(1, 0, 3, 3), s15 = sub(1, 0, 2, 3);
        double s16 = sub(2, 0, 3, 2), s18 = sub(1, 0, 3, 2), s19 = sub(1, 0, 2, 2);
        double s20 = sub(2, 0, 3, 1), s22 = sub(1, 0, 3, 1), s23 = sub(1, 0, 2, 1);
        const double f0[4] = {s00, s00, s02, s03}, f1[4] = {s04, s04, s06, s07}, f2[4] = {s08, s08, s10, s11};
        const double f3[4] = {s12, s12, s14, s15}, f4[4] = {s16, s16, s18, s19}, f5[4] = {s20, s20, s22, s23};
        const double v0[4] = {a[1][0], a[0][0], a[0][0], a[0][0]}, v1[4] = {a[1][1], a[0][1], a[0][1], a[0][1]};
        const double v2[4] = {a[1][2], a[0][2], a[0][2], a[0][2]}, v3[4] = {a[1][3], a[0][3], a[0][3], a[0][3]};
        Mat4 inv;
        for (int i = 0; i < 4; ++i) {
            double sa = (i & 1) ? -1.0 : 1.0, sb = -sa;
            inv.m[0][i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa;
            inv.m[1][i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb;
            inv.m[2][i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa;
            inv.m[3][i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb;
        }
        double det = ((a[0][0] * inv.m[0][0] + a[0][1] * inv.m[1][0]) + a[0][2] * inv.m[2][0]) + a[0][3] * inv.m[3][0];
        double rcp = 1.0 / det;
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) inv.m[c][r] *= rcp;
        return inv;
    }
};

}  // namespace nrrt_host
