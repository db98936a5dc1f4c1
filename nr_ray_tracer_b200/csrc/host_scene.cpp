// host_scene.cpp — host side of the drop-in: object graph -> reference BVH -> flat device layout.
//
// north_star keeps "the TOML/JSON scene loader and the BVH build" on the host.  This file is the
// C++ stand-in for that Rust host code (no Rust toolchain exists in the build image):
//   * per-object bounding boxes exactly as the reference computes them
//       AABB::new/pad_to_minimums/union/from_points   aabb.rs:16-76
//       SphereBuilder::build                           objects/sphere.rs:69-91
//       PlaneBuilder::build (bbox, normal, d, w)       objects/plane.rs:95-127
//       Translate::new / rotate_bbox / scale_bbox      objects/translate.rs:18-29, rotate.rs:13-62, scale.rs:10-58
//   * BVH::from — median split on the longest axis, stable sort by bbox.min[axis] with total_cmp,
//     one object per leaf (objects/object.rs:41-73).  The tree shape is part of the parity contract
//     (leaf order decides equal-t ties, object.rs:110-114).
//   * flattening into the layout of include/nrrt.h: index-based, no pointers, built per "space"
//     (everything reachable without crossing a transform wrapper); wrapper chains become instances
//     whose inner spaces are shared between instances of the same group.
//   * CameraBuilder::build (camera.rs:94-159).
//
// Unlike the reference (and the oracle) there are no hit() methods here: intersection lives in the
// CUDA kernels only.  Arithmetic is plain f64 in the reference's operation order; compile with
// -ffp-contract=off / -fmad=false so nothing is fused.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/nrrt.h"
#include "host_math.hpp"

namespace nrrt_host {

static thread_local std::string g_error;

struct HostError {
    std::string msg;
};
[[noreturn]] static void fail(const std::string& m) { throw HostError{m}; }

// ------------------------------------------------------------------ per-object data
struct ObjInfo {
    bool done = false;
    Aabb box;
    // GROUP: index-based BVH (object.rs:9-16)
    struct BNode {
        bool leaf;
        int object;  // leaf: graph object index or -1 (Leaf(None))
        Aabb box;    // node bbox (inner only)
        int left, right;
    };
    std::vector<BNode> bvh;
    int bvh_root = -1;
};

struct Flattener {
    const nrrt_graph_desc& g;
    std::vector<ObjInfo> info;

    // output arrays
    std::vector<nrrt_node> nodes;
    std::vector<nrrt_box> child_boxes;
    std::vector<double> sphere_rec;  // 4 doubles per sphere
    std::vector<double> sphere_speed;  // 3 doubles per sphere
    bool any_motion = false;
    std::vector<uint32_t> sphere_material, sphere_order, sphere_object;
    std::vector<double> plane_rec;   // 16 doubles per plane
    std::vector<uint32_t> plane_material, plane_order, plane_object;
    std::vector<nrrt_instance> instances;
    std::vector<uint32_t> instance_order;
    std::vector<nrrt_xform> xforms;

    struct SpaceRoot {
        uint32_t ref;
        Aabb box;
        uint32_t depth;   // worst-case traversal stack need below this root
        uint32_t levels;  // instance nesting levels below this root
    };
    std::map<uint32_t, SpaceRoot> space_memo;  // group object -> emitted space

    // NRRT_BUILD_SAH (§8(f) N3): inner nodes come from a binned surface-area-heuristic build instead of the
    // reference's median split.  Leaves, their records and their `order` numbers (the reference's depth-first leaf
    // order, which decides equal-t ties) are emitted exactly as in the reference tree; only the inner nodes differ.
    bool sah = false;

    Flattener(const nrrt_graph_desc& graph, bool use_sah) : g(graph), info(graph.n_objects), sah(use_sah) {}

    const nrrt_object& obj(uint32_t i) const {
        if (i >= g.n_objects) fail("object index out of range");
        return g.objects[i];
    }
    uint32_t child_of(const nrrt_object& o, uint32_t k) const {
        if (k >= o.n_children || o.first_child + k >= g.n_child_ids) fail("child index out of range");
        return g.child_ids[o.first_child + k];
    }

    // ---- bounding boxes + BVH build (recursive over the DAG, memoised)
    void prepare(uint32_t i, int depth) {
        if (depth > 64) fail("object nesting too deep (cycle?)");
        ObjInfo& in = info[i];
        if (in.done) return;
        const nrrt_object& o = obj(i);
        switch (o.kind) {
            case NRRT_OBJ_SPHERE: {
                if (o.material >= g.n_materials) fail("material index out of range");
                D3 c{o.v[0], o.v[1], o.v[2]};
                D3 r{o.v[3], o.v[3], o.v[3]};
                D3 c1 = c + D3{o.v[4], o.v[5], o.v[6]};  // center + speed.unwrap_or(ZERO), sphere.rs:75-76
                Aabb b0 = Aabb::from_points(c - r, c + r);
                Aabb b1 = Aabb::from_points(c1 - r, c1 + r);
                in.box = b0.unite(b1);
                break;
            }
            case NRRT_OBJ_QUAD:
            case NRRT_OBJ_TRIANGLE: {
                if (o.material >= g.n_materials) fail("material index out of range");
                D3 p{o.v[0], o.v[1], o.v[2]}, u{o.v[3], o.v[4], o.v[5]}, v{o.v[6], o.v[7], o.v[8]};
                Aabb b0 = Aabb::from_points(p, p + u + v);
                Aabb b1 = Aabb::from_points(p + u, p + v);
                in.box = b0.unite(b1);
                break;
            }
            case NRRT_OBJ_GROUP: {
                std::vector<int> kids;
                for (uint32_t k = 0; k < o.n_children; ++k) {
                    uint32_t c = child_of(o, k);
                    prepare(c, depth + 1);
                    kids.push_back((int)c);
                }
                in.bvh_root = build_bvh(in, kids.data(), kids.size());
                in.box = bvh_box(in, in.bvh_root);
                break;
            }
            case NRRT_OBJ_TRANSLATE: {
                uint32_t c = child_of(o, 0);
                prepare(c, depth + 1);
                in.box = info[c].box.translated(D3{o.v[0], o.v[1], o.v[2]});
                break;
            }
            case NRRT_OBJ_ROTATE_X:
            case NRRT_OBJ_ROTATE_Y:
            case NRRT_OBJ_ROTATE_Z: {
                uint32_t c = child_of(o, 0);
                prepare(c, depth + 1);
                Mat3 inv = Mat3::from_axis_angle(axis_of(o.kind), o.v[0]);
                in.box = corners_box(info[c].box, [&](D3 p) { return inv.mul(p); });
                break;
            }
            case NRRT_OBJ_SCALE: {
                uint32_t c = child_of(o, 0);
                prepare(c, depth + 1);
                Mat4 m = Mat4::from_scale(D3{o.v[0], o.v[1], o.v[2]});
                in.box = corners_box(info[c].box, [&](D3 p) { return m.transform_point3(p); });
                break;
            }
            default: fail("unknown object kind");
        }
        in.done = true;
    }

    static D3 axis_of(uint32_t kind) {
        return kind == NRRT_OBJ_ROTATE_X ? D3{1, 0, 0} : (kind == NRRT_OBJ_ROTATE_Y ? D3{0, 1, 0} : D3{0, 0, 1});
    }

    template <class F>
    static Aabb corners_box(const Aabb& b, F&& xf) {  // rotate.rs:13-36 / scale.rs:10-33
        const double inf = std::numeric_limits<double>::infinity();
        D3 mn{inf, inf, inf}, mx{-inf, -inf, -inf};
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) {
                    double x = (double)i * b.hi[0] + (1.0 - (double)i) * b.lo[0];
                    double y = (double)j * b.hi[1] + (1.0 - (double)j) * b.lo[1];
                    double z = (double)k * b.hi[2] + (1.0 - (double)k) * b.lo[2];
                    D3 t = xf(D3{x, y, z});
                    mn = D3{std::fmin(mn.x, t.x), std::fmin(mn.y, t.y), std::fmin(mn.z, t.z)};
                    mx = D3{std::fmax(mx.x, t.x), std::fmax(mx.y, t.y), std::fmax(mx.z, t.z)};
                }
        return Aabb::from_points(mn, mx);
    }

    Aabb bvh_box(const ObjInfo& in, int n) const {  // BVH::bbox, object.rs:77-87
        const ObjInfo::BNode& b = in.bvh[n];
        if (!b.leaf) return b.box;
        if (b.object >= 0) return info[b.object].box;
        return Aabb::empty();
    }

    int build_bvh(ObjInfo& in, int* objs, size_t n) {  // BVH::from, object.rs:41-73
        auto leaf = [&](int o) {
            in.bvh.push_back(ObjInfo::BNode{true, o, Aabb::empty(), -1, -1});
            return (int)in.bvh.size() - 1;
        };
        if (n == 0) return leaf(-1);
        if (n == 1) return leaf(objs[0]);
        if (n == 2) {
            int l = leaf(objs[0]), r = leaf(objs[1]);
            Aabb box = info[objs[0]].box.unite(info[objs[1]].box);
            in.bvh.push_back(ObjInfo::BNode{false, -1, box, l, r});
            return (int)in.bvh.size() - 1;
        }
        Aabb box = Aabb::empty();
        for (size_t i = 0; i < n; ++i) box = box.unite(info[objs[i]].box);
        int axis = box.longest_axis();
        std::stable_sort(objs, objs + n, [&](int a, int b) {
            return total_cmp(info[a].box.lo[axis], info[b].box.lo[axis]) < 0;
        });
        size_t mid = n / 2;
        int l = build_bvh(in, objs, mid);
        int r = build_bvh(in, objs + mid, n - mid);
        in.bvh.push_back(ObjInfo::BNode{false, -1, box, l, r});
        return (int)in.bvh.size() - 1;
    }

    // ---- flattening
    struct Emitted {
        uint32_t ref;
        Aabb box;        // box to store in the parent (exact for NODE, conservative for leaves)
        uint32_t depth;   // stack entries needed to traverse this subtree
        uint32_t levels;  // instance nesting levels inside this subtree
    };

    static nrrt_box to_box(const Aabb& b) {
        nrrt_box o;
        for (int a = 0; a < 3; ++a) o.lo[a] = b.lo[a], o.hi[a] = b.hi[a];
        return o;
    }

    // A space = one coordinate frame; `order` counts its leaves in DFS order.
    struct SpaceCtx {
        uint32_t order = 0;
    };

    // Hitable reached through a BVH leaf / wrapper: primitive, group (inlined) or wrapper chain.
    Emitted emit_hitable(uint32_t oi, SpaceCtx& sp) {
        const nrrt_object& o = obj(oi);
        switch (o.kind) {
            case NRRT_OBJ_SPHERE: {
                uint32_t idx = (uint32_t)sphere_material.size();
                for (int k = 0; k < 4; ++k) sphere_rec.push_back(o.v[k]);
                for (int k = 4; k < 7; ++k) {
                    sphere_speed.push_back(o.v[k]);
                    if (o.v[k] != 0.0) any_motion = true;  // Some(ZERO) behaves exactly like None
                }
                sphere_material.push_back(o.material);
                sphere_order.push_back(sp.order++);
                sphere_object.push_back(oi);
                return Emitted{NRRT_REF(NRRT_REF_SPHERE, idx), info[oi].box, 0, 0};
            }
            case NRRT_OBJ_QUAD:
            case NRRT_OBJ_TRIANGLE: {
                uint32_t idx = (uint32_t)plane_material.size();
                D3 p{o.v[0], o.v[1], o.v[2]}, u{o.v[3], o.v[4], o.v[5]}, v{o.v[6], o.v[7], o.v[8]};
                D3 n = u.cross(v);           // plane.rs:109
                D3 normal = n.normalize();   // :111
                double d = normal.dot(p);    // :113
                D3 w = n / n.dot(n);         // :114
                push3(plane_rec, normal);
                plane_rec.push_back(d);
                push3(plane_rec, p), push3(plane_rec, w), push3(plane_rec, u), push3(plane_rec, v);
                plane_material.push_back(o.material | (o.kind == NRRT_OBJ_TRIANGLE ? NRRT_PLANE_TRIANGLE_BIT : 0u));
                plane_order.push_back(sp.order++);
                plane_object.push_back(oi);
                return Emitted{NRRT_REF(NRRT_REF_PLANE, idx), info[oi].box, 0, 0};
            }
            case NRRT_OBJ_GROUP: return emit_bvh(oi, info[oi].bvh_root, sp);  // nested BVH in the same space: inlined
            default: return emit_instance(oi, sp);
        }
    }

    static void push3(std::vector<double>& v, D3 a) {
        v.push_back(a.x), v.push_back(a.y), v.push_back(a.z);
    }

    // ---- NRRT_BUILD_SAH: leaves in reference DFS order, then a SAH tree over them
    void collect_leaves(uint32_t group, int n, SpaceCtx& sp, std::vector<Emitted>& out) {
        const ObjInfo::BNode b = info[group].bvh[n];
        if (b.leaf) {
            if (b.object >= 0) out.push_back(emit_hitable((uint32_t)b.object, sp));  // assigns sp.order in DFS order
            return;
        }
        collect_leaves(group, b.left, sp, out);
        collect_leaves(group, b.right, sp, out);
    }
    static double half_area(const Aabb& b) {
        double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    static uint32_t ceil_log2(size_t n) {
        uint32_t k = 0;
        while (((size_t)1 << k) < n) ++k;
        return k;
    }
    // Builds nodes over items[lo, hi) (hi - lo >= 2); depth_left bounds the subtree height so the traversal stack
    // (NRRT_STACK_CAP, validated through max_stack) cannot overflow on adversarial inputs: when the budget gets
    // tight the split falls back to the balanced median.
    Emitted build_sah(std::vector<Emitted>& items, size_t lo, size_t hi, uint32_t depth_left) {
        const size_t n = hi - lo;
        Aabb box = items[lo].box, cbox = Aabb::empty();
        for (size_t i = lo; i < hi; ++i) {
            if (i > lo) box = box.unite(items[i].box);
            for (int a = 0; a < 3; ++a) {
                double c = 0.5 * (items[i].box.lo[a] + items[i].box.hi[a]);
                cbox.lo[a] = std::fmin(cbox.lo[a], c), cbox.hi[a] = std::fmax(cbox.hi[a], c);
            }
        }
        size_t mid = lo + n / 2;
        int axis = 0;
        for (int a = 1; a < 3; ++a)
            if (cbox.hi[a] - cbox.lo[a] > cbox.hi[axis] - cbox.lo[axis]) axis = a;
        auto centroid = [&](const Emitted& e, int a) { return 0.5 * (e.box.lo[a] + e.box.hi[a]); };
        bool finite = true;
        for (int a = 0; a < 3; ++a) finite = finite && std::isfinite(cbox.lo[a]) && std::isfinite(cbox.hi[a]);
        bool median = n <= 2 || ceil_log2(n) + 1 >= depth_left || !finite || !(cbox.hi[axis] - cbox.lo[axis] > 0.0);
        constexpr int NB = 16;
        auto bin_of = [&](const Emitted& e, int a) {  // NaN / out-of-range centroids (malformed input) land in bin 0
            const double x = (centroid(e, a) - cbox.lo[a]) / (cbox.hi[a] - cbox.lo[a]) * NB;
            return (x >= 0.0 && x < (double)NB) ? (int)x : (x >= (double)NB ? NB - 1 : 0);
        };
        if (!median) {
            double best_cost = std::numeric_limits<double>::infinity();
            int best_axis = -1, best_bin = 0;
            for (int a = 0; a < 3; ++a) {
                if (!(cbox.hi[a] - cbox.lo[a] > 0.0)) continue;
                Aabb bb[NB];
                size_t cnt[NB] = {0};
                for (int k = 0; k < NB; ++k) bb[k] = Aabb::empty();
                for (size_t i = lo; i < hi; ++i) {
                    int k = bin_of(items[i], a);
                    for (int x = 0; x < 3; ++x) {
                        bb[k].lo[x] = std::fmin(bb[k].lo[x], items[i].box.lo[x]);
                        bb[k].hi[x] = std::fmax(bb[k].hi[x], items[i].box.hi[x]);
                    }
                    cnt[k]++;
                }
                double right_area[NB];
                size_t right_cnt[NB];
                Aabb acc = Aabb::empty();
                size_t c = 0;
                for (int k = NB - 1; k > 0; --k) {
                    if (cnt[k])
                        for (int x = 0; x < 3; ++x)
                            acc.lo[x] = std::fmin(acc.lo[x], bb[k].lo[x]), acc.hi[x] = std::fmax(acc.hi[x], bb[k].hi[x]);
                    c += cnt[k];
                    right_area[k] = c ? half_area(acc) : 0.0;
                    right_cnt[k] = c;
                }
                acc = Aabb::empty();
                c = 0;
                for (int k = 0; k < NB - 1; ++k) {
                    if (cnt[k])
                        for (int x = 0; x < 3; ++x)
                            acc.lo[x] = std::fmin(acc.lo[x], bb[k].lo[x]), acc.hi[x] = std::fmax(acc.hi[x], bb[k].hi[x]);
                    c += cnt[k];
                    if (c == 0 || right_cnt[k + 1] == 0) continue;
                    double cost = half_area(acc) * (double)c + right_area[k + 1] * (double)right_cnt[k + 1];
                    if (cost < best_cost) best_cost = cost, best_axis = a, best_bin = k;
                }
            }
            if (best_axis < 0) {
                median = true;
            } else {
                auto first_right = std::stable_partition(items.begin() + lo, items.begin() + hi, [&](const Emitted& e) {
                    return bin_of(e, best_axis) <= best_bin;
                });
                mid = (size_t)(first_right - items.begin());
                if (mid == lo || mid == hi) median = true;
            }
        }
        if (median) {
            if (finite)
                std::stable_sort(items.begin() + lo, items.begin() + hi, [&](const Emitted& x, const Emitted& y) {
                    return centroid(x, axis) < centroid(y, axis);
                });
            mid = lo + n / 2;
        }
        uint32_t idx = (uint32_t)nodes.size();
        if (idx >= NRRT_REF_INDEX_MASK) fail("too many BVH nodes");
        nodes.emplace_back();
        child_boxes.emplace_back();
        child_boxes.emplace_back();
        const uint32_t dl = depth_left > 1 ? depth_left - 1 : 1;
        Emitted l = (mid - lo == 1) ? items[lo] : build_sah(items, lo, mid, dl);
        Emitted r = (hi - mid == 1) ? items[mid] : build_sah(items, mid, hi, dl);
        write_node(idx, l, r);
        return Emitted{NRRT_REF(NRRT_REF_NODE, idx), box, 1 + std::max(l.depth, r.depth), std::max(l.levels, r.levels)};
    }
    void write_node(uint32_t idx, const Emitted& l, const Emitted& r) {
        nrrt_node nd;
        std::memset(&nd, 0, sizeof nd);
        const Emitted* ch[2] = {&l, &r};
        for (int c = 0; c < 2; ++c) {
            nd.child[c] = ch[c]->ref;
            for (int a = 0; a < 3; ++a) {
                nd.lo[c][a] = (float)ch[c]->box.lo[a];
                nd.hi[c][a] = (float)ch[c]->box.hi[a];
            }
            child_boxes[2 * (size_t)idx + c] = to_box(ch[c]->box);
        }
        nodes[idx] = nd;
    }
    Emitted emit_group_sah(uint32_t group, SpaceCtx& sp) {
        std::vector<Emitted> items;
        collect_leaves(group, info[group].bvh_root, sp, items);
        if (items.empty()) return Emitted{NRRT_REF_NONE, Aabb::empty(), 0, 0};
        if (items.size() == 1) return items[0];
        return build_sah(items, 0, items.size(), ceil_log2(items.size()) + 7);
    }

    Emitted emit_bvh(uint32_t group, int n, SpaceCtx& sp) {
        const ObjInfo& in = info[group];
        if (sah && n == in.bvh_root && !in.bvh[n].leaf) return emit_group_sah(group, sp);
        const ObjInfo::BNode b = in.bvh[n];
        if (b.leaf) {
            if (b.object < 0) return Emitted{NRRT_REF_NONE, Aabb::empty(), 0, 0};
            return emit_hitable((uint32_t)b.object, sp);  // Leaf(Some(o)) forwards, no box test (object.rs:95-97)
        }
        uint32_t idx = (uint32_t)nodes.size();
        if (idx >= NRRT_REF_INDEX_MASK) fail("too many BVH nodes");
        nodes.emplace_back();
        child_boxes.emplace_back();
        child_boxes.emplace_back();
        Emitted l = emit_bvh(group, b.left, sp);   // left first: DFS leaf order
        Emitted r = emit_bvh(group, b.right, sp);
        write_node(idx, l, r);
        return Emitted{NRRT_REF(NRRT_REF_NODE, idx), b.box, 1 + std::max(l.depth, r.depth), std::max(l.levels, r.levels)};
    }

    // The shared inner space of a group used behind wrappers.
    SpaceRoot emit_space(uint32_t group) {
        auto it = space_memo.find(group);
        if (it != space_memo.end()) return it->second;
        SpaceCtx sp;
        Emitted e = emit_bvh(group, info[group].bvh_root, sp);
        SpaceRoot r{e.ref, e.box, e.depth, e.levels};
        space_memo[group] = r;
        return r;
    }

    Emitted emit_instance(uint32_t oi, SpaceCtx& sp) {
        uint32_t my_order = sp.order++;
        uint32_t first = (uint32_t)xforms.size();
        uint32_t cur = oi;
        // walk the wrapper chain, outermost first; single-object groups forward (Leaf(Some(o)))
        for (;;) {
            const nrrt_object& o = obj(cur);
            if (o.kind == NRRT_OBJ_GROUP) {
                const ObjInfo& in = info[cur];
                const ObjInfo::BNode& b = in.bvh[in.bvh_root];
                if (b.leaf && b.object >= 0) {
                    cur = (uint32_t)b.object;
                    continue;
                }
                break;
            }
            if (o.kind < NRRT_OBJ_TRANSLATE) break;
            xforms.push_back(make_xform(o));
            cur = child_of(o, 0);
        }
        uint32_t nx = (uint32_t)xforms.size() - first;
        nrrt_instance ins;
        std::memset(&ins, 0, sizeof ins);
        ins.first_xform = first;
        ins.n_xforms = nx;
        uint32_t depth = 0, levels = 0;
        const nrrt_object& io = obj(cur);
        if (io.kind == NRRT_OBJ_GROUP) {
            const ObjInfo& in = info[cur];
            const ObjInfo::BNode& b = in.bvh[in.bvh_root];
            if (b.leaf) {  // Leaf(None)
                ins.inner = NRRT_REF_NONE;
                ins.inner_box = to_box(Aabb::empty());
            } else {
                SpaceRoot r = emit_space(cur);
                ins.inner = r.ref;
                ins.inner_box = to_box(r.box);
                depth = r.depth;
                levels = r.levels;
            }
        } else {  // a bare primitive behind wrappers: its own one-leaf space
            SpaceCtx inner;
            Emitted e = emit_hitable(cur, inner);
            ins.inner = e.ref;
            ins.inner_box = to_box(e.box);
        }
        uint32_t idx = (uint32_t)instances.size();
        instances.push_back(ins);
        instance_order.push_back(my_order);
        return Emitted{NRRT_REF(NRRT_REF_INSTANCE, idx), info[oi].box, depth + 1, levels + 1};
    }

    // ---- four-slot nodes (nrrt_wnode): every other level of the binary trees folded away.  Runs after the binary
    // flattening; binary node b becomes wide node wide_of[b] whose slots are b's grandchildren (or b's child where
    // that child is a leaf), in depth-first order.  Only nodes reachable as a space root or as a slot are emitted.
    std::vector<nrrt_wnode> wnodes;
    std::vector<nrrt_box> wide_boxes;
    std::vector<uint32_t> wide_of;              // binary node -> wide node, 0xFFFFFFFF = not (yet) emitted
    std::vector<uint32_t> instance_wide_inner;  // per instance
    std::vector<uint32_t> wide_need;            // per wide node: traversal stack entries needed below it

    uint32_t wide_ref(uint32_t ref) {  // binary ref -> wide ref (emits the subtree on first use)
        if (ref == NRRT_REF_NONE || NRRT_REF_TYPE(ref) != NRRT_REF_NODE) return ref;
        return NRRT_REF(NRRT_REF_NODE, emit_wide(NRRT_REF_INDEX(ref)));
    }
    uint32_t emit_wide(uint32_t b) {
        if (wide_of[b] != 0xFFFFFFFFu) return wide_of[b];
        const uint32_t w = (uint32_t)wnodes.size();
        if (w >= NRRT_REF_INDEX_MASK) fail("too many BVH nodes");
        wide_of[b] = w;
        wnodes.emplace_back();
        wide_need.push_back(0);
        for (int k = 0; k < 8; ++k) wide_boxes.push_back(to_box(Aabb::empty()));
        struct Slot {
            uint32_t ref;       // binary ref
            nrrt_box box;       // the slot's own box
            nrrt_box gate;      // box of the folded-away parent
            bool gated;
        };
        Slot slots[4];
        int n = 0;
        const nrrt_node nb = nodes[b];
        for (int c = 0; c < 2; ++c) {
            const uint32_t r = nb.child[c];
            if (r == NRRT_REF_NONE) continue;
            if (NRRT_REF_TYPE(r) == NRRT_REF_NODE) {
                const uint32_t b2 = NRRT_REF_INDEX(r);
                const nrrt_node n2 = nodes[b2];
                for (int c2 = 0; c2 < 2; ++c2) {
                    if (n2.child[c2] == NRRT_REF_NONE) continue;
                    slots[n++] = Slot{n2.child[c2], child_boxes[2 * (size_t)b2 + c2], child_boxes[2 * (size_t)b + c], true};
                }
            } else {
                slots[n++] = Slot{r, child_boxes[2 * (size_t)b + c], to_box(Aabb::empty()), false};
            }
        }
        nrrt_wnode wn;
        std::memset(&wn, 0, sizeof wn);
        const float inf = std::numeric_limits<float>::infinity();
        uint32_t need = 0;
        for (int s = 0; s < 4; ++s) {
            if (s >= n) {  // unused slot: an empty box and no child
                for (int a = 0; a < 3; ++a) wn.lo[a][s] = inf, wn.hi[a][s] = -inf;
                wn.child[s] = NRRT_REF_NONE;
                continue;
            }
            for (int a = 0; a < 3; ++a) wn.lo[a][s] = (float)slots[s].box.lo[a], wn.hi[a][s] = (float)slots[s].box.hi[a];
            wn.meta[s] = slots[s].gated ? NRRT_WNODE_GATED : 0u;
            wide_boxes[8 * (size_t)w + 2 * s] = slots[s].box;
            wide_boxes[8 * (size_t)w + 2 * s + 1] = slots[s].gate;
            const uint32_t wr = wide_ref(slots[s].ref);  // recursion: depth-first, so slot order = leaf order
            wn.child[s] = wr;
            uint32_t below = 0;
            if (NRRT_REF_TYPE(wr) == NRRT_REF_NODE) below = wide_need[NRRT_REF_INDEX(wr)];
            else if (NRRT_REF_TYPE(wr) == NRRT_REF_INSTANCE) below = instance_need(NRRT_REF_INDEX(wr));
            need = std::max(need, below);
        }
        // all but the slot being descended into wait on the stack
        wide_need[w] = need + (n > 0 ? (uint32_t)n - 1 : 0);
        wnodes[w] = wn;
        return w;
    }
    // entering an instance leaves a level marker on the stack, then traverses the nested space
    uint32_t instance_need(uint32_t i) {
        const uint32_t wr = wide_ref(instances[i].inner);
        instance_wide_inner[i] = wr;
        uint32_t below = 0;
        if (wr != NRRT_REF_NONE && NRRT_REF_TYPE(wr) == NRRT_REF_NODE) below = wide_need[NRRT_REF_INDEX(wr)];
        else if (wr != NRRT_REF_NONE && NRRT_REF_TYPE(wr) == NRRT_REF_INSTANCE) below = instance_need(NRRT_REF_INDEX(wr));
        return below + 1;
    }
    // returns the wide root ref and the worst-case stack need of the whole scene
    uint32_t build_wide(uint32_t root_ref, uint32_t& max_stack) {
        wide_of.assign(nodes.size(), 0xFFFFFFFFu);
        instance_wide_inner.assign(instances.size(), NRRT_REF_NONE);
        const uint32_t wr = wide_ref(root_ref);
        uint32_t need = 0;
        if (wr != NRRT_REF_NONE && NRRT_REF_TYPE(wr) == NRRT_REF_NODE) need = wide_need[NRRT_REF_INDEX(wr)];
        else if (wr != NRRT_REF_NONE && NRRT_REF_TYPE(wr) == NRRT_REF_INSTANCE) need = instance_need(NRRT_REF_INDEX(wr));
        for (uint32_t i = 0; i < instances.size(); ++i)  // instances the root never reaches cannot exist, but be safe
            if (instance_wide_inner[i] == NRRT_REF_NONE && instances[i].inner != NRRT_REF_NONE) instance_need(i);
        max_stack = need + 2;
        return wr;
    }

    nrrt_xform make_xform(const nrrt_object& o) const {
        nrrt_xform x;
        std::memset(&x, 0, sizeof x);
        if (o.kind == NRRT_OBJ_TRANSLATE) {
            x.kind = NRRT_XF_TRANSLATE;
            for (int k = 0; k < 3; ++k) x.to_obj[k] = o.v[k];
        } else if (o.kind == NRRT_OBJ_SCALE) {
            x.kind = NRRT_XF_SCALE;
            Mat4 m = Mat4::from_scale(D3{o.v[0], o.v[1], o.v[2]});  // scale.rs:48
            Mat4 inv = m.inverse();                                   // scale.rs:49
            for (int c = 0; c < 4; ++c)
                for (int r = 0; r < 3; ++r) {
                    x.to_obj[3 * c + r] = inv.m[c][r];
                    x.to_world[3 * c + r] = m.m[c][r];
                }
        } else {
            x.kind = NRRT_XF_ROTATE;
            D3 axis = axis_of(o.kind);
            Mat3 rot = Mat3::from_axis_angle(axis, -o.v[0]);  // rotate.rs:52
            Mat3 inv = Mat3::from_axis_angle(axis, o.v[0]);   // rotate.rs:53
            const D3* rc[3] = {&rot.c0, &rot.c1, &rot.c2};
            const D3* ic[3] = {&inv.c0, &inv.c1, &inv.c2};
            for (int c = 0; c < 3; ++c) {
                x.to_obj[3 * c + 0] = rc[c]->x, x.to_obj[3 * c + 1] = rc[c]->y, x.to_obj[3 * c + 2] = rc[c]->z;
                x.to_world[3 * c + 0] = ic[c]->x, x.to_world[3 * c + 1] = ic[c]->y, x.to_world[3 * c + 2] = ic[c]->z;
            }
        }
        return x;
    }
};

}  // namespace nrrt_host

// ================================================================== C ABI
using namespace nrrt_host;

struct nrrt_host_scene {
    std::unique_ptr<Flattener> f;
    std::vector<nrrt_material> materials;
    std::vector<nrrt_texture> textures;
    std::vector<nrrt_image> images;
    std::vector<std::vector<uint8_t>> image_data;
    nrrt_scene_desc desc;
};

extern "C" {

const char* nrrt_host_last_error(void) { return g_error.c_str(); }

nrrt_host_scene* nrrt_host_build(const nrrt_graph_desc* g) { return nrrt_host_build_ex(g, NRRT_BUILD_REFERENCE); }

nrrt_host_scene* nrrt_host_build_ex(const nrrt_graph_desc* g, uint32_t flags) {
    g_error.clear();
    if (flags & ~(uint32_t)NRRT_BUILD_SAH) {
        g_error = "nrrt_host_build_ex: unknown flags";
        return nullptr;
    }
    if (!g || !g->objects || g->root >= g->n_objects) {
        g_error = "nrrt_host_build: null graph or bad root";
        return nullptr;
    }
    try {
        if (g->objects[g->root].kind != NRRT_OBJ_GROUP) fail("root object must be a GROUP (the scene list)");
        for (uint32_t i = 0; i < g->n_materials; ++i) {
            const nrrt_material& m = g->materials[i];
            if (m.kind > NRRT_MAT_DIFFUSE_LIGHT) fail("unknown material kind");
            if (m.kind != NRRT_MAT_DIELECTRIC && m.texture >= g->n_textures) fail("texture index out of range");
        }
        for (uint32_t i = 0; i < g->n_textures; ++i) {
            const nrrt_texture& t = g->textures[i];
            if (t.kind > NRRT_TEX_MARBLE) fail("unknown texture kind");
            if (t.kind == NRRT_TEX_CHECKER && (t.a >= g->n_textures || t.b >= g->n_textures))
                fail("checker sub-texture out of range");
            if (t.kind == NRRT_TEX_CHECKER && (t.a >= i || t.b >= i))
                fail("checker sub-textures must precede the checker (scene_config.rs:65-81)");
            if (t.kind == NRRT_TEX_IMAGE && t.a >= g->n_images) fail("image index out of range");
        }
        auto hs = std::make_unique<nrrt_host_scene>();
        hs->f = std::make_unique<Flattener>(*g, (flags & NRRT_BUILD_SAH) != 0);
        Flattener& f = *hs->f;
        f.prepare(g->root, 0);
        Flattener::SpaceCtx world;
        Flattener::Emitted root = f.emit_bvh(g->root, f.info[g->root].bvh_root, world);
        if (root.levels > NRRT_MAX_INSTANCE_DEPTH) fail("instance nesting deeper than NRRT_MAX_INSTANCE_DEPTH");

        hs->materials.assign(g->materials, g->materials + g->n_materials);
        hs->textures.assign(g->textures, g->textures + g->n_textures);
        for (uint32_t i = 0; i < g->n_images; ++i) {
            const nrrt_image& im = g->images[i];
            if (!im.rgb || im.width == 0 || im.height == 0) fail("empty image");
            size_t n = (size_t)im.width * im.height * 3;
            hs->image_data.emplace_back(im.rgb, im.rgb + n);
        }
        for (uint32_t i = 0; i < g->n_images; ++i)
            hs->images.push_back(nrrt_image{g->images[i].width, g->images[i].height, hs->image_data[i].data()});

        nrrt_scene_desc& d = hs->desc;
        std::memset(&d, 0, sizeof d);
        d.abi_version = NRRT_ABI_VERSION;
        d.n_nodes = (uint32_t)f.nodes.size();
        d.nodes = f.nodes.data();
        d.child_boxes = f.child_boxes.data();
        d.root = root.ref;
        d.root_box = Flattener::to_box(root.box);
        d.n_spheres = (uint32_t)f.sphere_material.size();
        d.sphere_rec = f.sphere_rec.data();
        d.sphere_material = f.sphere_material.data();
        d.sphere_order = f.sphere_order.data();
        d.sphere_object = f.sphere_object.data();
        d.n_planes = (uint32_t)f.plane_material.size();
        d.plane_rec = f.plane_rec.data();
        d.plane_material = f.plane_material.data();
        d.plane_order = f.plane_order.data();
        d.plane_object = f.plane_object.data();
        d.n_instances = (uint32_t)f.instances.size();
        d.instances = f.instances.data();
        d.instance_order = f.instance_order.data();
        d.n_xforms = (uint32_t)f.xforms.size();
        d.xforms = f.xforms.data();
        d.n_materials = (uint32_t)hs->materials.size();
        d.materials = hs->materials.data();
        d.n_textures = (uint32_t)hs->textures.size();
        d.textures = hs->textures.data();
        d.n_images = (uint32_t)hs->images.size();
        d.images = hs->images.data();
        d.sphere_speed = f.any_motion ? f.sphere_speed.data() : nullptr;
        uint32_t wide_stack = 0;
        d.wide_root = f.build_wide(root.ref, wide_stack);
        d.max_stack = wide_stack;
        d.n_wnodes = (uint32_t)f.wnodes.size();
        d.wnodes = f.wnodes.data();
        d.wide_boxes = f.wide_boxes.data();
        d.instance_wide_inner = f.instance_wide_inner.data();
        return hs.release();
    } catch (const HostError& e) {
        g_error = e.msg;
    } catch (const std::exception& e) {
        g_error = e.what();
    }
    return nullptr;
}

const nrrt_scene_desc* nrrt_host_scene_desc(const nrrt_host_scene* s) { return s ? &s->desc : nullptr; }

void nrrt_host_free(nrrt_host_scene* s) { delete s; }

int nrrt_host_camera_build(const nrrt_camera_config* c, nrrt_camera* out) {  // camera.rs:94-159
    if (!c || !out || c->width == 0 || c->height == 0) return NRRT_ERR_INVALID;
    const double PI = 3.14159265358979323846264338327950288;
    std::memset(out, 0, sizeof *out);
    out->width = c->width;
    out->height = c->height;
    out->ray_max_bounces = c->ray_max_bounces;
    out->samples_per_pixel = std::max<uint32_t>(c->samples_per_pixel, 1u);       // :104
    double defocus_angle = std::min(std::max(c->defocus_angle, 0.0), PI);         // :106
    double h = std::tan(c->field_of_view / 2.);                                   // :110
    double vh = c->focus_dist * h * 2.0;                                          // :112
    double vw = vh * ((double)c->width / (double)c->height);                      // :113
    D3 from{c->look_from[0], c->look_from[1], c->look_from[2]};
    D3 at{c->look_at[0], c->look_at[1], c->look_at[2]};
    D3 up{c->view_up[0], c->view_up[1], c->view_up[2]};
    D3 w = (from - at).normalize();                                               // :115
    D3 u = up.cross(w).normalize();                                               // :116
    D3 v = w.cross(u).normalize();                                                // :117
    D3 vu = u * vw;                                                               // :119
    D3 vv = (-v) * vh;                                                            // :120
    D3 du = vu / (double)c->width;                                                // :122
    D3 dv = vv / (double)c->height;                                               // :123
    D3 tl = from - w * c->focus_dist - vu / 2.0 - vv / 2.0 + (du + dv) / 2.0;     // :125-131
    double radius = c->focus_dist * std::tan(defocus_angle / 2.0);                // :133
    D3 ddu = u * radius, ddv = v * radius;                                        // :134-135
    auto put = [](double* d, D3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
    for (int k = 0; k < 3; ++k) out->background[k] = c->background[k];
    put(out->look_from, from);
    put(out->defocus_disk_u, ddu);
    put(out->defocus_disk_v, ddv);
    put(out->pixel_delta_u, du);
    put(out->pixel_delta_v, dv);
    put(out->viewport_top_left, tl);
    return NRRT_OK;
}

}  // extern "C"
