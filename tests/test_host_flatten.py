"""Host BVH build + flattening (csrc/host_scene.cpp) checked on the CPU:
structure invariants, camera maths vs the oracle, and — through a pure-Python interpreter of the flat layout
with reference traversal semantics (tests/flat_interp.py) — that the flattened tree gives the oracle's answers."""
import numpy as np
import pytest

from nr_ray_tracer_b200 import _abi as A
from nr_ray_tracer_b200 import api
from oracle import oracle as O
from tests import flat_interp, kat
from tests.scenes_util import ALL_SCENES, load


@pytest.mark.parametrize("name", ALL_SCENES)
def test_flat_structure_invariants(name):
    g = load(name)
    hs = api.HostScene(g)
    d = hs.desc
    nodes = hs.nodes()
    boxes = hs.child_boxes()
    assert d.abi_version == A.ABI_VERSION
    assert d.max_stack <= 32
    # f32 culling boxes are the nearest-rounded f64 boxes
    if d.n_nodes:
        valid = nodes["child"] != A.REF_NONE
        lo64, hi64 = boxes[:, :, 0, :], boxes[:, :, 1, :]
        assert np.array_equal(nodes["lo"][valid], lo64[valid].astype(np.float32))
        assert np.array_equal(nodes["hi"][valid], hi64[valid].astype(np.float32))
        assert (lo64[valid] <= hi64[valid]).all()
        # aabb.rs:16-30: every box axis is at least 1e-4 wide (up to rounding)
        assert ((hi64[valid] - lo64[valid]) >= 1e-4 * (1 - 1e-9)).all()
    # every node index is referenced exactly once (root or as a child) within its tree, leaves once each
    refs = nodes["child"].reshape(-1) if d.n_nodes else np.zeros(0, dtype=np.uint32)
    types, idx = refs >> A.REF_TYPE_SHIFT, refs & A.REF_INDEX_MASK
    for ty, n in ((A.REF_SPHERE, d.n_spheres), (A.REF_PLANE, d.n_planes)):
        seen = np.sort(idx[types == ty])
        inner_prims = [hs.desc.instances[i].inner for i in range(d.n_instances)
                       if (hs.desc.instances[i].inner >> A.REF_TYPE_SHIFT) == ty]
        seen = np.sort(np.concatenate([seen, np.array([r & A.REF_INDEX_MASK for r in inner_prims], dtype=np.uint32)]))
        if (d.root >> A.REF_TYPE_SHIFT) == ty:
            seen = np.sort(np.append(seen, d.root & A.REF_INDEX_MASK))
        assert seen.tolist() == list(range(n)), "each primitive record is one leaf"
    # one object per leaf, median split => a binary tree over L leaves has L-1 inner nodes in each space
    n_leaves = int((types != A.REF_NODE).sum())
    if d.n_instances == 0 and d.n_nodes:
        assert d.n_nodes == n_leaves - 1
    # DFS emission: children of node i come after i
    node_children = idx[types == A.REF_NODE]
    parents = np.repeat(np.arange(d.n_nodes), 2)[types == A.REF_NODE]
    assert (node_children > parents).all()


def test_spheres_tree_matches_survey_measurement():
    # spheres.toml -> 487 inner nodes, 488 leaves, leaf depth 8-9
    hs = api.HostScene(load("spheres.toml"))
    assert hs.desc.n_nodes == 487 and hs.desc.n_spheres == 488
    nodes = hs.nodes()
    depth = {0: 0}
    leaf_depths = []
    for i in range(len(nodes)):
        for c in nodes["child"][i]:
            if (c >> A.REF_TYPE_SHIFT) == A.REF_NODE:
                depth[int(c & A.REF_INDEX_MASK)] = depth[i] + 1
            else:
                leaf_depths.append(depth[i] + 1)
    assert min(leaf_depths) == 8 and max(leaf_depths) == 9
    # leaf order ids are the DFS positions 0..487
    order = np.ctypeslib.as_array(hs.desc.sphere_order, shape=(488,))
    assert order.tolist() == list(range(488))


def test_leaf_order_follows_reference_sort_not_input_order():
    # three spheres given in reverse x order: BVH::from sorts by bbox.min.x (object.rs:59-64), split 1 | 2
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    t = g.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g.add_material(A.MAT_LAMBERTIAN, t)
    ids = [g.add_object(A.OBJ_SPHERE, m, v=(x, 0, 0, 1)) for x in (6.0, 3.0, 0.0)]
    g.root = g.add_object(A.OBJ_GROUP, children=ids)
    hs = api.HostScene(g)
    objs = np.ctypeslib.as_array(hs.desc.sphere_object, shape=(3,)).tolist()
    assert objs == [ids[2], ids[1], ids[0]]
    nodes = hs.nodes()
    assert (nodes["child"][0][0] >> A.REF_TYPE_SHIFT) == A.REF_SPHERE      # left = single leaf (mid = 3/2 = 1)
    assert (nodes["child"][0][1] >> A.REF_TYPE_SHIFT) == A.REF_NODE


def test_longest_axis_tie_picks_highest_axis():
    # cube of spheres: all extents equal -> max_by keeps the last maximum = z (aabb.rs:101-108)
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    t = g.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g.add_material(A.MAT_LAMBERTIAN, t)
    pts = [(x, y, z) for x in (0.0, 4.0) for y in (0.0, 4.0) for z in (0.0, 4.0)]
    ids = [g.add_object(A.OBJ_SPHERE, m, v=p + (1.0,)) for p in pts]
    g.root = g.add_object(A.OBJ_GROUP, children=ids)
    hs = api.HostScene(g)
    boxes = hs.child_boxes()
    # root split on z: left child box spans z in [-1, 1], right child z in [3, 5]
    assert boxes[0, 0, 0, 2] == -1.0 and boxes[0, 0, 1, 2] == 1.0
    assert boxes[0, 1, 0, 2] == 3.0 and boxes[0, 1, 1, 2] == 5.0


@pytest.mark.parametrize("name", ["spheres.toml", "cornell-box-scene.json", "scale.json", "cube-scene.json",
                                  "simple-lights.toml", "utah-teapot-scene.json"])
def test_flat_layout_traversed_with_reference_semantics_matches_oracle(name):
    g = load(name)
    hs = api.HostScene(g)
    fs = flat_interp.FlatScene(hs)
    sp = kat.special_rays(g)
    rng = np.random.default_rng(0)
    rays = np.concatenate([kat.random_rays(g, 250, seed=3), kat.aimed_rays(g, 350, seed=4),
                           sp[rng.permutation(len(sp))[:300]]])
    ref, _ = O.OracleScene(g).trace_rays(rays)
    for i, r in enumerate(rays):
        h = flat_interp.trace(fs, r[:3], r[3:])
        if ref["object"][i] == 0xFFFFFFFF:
            assert h is None, (i, r, h)
        else:
            assert h is not None and h[1] == ref["object"][i] and h[0] == ref["t"][i], (i, r, h, ref[i])


@pytest.mark.parametrize("bvh", ["reference", "sah"])
@pytest.mark.parametrize("name", ["spheres.toml", "cornell-box-scene.json", "scale.json", "cube-scene.json",
                                  "simple-lights.toml", "utah-teapot-scene.json", "cornell-teapot-scene.json"])
def test_four_slot_nodes_fold_the_binary_tree_and_give_the_oracle_hits(name, bvh):
    """nrrt_wnode (what the kernels walk): every other level of the binary tree folded away.  Structure: each leaf in
    exactly one slot, slots in depth-first leaf order, f32 slot boxes = the rounded f64 boxes, gates enclose their
    slots, stack bound respected.  Semantics: walked with the reference's exact tests (gate box for gated slots, own
    box for inner slots, none for leaves) it returns the oracle's hits, as the binary layout does."""
    g = load(name)
    hs = api.HostScene(g, bvh=bvh)
    d = hs.desc
    wn, wb = hs.wnodes(), hs.wide_boxes()
    assert d.n_wnodes == len(wn) and d.max_stack <= 32
    if d.n_nodes:
        assert 0 < d.n_wnodes <= d.n_nodes
    refs = wn["child"].reshape(-1)
    used = refs != A.REF_NONE
    types, idx = refs >> A.REF_TYPE_SHIFT, refs & A.REF_INDEX_MASK
    # every leaf of the binary layout sits in exactly one slot (or is a space root / instance inner)
    brefs = hs.nodes()["child"].reshape(-1)
    bleaves = sorted(brefs[(brefs != A.REF_NONE) & ((brefs >> A.REF_TYPE_SHIFT) != A.REF_NODE)].tolist())
    wleaves = sorted(refs[used & (types != A.REF_NODE)].tolist())
    assert wleaves == bleaves
    # every wide node is referenced exactly once (slot, scene root or a nested space's root)
    roots = [int(d.wide_root)] + [int(r) for r in hs.instance_wide_inner()]
    node_refs = sorted(idx[used & (types == A.REF_NODE)].tolist() +
                       list({r & A.REF_INDEX_MASK for r in roots if r != A.REF_NONE and (r >> A.REF_TYPE_SHIFT) == A.REF_NODE}))
    assert node_refs == list(range(d.n_wnodes))
    lo32, hi32 = wn["lo"].transpose(0, 2, 1).reshape(-1, 3), wn["hi"].transpose(0, 2, 1).reshape(-1, 3)
    own, gate = wb[:, :, 0].reshape(-1, 2, 3), wb[:, :, 1].reshape(-1, 2, 3)
    assert np.array_equal(lo32[used], own[used][:, 0].astype(np.float32))
    assert np.array_equal(hi32[used], own[used][:, 1].astype(np.float32))
    assert np.isinf(lo32[~used]).all() and np.isinf(hi32[~used]).all()       # unused slots: empty boxes
    gated = used & ((wn["meta"].reshape(-1) & A.WNODE_GATED) != 0)
    assert gated.any() or d.n_nodes < 2
    assert (gate[gated][:, 0] <= own[gated][:, 0]).all() and (gate[gated][:, 1] >= own[gated][:, 1]).all()
    # slot order is depth-first leaf order: within one wide node of a primitive-only subtree, orders ascend
    if bvh == "reference" and d.n_instances == 0 and d.n_spheres:
        order = np.ctypeslib.as_array(d.sphere_order, shape=(d.n_spheres,))
        for row in wn["child"]:
            o = [int(order[r & A.REF_INDEX_MASK]) for r in row if r != A.REF_NONE and (r >> A.REF_TYPE_SHIFT) == A.REF_SPHERE]
            assert o == sorted(o)
    fs = flat_interp.FlatScene(hs)
    sp = kat.special_rays(g)
    rng = np.random.default_rng(1)
    rays = np.concatenate([kat.random_rays(g, 200, seed=13), kat.aimed_rays(g, 300, seed=14),
                           sp[rng.permutation(len(sp))[:250]]])
    ref, _ = O.OracleScene(g).trace_rays(rays)
    for i, r in enumerate(rays):
        h = flat_interp.trace(fs, r[:3], r[3:], tie_by_order=(bvh == "sah"), wide=True)
        if ref["object"][i] == 0xFFFFFFFF:
            assert h is None, (i, r, h)
        else:
            assert h is not None and h[1] == ref["object"][i] and h[0] == ref["t"][i], (i, r, h, ref[i])


@pytest.mark.parametrize("name", ALL_SCENES)
def test_camera_build_matches_oracle_bit_for_bit(name):
    g = load(name, width=1920, height=1080)
    cfg = g.camera.to_builder_config()
    assert bytes(api.camera_build(cfg)) == bytes(O.camera_build(cfg))


def test_camera_build_hand_values():
    # look down -z from the origin, fov 90deg, focus 1, square image 2x2 (camera.rs:94-159)
    cfg = A.CameraConfig(width=2, height=2, samples_per_pixel=0, ray_max_bounces=3)
    for i, v in enumerate((0.0, 0.0, 0.0)):
        cfg.look_from[i] = v
    cfg.look_at[2] = -1.0
    cfg.view_up[1] = 1.0
    cfg.field_of_view, cfg.focus_dist, cfg.defocus_angle = np.pi / 2, 1.0, -3.0
    cam = api.camera_build(cfg)
    assert cam.samples_per_pixel == 1                                  # clamped to >= 1 (:104)
    assert np.allclose(list(cam.pixel_delta_u), [1, 0, 0]) and np.allclose(list(cam.pixel_delta_v), [0, -1, 0])
    assert np.allclose(list(cam.viewport_top_left), [-0.5, 0.5, -1.0])
    assert list(cam.defocus_disk_u) == [0, 0, 0] and list(cam.defocus_disk_v) == [0, 0, 0]  # angle clamped to 0 (:106)
    cfg.width = 0
    with pytest.raises(api.NrrtError):
        api.camera_build(cfg)


def test_shared_group_behind_two_instances_is_flattened_once():
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    t = g.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g.add_material(A.MAT_LAMBERTIAN, t)
    prims = [g.add_object(A.OBJ_SPHERE, m, v=(float(i), 0, 0, 0.4)) for i in range(8)]
    grp = g.add_object(A.OBJ_GROUP, children=prims)
    insts = [g.add_object(A.OBJ_TRANSLATE, children=[grp], v=(0, 3.0 * k, 0)) for k in range(5)]
    g.root = g.add_object(A.OBJ_GROUP, children=insts)
    hs = api.HostScene(g)
    d = hs.desc
    assert d.n_instances == 5 and d.n_spheres == 8       # inner space shared, not copied per instance
    assert d.n_nodes == 4 + 7
    inner = {d.instances[i].inner for i in range(5)}
    assert len(inner) == 1
    # and it traces like the oracle
    fs = flat_interp.FlatScene(hs)
    rays = kat.aimed_rays(g, 200, seed=1)
    ref, _ = O.OracleScene(g).trace_rays(rays)
    for i, r in enumerate(rays):
        h = flat_interp.trace(fs, r[:3], r[3:])
        assert (h is None) == (ref["object"][i] == 0xFFFFFFFF)
        if h is not None:
            assert h[1] == ref["object"][i] and h[0] == ref["t"][i]


def test_moving_sphere_bbox_is_union_of_both_shutter_ends_and_speed_is_exported():
    """sphere.rs:72-78: bbox = union(box at center, box at center + speed); static scenes export no speed array."""
    from nr_ray_tracer_b200 import api
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    t = g.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g.add_material(A.MAT_LAMBERTIAN, t)
    a = g.add_object(A.OBJ_SPHERE, m, v=(0, 0, 0, 1, 3, -2, 0.5))
    b = g.add_object(A.OBJ_SPHERE, m, v=(10, 0, 0, 1))
    g.root = g.add_object(A.OBJ_GROUP, children=[a, b])
    hs = api.HostScene(g)
    d = hs.desc
    assert d.abi_version == A.ABI_VERSION and d.n_spheres == 2 and bool(d.sphere_speed)
    speed = np.ctypeslib.as_array(d.sphere_speed, shape=(2, 3))
    obj = np.ctypeslib.as_array(d.sphere_object, shape=(2,))
    assert speed[list(obj).index(a)].tolist() == [3.0, -2.0, 0.5] and speed[list(obj).index(b)].tolist() == [0, 0, 0]
    assert list(d.root_box.lo) == [-1.0, -3.0, -1.0] and list(d.root_box.hi) == [11.0, 1.0, 1.5]
    # a scene whose spheres all have zero speed exports no speed array -> static kernels
    g2 = SceneGraph()
    t2 = g2.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m2 = g2.add_material(A.MAT_LAMBERTIAN, t2)
    g2.root = g2.add_object(A.OBJ_GROUP, children=[g2.add_object(A.OBJ_SPHERE, m2, v=(0, 0, 0, 1, 0, 0, 0)),
                                                   g2.add_object(A.OBJ_SPHERE, m2, v=(3, 0, 0, 1))])
    assert not bool(api.HostScene(g2).desc.sphere_speed)


@pytest.mark.parametrize("name", ["spheres.toml", "cornell-box-scene.json", "cube-scene.json", "utah-teapot-scene.json"])
def test_sah_build_keeps_leaves_and_order_and_gives_oracle_hits(name):
    """NRRT_BUILD_SAH (§8(f) N3): same leaves, records and reference DFS order numbers; only inner nodes differ.
    Traversed with the reference's box / primitive tests and order-based tie-break it returns the oracle's hits."""
    g = load(name)
    ref_hs, hs = api.HostScene(g), api.HostScene(g, bvh="sah")
    a, b = ref_hs.desc, hs.desc
    assert (a.n_nodes, a.n_spheres, a.n_planes, a.n_instances) == (b.n_nodes, b.n_spheres, b.n_planes, b.n_instances)
    for field, n in (("sphere_order", a.n_spheres), ("plane_order", a.n_planes), ("instance_order", a.n_instances),
                     ("sphere_object", a.n_spheres), ("plane_object", a.n_planes)):
        if n:
            assert np.array_equal(np.ctypeslib.as_array(getattr(a, field), shape=(n,)),
                                  np.ctypeslib.as_array(getattr(b, field), shape=(n,))), field
    assert list(a.root_box.lo) == list(b.root_box.lo) and list(a.root_box.hi) == list(b.root_box.hi)
    assert b.max_stack <= 32
    # every leaf is referenced exactly once by the SAH nodes, and child boxes enclose their subtrees
    nodes = hs.nodes()
    refs = nodes["child"].reshape(-1)
    leaves = refs[(refs >> A.REF_TYPE_SHIFT) != A.REF_NODE]
    assert len(set(leaves.tolist())) == len(leaves)
    if name != "spheres.toml":
        assert not np.array_equal(nodes["child"], ref_hs.nodes()["child"])   # it really is another tree
    fs = flat_interp.FlatScene(hs)
    rays = np.concatenate([kat.random_rays(g, 300, seed=5), kat.aimed_rays(g, 500, seed=6)])
    ref, _ = O.OracleScene(g).trace_rays(rays)
    for i, r in enumerate(rays):
        h = flat_interp.trace(fs, r[:3], r[3:], tie_by_order=True)
        if ref["object"][i] == 0xFFFFFFFF:
            assert h is None, (i, r, h)
        else:
            assert h is not None and h[1] == ref["object"][i] and h[0] == ref["t"][i], (i, r, h, ref[i])


def test_sah_surface_area_cost_is_lower_than_median_split_on_the_mesh():
    g = load("utah-teapot-scene.json")

    def sah_cost(hs):
        cb = hs.child_boxes()                      # (n, 2, 2, 3)
        ext = cb[:, :, 1, :] - cb[:, :, 0, :]
        area = ext[..., 0] * ext[..., 1] + ext[..., 1] * ext[..., 2] + ext[..., 2] * ext[..., 0]
        return float(area.sum())                   # sum of child-box areas ~ expected node + leaf visits
    assert sah_cost(api.HostScene(g, bvh="sah")) < 0.8 * sah_cost(api.HostScene(g))


def test_non_finite_geometry_does_not_break_either_build():
    """NaN / inf coordinates (malformed input): both tree builds terminate with a valid flat scene (the reference
    sorts with total_cmp; the SAH binning must not index out of range)."""
    from nr_ray_tracer_b200 import api
    from nr_ray_tracer_b200.scene_config import SceneGraph
    g = SceneGraph()
    t = g.add_texture(kind=A.TEX_SOLID, color=(1, 1, 1))
    m = g.add_material(A.MAT_LAMBERTIAN, t)
    rng = np.random.default_rng(0)
    kids = []
    for i in range(300):
        c = rng.uniform(-5, 5, 3)
        if i % 37 == 0:
            c[0] = float("nan")
        if i % 53 == 0:
            c[1] = float("inf")
        kids.append(g.add_object(A.OBJ_SPHERE, m, v=(*c, 0.3)))
    g.root = g.add_object(A.OBJ_GROUP, children=kids)
    for bvh in ("reference", "sah"):
        hs = api.HostScene(g, bvh=bvh)
        assert hs.desc.n_nodes == 299 and hs.desc.n_spheres == 300 and hs.desc.max_stack <= 32
        refs = hs.nodes()["child"].reshape(-1)
        leaves = refs[(refs >> A.REF_TYPE_SHIFT) != A.REF_NODE]
        assert len(set(leaves.tolist())) == 300
