#!/usr/bin/env python3
"""Generates rust/ray-tracer-lib-describe.patch: the change a maintainer applies to the reference tree so that
`Scene::render` (ray-tracer-lib/src/scene.rs:13-18) runs on the B200 through nrrt-sys.

Every concrete Hitable / Material / Texture keeps its fields private (sphere.rs:20-26, plane.rs:33-43, ...), so the
object graph can only be described from inside the crate: one `describe` method per trait, implemented next to each
`bbox` / `scatter` / `get_color`, feeding a `GraphBuilder` (new file graph.rs) that memoises by `Arc` pointer so shared
instances stay shared.  `Rotate` additionally remembers its axis and angle (it only stores the matrices today).

Run in the build container (needs /root/reference to read the files being patched):
    python tools/make_describe_patch.py
The edits are anchored on exact reference lines; the script fails if an anchor no longer matches."""
import difflib
import os
import sys

REF = os.environ.get("NRRT_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "rust", "ray-tracer-lib-describe.patch")
LIB = "packages/ray-tracer-lib"

# (file, anchor line (exact, first occurrence after `after` if given), text inserted AFTER the anchor)
EDITS = [
    (f"{LIB}/Cargo.toml", 'noise = "0.9.0"', 'nrrt-sys = { path = "../nrrt-sys" }\n'),
    (f"{LIB}/src/lib.rs", "pub mod camera;", "pub mod graph;\n"),
    (f"{LIB}/src/prelude.rs", "pub use crate::camera::*;", "pub use crate::graph::*;\n"),
    (f"{LIB}/src/hitable.rs", "    fn hit(&self, ray: &Ray, hit_range: Interval) -> Option<HitRecord>;",
     "    /// Appends this object to the graph description handed to nrrt_host_build; returns its object index.\n"
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32;\n"),
    (f"{LIB}/src/materials/material.rs", "        DVec3::ZERO\n    }",
     "\n    /// {kind, texture, param} of include/nrrt.h; returns the material index.\n"
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32;\n"),
    (f"{LIB}/src/textures/texture.rs", "    fn get_color(&self, uv_coord: DVec2, point: DVec3) -> DVec3;",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32;\n"),
    # ---- objects
    (f"{LIB}/src/objects/sphere.rs", "impl Hitable for Sphere {\n    fn bbox(&self) -> AABB {\n        self.bbox\n    }",
     "\n    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let m = g.material(&self.material);\n"
     "        let s = self.speed.unwrap_or(DVec3::ZERO);  // with_speed(): v[4..6], zero = static\n"
     "        g.object(nrrt_sys::NRRT_OBJ_SPHERE, m, &[],\n"
     "                 &[self.center.x, self.center.y, self.center.z, self.radius, s.x, s.y, s.z])\n"
     "    }\n"),
    (f"{LIB}/src/objects/plane.rs", "impl Hitable for Plane {\n    fn bbox(&self) -> AABB {\n        self.bbox\n    }",
     "\n    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let m = g.material(&self.material);\n"
     "        let kind = match self.shape { Shape::Quad => nrrt_sys::NRRT_OBJ_QUAD, Shape::Triangle => nrrt_sys::NRRT_OBJ_TRIANGLE };\n"
     "        g.object(kind, m, &[], &[self.p.x, self.p.y, self.p.z, self.u.x, self.u.y, self.u.z, self.v.x, self.v.y, self.v.z])\n"
     "    }\n"),
    (f"{LIB}/src/objects/translate.rs", "impl Hitable for Translate {\n    fn bbox(&self) -> AABB {\n        self.bbox\n    }",
     "\n    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let child = g.hitable(&self.object);\n"
     "        g.object(nrrt_sys::NRRT_OBJ_TRANSLATE, 0, &[child], &[self.offset.x, self.offset.y, self.offset.z])\n"
     "    }\n"),
    (f"{LIB}/src/objects/rotate.rs", "    rotation_mat_inv: DMat3,\n    bbox: AABB,",
     "    axis: DVec3,   // kept for Hitable::describe (only the matrices were stored)\n    angle: f64,\n"),
    (f"{LIB}/src/objects/rotate.rs", "            rotation_mat_inv,\n            bbox,", "            axis,\n            angle,\n"),
    (f"{LIB}/src/objects/rotate.rs", "impl Hitable for Rotate {\n    fn bbox(&self) -> AABB {\n        self.bbox\n    }",
     "\n    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let child = g.hitable(&self.object);\n"
     "        let kind = if self.axis == DVec3::X { nrrt_sys::NRRT_OBJ_ROTATE_X }\n"
     "                   else if self.axis == DVec3::Y { nrrt_sys::NRRT_OBJ_ROTATE_Y } else { nrrt_sys::NRRT_OBJ_ROTATE_Z };\n"
     "        g.object(kind, 0, &[child], &[self.angle])\n"
     "    }\n"),
    (f"{LIB}/src/objects/scale.rs", "impl Hitable for Scale {\n    fn bbox(&self) -> AABB {\n        self.bbox\n    }",
     "\n    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let child = g.hitable(&self.object);\n"
     "        let m = &self.scale_matrix;  // DMat4::from_scale(scale): the scale vector is the diagonal\n"
     "        g.object(nrrt_sys::NRRT_OBJ_SCALE, 0, &[child], &[m.x_axis.x, m.y_axis.y, m.z_axis.z])\n"
     "    }\n"),
    (f"{LIB}/src/objects/object.rs", "impl Hitable for BVH {",
     "    /// The tree goes over as its leaf list, in order: nrrt_host_build re-runs BVH::from with the reference's own\n"
     "    /// algorithm (object.rs:41-73) on it, so the device walks the identical tree.\n"
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let leaves: Vec<Arc<dyn Hitable + Send + Sync>> = self.clone().into();\n"
     "        let children: Vec<u32> = leaves.iter().map(|o| g.hitable(o)).collect();\n"
     "        g.object(nrrt_sys::NRRT_OBJ_GROUP, 0, &children, &[])\n"
     "    }\n\n"),
    # ---- materials
    (f"{LIB}/src/materials/lambertian.rs", "impl Material for Lambertian {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let t = g.texture(&self.texture);\n"
     "        g.add_material(nrrt_sys::NRRT_MAT_LAMBERTIAN, t, 0.0)\n"
     "    }\n\n"),
    (f"{LIB}/src/materials/metal.rs", "impl Material for Metal {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let t = g.texture(&self.texture);\n"
     "        g.add_material(nrrt_sys::NRRT_MAT_METAL, t, self.fuzz)\n"
     "    }\n\n"),
    (f"{LIB}/src/materials/dielectric.rs", "impl Material for Dielectric {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        g.add_material(nrrt_sys::NRRT_MAT_DIELECTRIC, 0, self.refraction_index)\n"
     "    }\n\n"),
    (f"{LIB}/src/materials/diffuse_light.rs", "impl Material for DiffuseLight {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let t = g.texture(&self.texture);\n"
     "        g.add_material(nrrt_sys::NRRT_MAT_DIFFUSE_LIGHT, t, self.intensity)\n"
     "    }\n\n"),
    # ---- textures
    (f"{LIB}/src/textures/solid_color.rs", "impl Texture for SolidColor {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        g.add_texture(nrrt_sys::NRRT_TEX_SOLID, 0, 0, 0, 0, self.color, [0.0; 3])\n"
     "    }\n\n"),
    (f"{LIB}/src/textures/checker.rs", "impl Texture for Checker {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        let (a, b) = (g.texture(&self.even_texture), g.texture(&self.odd_texture));  // sub-textures first\n"
     "        g.add_texture(nrrt_sys::NRRT_TEX_CHECKER, a, b, 0, 0, DVec3::ZERO, [self.scale, 0.0, 0.0])\n"
     "    }\n\n"),
    (f"{LIB}/src/textures/image.rs", "impl Texture for Image {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        // into_rgb32f() stored u8/255 (image.rs:24): x255 and round gives the decoded bytes back exactly\n"
     "        let rgb: Vec<u8> = self.image.as_raw().iter().map(|v| (v * 255.0).round() as u8).collect();\n"
     "        let img = g.add_image(self.image.width(), self.image.height(), rgb);\n"
     "        g.add_texture(nrrt_sys::NRRT_TEX_IMAGE, img, 0, 0, 0, DVec3::ZERO, [0.0; 3])\n"
     "    }\n\n"),
    (f"{LIB}/src/textures/noise.rs", "impl Texture for PerlinRidgedNoise {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        g.add_texture(nrrt_sys::NRRT_TEX_NOISE, 0, 0, self.seed, self.octaves as u32, DVec3::ZERO,\n"
     "                      [self.frequency, self.lacunarity, self.persistence])\n"
     "    }\n\n"),
    (f"{LIB}/src/textures/marble.rs", "impl Texture for Marble {",
     "    fn describe(&self, g: &mut crate::graph::GraphBuilder) -> u32 {\n"
     "        g.add_texture(nrrt_sys::NRRT_TEX_MARBLE, 0, 0, self.seed, 7, DVec3::ZERO, [self.frequency, 0.0, 0.0])\n"
     "    }\n\n"),
    # ---- camera: hand the built Camera over (its fields are private, camera.rs:206-227)
    (f"{LIB}/src/camera.rs", "    pub fn get_image_size(&self) -> ImageSize {\n        self.image_size\n    }",
     "\n    /// The ten fields of the built camera in nrrt_camera layout (include/nrrt.h).\n"
     "    pub fn to_nrrt(&self) -> nrrt_sys::nrrt_camera {\n"
     "        let a = |v: DVec3| [v.x, v.y, v.z];\n"
     "        nrrt_sys::nrrt_camera {\n"
     "            width: self.image_size.width as u32, height: self.image_size.height as u32,\n"
     "            samples_per_pixel: self.samples_per_pixel as u32, ray_max_bounces: self.ray_max_bounces as u32,\n"
     "            background: a(self.background_color), look_from: a(self.look_from),\n"
     "            defocus_disk_u: a(self.defocus_disk_u), defocus_disk_v: a(self.defocus_disk_v),\n"
     "            pixel_delta_u: a(self.viewport_pixel_delta_u), pixel_delta_v: a(self.viewport_pixel_delta_v),\n"
     "            viewport_top_left: a(self.viewport_top_left),\n"
     "        }\n"
     "    }\n"),
]
# (file, exact old text, new text): replacements
REPLACE = [
    (f"{LIB}/src/scene.rs", "        self.camera.render(&self.objects, progress)",
     "        // B200 path (no CPU fallback: NRRT_ERR_NO_DEVICE aborts with the library's message)\n"
     "        crate::graph::render_on_gpu(&self.camera, &self.objects, progress)"),
    (f"{LIB}/src/objects/rotate.rs", "        Self {\n            object,\n            rotation_mat,", None),  # anchor check only
]

GRAPH_RS = '''//! Object graph -> nrrt_graph_desc (include/nrrt.h) and the call into libnrrt_b200.so.
//! New file of the integration patch; see INTEGRATION.md in the B200 repository.
use std::collections::HashMap;
use std::ffi::{c_void, CStr};
use std::sync::Arc;

use glam::DVec3;
use image::Rgb32FImage;
use nrrt_sys::*;

use crate::camera::Camera;
use crate::hitable::Hitable;
use crate::materials::Material;
use crate::objects::BVH;
use crate::textures::Texture;

/// Collects objects / materials / textures in the index-based form of nrrt_graph_desc.  Shared `Arc`s are described
/// once (memo by pointer), so instanced groups stay shared on the device.
#[derive(Default)]
pub struct GraphBuilder {
    objects: Vec<nrrt_object>,
    child_ids: Vec<u32>,
    materials: Vec<nrrt_material>,
    textures: Vec<nrrt_texture>,
    images: Vec<(u32, u32, Vec<u8>)>,
    seen_objects: HashMap<*const (), u32>,
    seen_materials: HashMap<*const (), u32>,
    seen_textures: HashMap<*const (), u32>,
}

impl GraphBuilder {
    pub fn hitable(&mut self, o: &Arc<dyn Hitable + Send + Sync>) -> u32 {
        let key = Arc::as_ptr(o) as *const ();
        if let Some(&i) = self.seen_objects.get(&key) { return i; }
        let i = o.describe(self);
        self.seen_objects.insert(key, i);
        i
    }
    pub fn material(&mut self, m: &Arc<dyn Material + Send + Sync>) -> u32 {
        let key = Arc::as_ptr(m) as *const ();
        if let Some(&i) = self.seen_materials.get(&key) { return i; }
        let i = m.describe(self);
        self.seen_materials.insert(key, i);
        i
    }
    pub fn texture(&mut self, t: &Arc<dyn Texture + Send + Sync>) -> u32 {
        let key = Arc::as_ptr(t) as *const ();
        if let Some(&i) = self.seen_textures.get(&key) { return i; }
        let i = t.describe(self);
        self.seen_textures.insert(key, i);
        i
    }
    pub fn object(&mut self, kind: u32, material: u32, children: &[u32], v: &[f64]) -> u32 {
        let mut o = nrrt_object { kind, material, first_child: self.child_ids.len() as u32,
                                  n_children: children.len() as u32, v: [0.0; 9] };
        o.v[..v.len()].copy_from_slice(v);
        self.child_ids.extend_from_slice(children);
        self.objects.push(o);
        (self.objects.len() - 1) as u32
    }
    pub fn add_material(&mut self, kind: u32, texture: u32, param: f64) -> u32 {
        self.materials.push(nrrt_material { kind, texture, param });
        (self.materials.len() - 1) as u32
    }
    #[allow(clippy::too_many_arguments)]
    pub fn add_texture(&mut self, kind: u32, a: u32, b: u32, seed: u32, octaves: u32, color: DVec3, f: [f64; 3]) -> u32 {
        self.textures.push(nrrt_texture { kind, a, b, seed, octaves, _pad: 0, color: [color.x, color.y, color.z],
                                          f0: f[0], f1: f[1], f2: f[2] });
        (self.textures.len() - 1) as u32
    }
    pub fn add_image(&mut self, width: u32, height: u32, rgb: Vec<u8>) -> u32 {
        self.images.push((width, height, rgb));
        (self.images.len() - 1) as u32
    }
}

fn message(p: *const std::os::raw::c_char) -> String {
    if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() }
}

/// Body of `Scene::render` (scene.rs:13-18): Camera::render (camera.rs:302-343) on the GPU.
pub fn render_on_gpu<P>(camera: &Camera, objects: &BVH, progress: Option<P>) -> Rgb32FImage where P: Fn() + Sync {
    check_abi();
    let mut g = GraphBuilder::default();
    let root = objects.describe(&mut g);
    let images: Vec<nrrt_image> =
        g.images.iter().map(|(w, h, rgb)| nrrt_image { width: *w, height: *h, rgb: rgb.as_ptr() }).collect();
    let desc = nrrt_graph_desc {
        n_objects: g.objects.len() as u32, objects: g.objects.as_ptr(),
        n_child_ids: g.child_ids.len() as u32, child_ids: g.child_ids.as_ptr(),
        n_materials: g.materials.len() as u32, materials: g.materials.as_ptr(),
        n_textures: g.textures.len() as u32, textures: g.textures.as_ptr(),
        n_images: images.len() as u32, images: images.as_ptr(),
        root,
    };
    unsafe {
        let host = nrrt_host_build(&desc);  // the reference's BVH build + flatten, on the host
        assert!(!host.is_null(), "nrrt_host_build: {}", message(nrrt_host_last_error()));
        let mut ctx = std::ptr::null_mut();
        let rc = nrrt_create(0, &mut ctx);
        assert!(rc == NRRT_OK, "nrrt_create: {}", message(nrrt_last_error(std::ptr::null())));
        let rc = nrrt_scene_upload(ctx, nrrt_host_scene_desc(host));
        assert!(rc == NRRT_OK, "nrrt_scene_upload: {}", message(nrrt_last_error(ctx)));
        let cam = camera.to_nrrt();
        let mut pixels = vec![0f32; cam.width as usize * cam.height as usize * 3];
        // the reference ticks once per finished pixel (camera.rs:333-335); the library reports pixel counts
        extern "C" fn tick<P: Fn() + Sync>(done: u64, _total: u64, user: *mut c_void) {
            let (f, last) = unsafe { &mut *(user as *mut (&P, u64)) };
            for _ in *last..done { f() }
            *last = done;
        }
        let mut user = progress.as_ref().map(|p| (p, 0u64));
        let opts = nrrt_render_opts { seed: 0, mode: NRRT_MODE_AUTO, rank: 0, world: 1, rows_per_block: 0, max_slots: 0,
                                      flags: 0 };  // seed_from_u64(0), camera.rs:318
        let cb: nrrt_progress_fn = if user.is_some() { Some(tick::<P>) } else { None };
        let up = user.as_mut().map_or(std::ptr::null_mut(), |u| u as *mut _ as *mut c_void);
        let rc = nrrt_render(ctx, &cam, &opts, pixels.as_mut_ptr(), cb, up, std::ptr::null_mut());
        assert!(rc == NRRT_OK, "nrrt_render: {}", message(nrrt_last_error(ctx)));
        nrrt_destroy(ctx);
        nrrt_host_free(host);
        Rgb32FImage::from_vec(cam.width, cam.height, pixels).unwrap()  // camera.rs:342
    }
}
'''


def main():
    files = {}
    for path, anchor, ins in EDITS:
        full = os.path.join(REF, path)
        if path not in files:
            files[path] = [open(full).read()] * 2
        cur = files[path][1]
        k = cur.find(anchor)
        if k < 0:
            sys.exit(f"anchor not found in {path}: {anchor!r}")
        end = k + len(anchor)
        if cur[end:end + 1] == "\n":
            end += 1
        files[path][1] = cur[:end] + ins + cur[end:]
    for path, old, new in REPLACE:
        full = os.path.join(REF, path)
        if path not in files:
            files[path] = [open(full).read()] * 2
        if old not in files[path][1]:
            sys.exit(f"text to replace not found in {path}: {old!r}")
        if new is not None:
            files[path][1] = files[path][1].replace(old, new, 1)
    out = []
    for path in sorted(files):
        a, b = files[path]
        out += difflib.unified_diff(a.splitlines(True), b.splitlines(True), "a/" + path, "b/" + path, n=2)
    gpath = f"{LIB}/src/graph.rs"
    out += difflib.unified_diff([], GRAPH_RS.splitlines(True), "/dev/null", "b/" + gpath, n=0)
    header = ("# Integration patch for NealRame/nr-ray-tracer (apply at the reference root with `patch -p1`), generated by\n"
              "# tools/make_describe_patch.py of the B200 repository.  It adds the flatten hook (describe) and routes\n"
              "# Scene::render through nrrt-sys.  Also copy rust/nrrt-sys to packages/nrrt-sys and add it to the workspace\n"
              "# members in Cargo.toml.  Not compiled in the B200 repository's build image (no cargo there).\n")
    open(OUT, "w").write(header + "".join(out))
    print(OUT, len(out), "lines")


if __name__ == "__main__":
    main()
