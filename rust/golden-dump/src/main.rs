//! Known-answer dump from the REAL nr-ray-tracer crates, in the layout tests/test_oracle.py::test_against_rust_dump
//! reads.  Not compiled in the build image of the B200 repository (no cargo there); see ../README.md.
//!
//!   cargo run --release -- <reference scenes dir> <tests/golden/rust_dump>
//!
//! Run it from the reference checkout's root (scene files name textures / sub-scenes relative to the CWD,
//! scene_config.rs:88, :337).  For every `<scene>.rays.bin` in the dump directory (n x 6 f64, written by
//! export_rays.py) it writes `<scene>.hits.bin`: n x 11 f64 = hit (0/1), t, point xyz, normal xyz, u, v, front_face —
//! the HitRecord of `Scene.objects.hit(ray, [0.001, +inf])` (objects/object.rs:89-121, camera.rs:280).
//! Then the third-party arithmetic the oracle restates from memory:
//!   perm.bin            9 x 256 u8     noise 0.9.0 PermutationTable::new(seed), seeds 0..=8
//!   textures.bin        f64            |Fbm<Perlin>| (textures/noise.rs) and Marble (marble.rs) at texture_points.bin
//!   dmat4_inverse.bin   4 x 16 f64     DMat4::from_scale(s).inverse(), column major (objects/scale.rs:48-49)
//!   glam_misc.bin       f64            DMat3::from_axis_angle, DVec3::reflect / refract / normalize samples

// the reference's scene loader, compiled in by path (it lives in the binary crate)
#[path = "../../reference/packages/ray-tracer/src/cli.rs"]
#[allow(dead_code)]
mod cli;
#[path = "../../reference/packages/ray-tracer/src/constants.rs"]
#[allow(dead_code)]
mod constants;
#[path = "../../reference/packages/ray-tracer/src/scene_config.rs"]
#[allow(dead_code)]
mod scene_config;

use std::fs;
use std::io::Write;
use std::path::{Path, PathBuf};

use anyhow::{anyhow, Result};
use glam::{DMat3, DMat4, DVec2, DVec3};
use noise::permutationtable::{NoiseHasher, PermutationTable};
use nr_ray_tracer_lib::prelude::*;

use scene_config::SceneConfig;

fn read_f64s(path: &Path) -> Result<Vec<f64>> {
    let bytes = fs::read(path)?;
    if bytes.len() % 8 != 0 {
        return Err(anyhow!("{}: not a whole number of f64", path.display()));
    }
    Ok(bytes.chunks_exact(8).map(|c| f64::from_le_bytes(c.try_into().unwrap())).collect())
}

fn write_f64s(path: &Path, v: &[f64]) -> Result<()> {
    let mut f = fs::File::create(path)?;
    for x in v {
        f.write_all(&x.to_le_bytes())?;
    }
    Ok(())
}

/// BVH::hit on the golden rays of one scene file.
fn dump_scene(scene_file: &Path, rays_file: &Path, out_file: &Path) -> Result<()> {
    let scene = SceneConfig::try_load_scene(scene_file)?.try_build()?;
    let rays = read_f64s(rays_file)?;
    let mut out = Vec::with_capacity(rays.len() / 6 * 11);
    for r in rays.chunks_exact(6) {
        let ray = Ray::new(DVec3::new(r[0], r[1], r[2]), DVec3::new(r[3], r[4], r[5]));
        match scene.objects.hit(&ray, Interval::new(0.001, f64::INFINITY)) {
            Some(h) => out.extend_from_slice(&[
                1.0, h.t, h.point.x, h.point.y, h.point.z, h.normal.x, h.normal.y, h.normal.z,
                h.texture_coordinates.x, h.texture_coordinates.y, if h.front_face { 1.0 } else { 0.0 },
            ]),
            None => out.extend_from_slice(&[0.0; 11]),
        }
    }
    write_f64s(out_file, &out)
}

fn main() -> Result<()> {
    let args: Vec<String> = std::env::args().collect();
    if args.len() != 3 {
        return Err(anyhow!("usage: nrrt-golden-dump <reference scenes dir> <dump dir>"));
    }
    let (scenes, dump) = (PathBuf::from(&args[1]), PathBuf::from(&args[2]));

    // ---- 1. BVH::hit per scene
    for entry in fs::read_dir(&dump)? {
        let p = entry?.path();
        let name = p.file_name().unwrap().to_string_lossy().to_string();
        if let Some(scene_name) = name.strip_suffix(".rays.bin") {
            let scene_file = scenes.join(scene_name);
            if !scene_file.exists() {
                eprintln!("skip {scene_name}: not in the reference (e.g. the teapot stand-in)");
                continue;
            }
            match dump_scene(&scene_file, &p, &dump.join(format!("{scene_name}.hits.bin"))) {
                Ok(()) => println!("{scene_name}: ok"),
                // legacy-schema files (spheres/earth/noise.toml) are rejected by the current loader, SURVEY.md note B
                Err(e) => eprintln!("skip {scene_name}: {e}"),
            }
        }
    }

    // ---- 2. noise 0.9.0 permutation tables: hash(&[i]) == values[i & 255] for a one-element key
    let mut perm = Vec::new();
    for seed in 0u32..=8 {
        let table = PermutationTable::new(seed);
        for i in 0..256isize {
            perm.push(table.hash(&[i]) as u8);
        }
    }
    fs::write(dump.join("perm.bin"), &perm)?;

    // ---- 3. Texture::get_color at fixed points (textures/noise.rs:135-145, marble.rs:86-97)
    let pts = read_f64s(&dump.join("texture_points.bin"))?;
    let mut tex = Vec::new();
    // (seed, octaves, lacunarity, persistence, frequency): noise.toml's sphere + defaults + the _textures golden scene
    let noise_cfgs: [(u32, Option<usize>, Option<f64>, Option<f64>, Option<f64>); 4] = [
        (0, None, None, None, None),
        (0, Some(8), None, None, Some(0.2)),
        (3, Some(5), Some(2.1), Some(0.45), Some(1.7)),
        (7, Some(1), None, Some(0.9), Some(3.0)),
    ];
    for (seed, oct, lac, per, freq) in noise_cfgs {
        let mut b = PerlinRidgedNoiseBuilder::default();
        b.with_seed(Some(seed)).with_octaves(oct).with_lacunarity(lac).with_persistence(per).with_frequency(freq);
        let t = b.build();
        for p in pts.chunks_exact(3) {
            tex.push(t.get_color(DVec2::ZERO, DVec3::new(p[0], p[1], p[2])).x);
        }
    }
    for (seed, freq) in [(0u32, None), (1, Some(0.8)), (0, Some(0.2))] {
        let mut b = MarbleBuilder::default();
        b.with_seed(Some(seed)).with_frequency(freq);
        let t = b.build();
        for p in pts.chunks_exact(3) {
            tex.push(t.get_color(DVec2::ZERO, DVec3::new(p[0], p[1], p[2])).x);
        }
    }
    write_f64s(&dump.join("textures.bin"), &tex)?;

    // ---- 4. DMat4::from_scale(s).inverse() (scale.rs:48-49): the two Cornell scales + awkward ones
    let mut inv = Vec::new();
    for s in [DVec3::splat(0.25), DVec3::new(0.25, 0.75, 0.25), DVec3::new(2.0, 3.0, 5.0), DVec3::new(1e-3, 1.0, 1e3)] {
        inv.extend_from_slice(&DMat4::from_scale(s).inverse().to_cols_array());
    }
    write_f64s(&dump.join("dmat4_inverse.bin"), &inv)?;

    // ---- 5. glam closed forms used on the hot path (rotate.rs:52-53, metal.rs:80, dielectric.rs:58-60)
    let mut misc = Vec::new();
    for (axis, angle) in [(DVec3::Y, std::f64::consts::FRAC_PI_4), (DVec3::Y, -std::f64::consts::FRAC_PI_3),
                          (DVec3::Z, std::f64::consts::FRAC_PI_2), (DVec3::X, 0.3)] {
        misc.extend_from_slice(&DMat3::from_axis_angle(axis, -angle).to_cols_array());
        misc.extend_from_slice(&DMat3::from_axis_angle(axis, angle).to_cols_array());
    }
    let i = DVec3::new(0.3, -0.8, 0.52).normalize();
    let n = DVec3::new(0.1, 0.97, -0.2).normalize();
    for v in [i, n, i.reflect(n), i.refract(n, 1.0 / 1.5), i.refract(n, 1.5), (-i).refract(n, 1.5)] {
        misc.extend_from_slice(&[v.x, v.y, v.z]);
    }
    write_f64s(&dump.join("glam_misc.bin"), &misc)?;
    println!("perm / textures / dmat4_inverse / glam_misc: ok");
    Ok(())
}
