// Links libnrrt_b200.so.  NRRT_B200_LIB_DIR = directory holding the library (default: the in-tree build,
// ../../nr_ray_tracer_b200 relative to this crate, produced by `python -m nr_ray_tracer_b200.build`).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("NRRT_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../nr_ray_tracer_b200")
    });
    let dir = dir.canonicalize().unwrap_or(dir);
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=nrrt_b200");
    // the library is not installed system-wide: let binaries find it where it was built
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=NRRT_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=build.rs");
}
