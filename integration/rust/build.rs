// build.rs for pathtracer-rs with the `b200` feature: builds libptrs_b200.so from the pathtracer-b200 checkout
// named by PTRS_B200_DIR and links it.  Takes the place of the nvcc -> PTX step of the reference's build.rs:14-36
// (the OptiX stub).  Not compiled in the pathtracer-b200 repository (no Rust toolchain in that image).
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var_os("CARGO_FEATURE_B200").is_none() {
        return;
    }
    let root = PathBuf::from(env::var("PTRS_B200_DIR").expect("set PTRS_B200_DIR to the pathtracer-b200 checkout"));
    let status = Command::new("make")
        .arg("-C")
        .arg(&root)
        .arg("pathtracer_rs_b200/lib/libptrs_b200.so")
        .status()
        .expect("failed to run make");
    assert!(status.success(), "building libptrs_b200.so failed");
    let lib_dir = root.join("pathtracer_rs_b200/lib");
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=ptrs_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-changed={}", root.join("include/ptrs_b200.h").display());
    println!("cargo:rerun-if-env-changed=PTRS_B200_DIR");
}
