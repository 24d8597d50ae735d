// build.rs — added to the reference's `engine` crate (engine/Cargo.toml:7-10 already declares the
// `cc` build-dependency; the reference never shipped a build.rs).  With `--features gpu` it builds
// librama_b200.a with nvcc for sm_100a only and links it plus cudart; NCCL is dlopen'ed at run time
// by the library itself (only when tensor parallelism is requested), so it is not linked here.
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var("CARGO_FEATURE_GPU").is_err() {
        return;
    }
    let rama = PathBuf::from(env::var("RAMA_B200_DIR").unwrap_or_else(|_| "../rama_b200".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let obj = out.join("rama_b200_api.o");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let ok = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", "-o"])
        .arg(&obj)
        .arg(rama.join("csrc/api.cu"))
        .status()
        .expect("nvcc not found")
        .success();
    assert!(ok, "nvcc failed");
    let ok = Command::new("ar").arg("rcs").arg(out.join("librama_b200.a")).arg(&obj).status().unwrap().success();
    assert!(ok, "ar failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=rama_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rerun-if-changed={}", rama.join("csrc").display());
    println!("cargo:rerun-if-changed={}", rama.join("../include/rama_b200.h").display());
}
