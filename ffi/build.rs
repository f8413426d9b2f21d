// Build libh2v for sm_100a with nvcc and link it statically-by-path.  No CPU fallback is compiled.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("halo2_vectordb_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libh2v.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(["-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", "-split-compile", "0"])
        .args(["-gencode", "arch=compute_100a,code=sm_100a"])
        .arg(format!("-I{}", root.join("include").display()))
        .arg(format!("-I{}", csrc.display()))
        .arg("-o")
        .arg(&lib)
        .arg(csrc.join("h2v.cu"))
        .status()
        .expect("nvcc not found: libh2v has no CPU fallback, a CUDA 12.9+ toolkit is required");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=h2v");
    for f in ["h2v.cu", "msm.cuh", "ntt.cuh", "ec.cuh", "ff.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("h2v.h").display());
}
