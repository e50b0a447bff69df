// build.rs — compiles the CUDA sources of this repository with nvcc for
// sm_100a and links the resulting shared library.  NOT COMPILED HERE (no
// cargo in the build image); mirrors coupe_b200/csrc/Makefile line by line.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("coupe_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libcoupe_b200.so");

    if env::var_os("CARGO_FEATURE_PREBUILT").is_none() {
        let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
        let status = Command::new(nvcc)
            .args([
                "-gencode", "arch=compute_100a,code=sm_100a",
                "-O3", "-std=c++17", "-lineinfo",
                // the bisection arithmetic must round like the reference's f32/f64 code
                "-fmad=false",
                "-Xcompiler", "-fPIC,-O3",
                "-shared", "-o",
            ])
            .arg(&lib)
            .arg(csrc.join("engine.cu"))
            .arg(csrc.join("ffi.cu"))
            .arg(csrc.join("tools.cu"))
            .arg(csrc.join("mj.cu"))
            .arg(csrc.join("grid.cu"))
            .args(["-ldl", "-lpthread"])
            .status()
            .expect("nvcc not found: set NVCC or enable the `prebuilt` feature");
        assert!(status.success(), "nvcc failed");
        println!("cargo:rustc-link-search=native={}", out.display());
    } else {
        println!("cargo:rustc-link-search=native={}", root.join("coupe_b200/lib").display());
    }
    println!("cargo:rustc-link-lib=dylib=coupe_b200");
    for f in ["engine.cu", "ffi.cu", "tools.cu", "mj.cu", "grid.cu", "rcb_kernels.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/coupe_b200.h").display());
}
