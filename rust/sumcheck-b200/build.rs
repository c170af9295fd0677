// SOURCE ONLY -- not compiled in this environment.
// Builds libsumcheck_b200 with nvcc for sm_100a from the CUDA sources of this repository and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("SUMCHECK_B200_ROOT").unwrap_or_else(|_| "../..".into()));
    let csrc = root.join("thaler_study_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"];
    for (src, obj, arch) in [("engine.cu", "engine.o", true), ("g4.cu", "g4.o", true), ("protocol.cpp", "protocol.o", false)] {
        let mut c = Command::new(&nvcc);
        if arch {
            c.args(["-gencode", "arch=compute_100a,code=sm_100a"]);
        }
        c.args(common)
            .arg(format!("-I{}", root.join("include").display()))
            .arg(format!("-I{}", csrc.display()))
            .args(["-c", "-o"])
            .arg(out.join(obj))
            .arg(csrc.join(src));
        assert!(c.status().expect("nvcc not found").success(), "nvcc failed on {src}");
    }
    let status = Command::new("ar")
        .args(["crs"])
        .arg(out.join("libsumcheck_b200.a"))
        .arg(out.join("engine.o"))
        .arg(out.join("g4.o"))
        .arg(out.join("protocol.o"))
        .status()
        .unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=sumcheck_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
