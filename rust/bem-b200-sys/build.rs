// build.rs -- compile the CUDA translation units with nvcc for sm_100a and link them.
// Mirrors math_audio_b200/build.py (the Python build used in this repository's CI image).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("math_audio_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    // (translation unit, extra flags): the decision-taking kernels forbid FMA contraction
    let units: [(&str, &[&str]); 13] = [
        ("assembly_exact.cu", &["-fmad=false"]),
        ("assembly_far.cu", &[]),
        ("linalg.cu", &[]),
        ("gmres.cu", &[]),
        ("gmres_fused.cu", &[]),
        ("block_gmres.cu", &[]),
        ("schwarz.cu", &[]),
        ("block_matvec.cu", &[]),
        ("postprocess.cu", &["-fmad=false"]),
        ("room.cu", &[]),
        ("direct.cu", &[]),
        ("sweep.cu", &[]),
        ("api.cu", &[]),
    ];
    let mut objs = Vec::new();
    for (unit, extra) in units.iter() {
        let obj = out.join(format!("{unit}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
            .args(["-Xcompiler", "-fPIC,-ffp-contract=off"])
            .args(extra.iter())
            .arg("-c").arg(csrc.join(unit)).arg("-o").arg(&obj)
            .status().expect("nvcc not found: there is no CPU fallback for libbemb200");
        assert!(status.success(), "nvcc failed on {unit}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(unit).display());
    }
    let lib = out.join("libbemb200.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"]).arg(&lib)
        .args(&objs).args(["-Xcompiler", "-fPIC", "-ldl"]).status().unwrap();
    assert!(status.success(), "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=bemb200");
    println!("cargo:rerun-if-changed={}", root.join("include/bemb200.h").display());
}
