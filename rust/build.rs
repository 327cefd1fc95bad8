// Compiles the CUDA sources for sm_100a and links them as libfemgpu (same flags as csrc/Makefile).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("finite_element_method_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = vec![];
    for (file, extra) in [("api.cu", None), ("prep.cu", Some("-fmad=false")), ("symbolic.cu", None),
                          ("numeric.cu", None), ("dist.cu", None), ("separate.cu", None),
                          ("results.cu", Some("-fmad=false")), ("solve.cu", None)] {
        let obj = out.join(file).with_extension("o");
        let mut c = Command::new(&nvcc);
        c.args(["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                "-gencode", "arch=compute_100a,code=sm_100a", "-c"]);
        if let Some(e) = extra { c.arg(e); }
        c.arg(csrc.join(file)).arg("-o").arg(&obj);
        assert!(c.status().expect("nvcc not found").success(), "nvcc failed on {file}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(file).display());
    }
    let lib = out.join("libfemgpu.so");
    let mut l = Command::new(&nvcc);
    l.args(["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o"]).arg(&lib).args(&objs)
        .args(["-lcudart", "-ldl"]);
    assert!(l.status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=femgpu");
    println!("cargo:rerun-if-changed={}", root.join("include/femgpu.h").display());
}
