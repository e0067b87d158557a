// Compiles the sm_100a kernels with nvcc (same flags as nphysics_b200/csrc/Makefile) and links them.
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../nphysics_b200/csrc");
    let srcs = ["api.cu", "bodies.cu", "schedule.cu", "assemble.cu", "solve.cu"];
    let mut objs = Vec::new();
    for s in srcs.iter() {
        let src = csrc.join(s);
        println!("cargo:rerun-if-changed={}", src.display());
        let obj = out.join(s).with_extension("o");
        let status = Command::new("nvcc")
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                    "-fmad=false", "-Xcompiler", "-fPIC", "--extended-lambda", "-c", "-o"])
            .arg(&obj)
            .arg(&src)
            .status()
            .expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {}", s);
        objs.push(obj);
    }
    let lib = out.join("libnphysics_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().expect("ar not found");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=nphysics_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
