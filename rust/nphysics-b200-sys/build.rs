// Compiles the sm_100a kernels with nvcc (same flags as nphysics_b200/csrc/Makefile) and links them.
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../nphysics_b200/csrc");
    // (source, keep the reference's multiply-add order).  solve_coloured.cu holds the coloured-mode staged
    // kernels, which may contract to FMA (csrc/Makefile).
    let srcs = [("api.cu", true), ("bodies.cu", true), ("schedule.cu", true), ("assemble.cu", true), ("assemble_coloured.cu", true), ("solve.cu", true),
                ("activation.cu", true), ("narrowphase.cu", true), ("multibody.cu", true), ("solve_coloured.cu", false)];
    let mut objs = Vec::new();
    for (s, no_fma) in srcs.iter() {
        let src = csrc.join(s);
        println!("cargo:rerun-if-changed={}", src.display());
        let obj = out.join(s).with_extension("o");
        let mut args = vec!["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"];
        if *no_fma { args.push("-fmad=false"); }
        args.extend_from_slice(&["-Xcompiler", "-fPIC", "--extended-lambda", "-c", "-o"]);
        let status = Command::new("nvcc")
            .args(&args)
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
