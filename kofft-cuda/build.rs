//! Compiles the hand-written CUDA kernels with `nvcc -gencode arch=compute_100a,code=sm_100a`
//! into a static library and links it together with cudart.
//!
//! The kernel sources live in `../kofft_b200/csrc` of this repository (the same files the
//! Python mirror builds into `libkofft_cuda.so`):
//!   fft_inst.cu (once per L = 5..14 with -DKOFFT_L), small_inst.cu, fft_large_inst.cu, istft_inst.cu,
//!   ola.cu, dist_kernels.cu, bluestein.cu, kofft_cuda.cu (the C ABI) and host_tables.cpp (bit-exact twiddle/window generators; must
//!   be built with -ffp-contract=off).
use std::{env, path::PathBuf, process::Command};

fn run(cmd: &mut Command) {
    let status = cmd.status().unwrap_or_else(|e| panic!("failed to spawn {:?}: {e}", cmd));
    assert!(status.success(), "{:?} failed", cmd);
}

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../kofft_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let cuda_lib = env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".into());
    let flags = [
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-ffp-contract=off",
    ];
    let mut objs = Vec::new();
    for l in 5..=14 {
        let o = out.join(format!("fft_L{l}.o"));
        run(Command::new(&nvcc).args(flags).arg(format!("-DKOFFT_L={l}")).arg("-c")
            .arg(csrc.join("fft_inst.cu")).arg("-o").arg(&o));
        objs.push(o);
    }
    for src in ["small_inst.cu", "fft_large_inst.cu", "fft_f64_inst.cu", "istft_inst.cu", "ola.cu", "dist_kernels.cu", "bluestein.cu",
                "kofft_cuda.cu"] {
        let o = out.join(format!("{src}.o"));
        run(Command::new(&nvcc).args(flags).arg("-c").arg(csrc.join(src)).arg("-o").arg(&o));
        objs.push(o);
    }
    let o = out.join("host_tables.o");
    run(Command::new(env::var("CXX").unwrap_or_else(|_| "g++".into()))
        .args(["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-c"])
        .arg(csrc.join("host_tables.cpp")).arg("-o").arg(&o));
    objs.push(o);
    let lib = out.join("libkofft_cuda.a");
    let _ = std::fs::remove_file(&lib);
    run(Command::new("ar").arg("crs").arg(&lib).args(&objs));

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=kofft_cuda");
    println!("cargo:rustc-link-search=native={cuda_lib}");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed=../include/kofft_cuda.h");
}
