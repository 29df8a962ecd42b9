// build.rs arm for `--features cuda-sketch-blackwell` (goes in front of the reference's PTX arm, build.rs:9-64,
// which stays as the else branch for the older features).
//
// The reference compiles src/cuda_kernel.cu to PTX and JIT-loads it through cudarc (build.rs:31-39,
// src/sketch_cuda.rs:19-20,52-60).  Here nvcc compiles the CUDA sources of libhypergen_b200 (vendored under
// cuda/b200/) for sm_100a into a static library, and bindgen - already a build-dependency, today pointed at the
// empty src/cuda_kernel.h (build.rs:46-61) - reads the real header.
use std::{env, path::PathBuf};

fn main() {
    if cfg!(feature = "cuda-sketch-blackwell") {
        let out_dir = PathBuf::from(env::var("OUT_DIR").unwrap());
        let dir = PathBuf::from("cuda/b200");
        println!("cargo:rerun-if-changed={}", dir.display());
        let sources = [
            "api.cu", "kmer_hash.cu", "encode.cu", "dist_simt.cu", "dist_tc.cu", "dist_narrow.cu", "probe.cu",
            "fasta.cu", "sort.cu", "peer.cu",
        ];
        cc::Build::new()
            .cuda(true)
            .cudart("static")
            .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
            .flag("-lineinfo").flag("-O3").flag("-std=c++17")
            .include(dir.join("include"))
            .files(sources.iter().map(|f| dir.join("csrc").join(f)))
            .compile("hypergen_b200");
        println!("cargo:rustc-link-lib=cuda"); // cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint
        let bindings = bindgen::Builder::default()
            .header(dir.join("include/hypergen_b200.h").to_str().unwrap())
            .allowlist_function("hg_.*")
            .allowlist_type("hg_.*")
            .generate()
            .expect("bindgen over hypergen_b200.h");
        bindings.write_to_file(out_dir.join("hg_bindings.rs")).unwrap();
        return;
    }
    // ... the reference's existing PTX arm (build.rs:9-64) follows unchanged ...
}
