//! `sketch_cuda::sketch_cuda` under `cuda-sketch-blackwell` (replaces src/sketch_cuda.rs:43-117).
//!
//! The reference reads each file (`fastx_reader::read_merge_seq`, src/fastx_reader.rs:6-29), hashes it on GPU 0
//! (src/sketch_cuda.rs:119-166) and encodes / norms / compresses on the CPU (:85-100).  Here the raw file bytes of a
//! batch go to the library in one call; record merging, k-mer hashing, the hash set, HD encoding, the norm and the
//! bit-packing run on the GPUs (files split over all GPUs of the box), and the FileSketch records come back complete.
use crate::hypergen_b200_sys::*;
use crate::types::{FileSketch, SketchParams};
use crate::utils;
use rayon::prelude::*;

#[cfg(feature = "cuda-sketch-blackwell")]
pub fn sketch_cuda(params: SketchParams) {
    let files = utils::get_fasta_files(&params.path); // unchanged: record order of the sketch file (utils.rs:208-221)
    let mut group = std::ptr::null_mut();
    check(unsafe { hg_group_create(0, std::ptr::null(), &mut group) }); // 0: every visible GPU
    let p = hg_sketch_params { scaled: params.scaled, seed: params.seed, hv_d: params.hv_d as u32, ksize: params.ksize,
                               canonical: params.canonical as u8, reserved: [0; 2] };
    let d = params.hv_d;
    let mut all = Vec::with_capacity(files.len());
    // one page-locked staging buffer for the run: H2D from it runs at the PCIe rate and overlaps the kernels
    let (mut stage, mut stage_cap) = (std::ptr::null_mut::<u8>(), 0u64);
    for batch in files.chunks(256 * unsafe { hg_group_size(group) } as usize) {
        let sizes: Vec<u64> = batch.par_iter().map(|f| std::fs::metadata(f).unwrap().len()).collect();
        let mut file_off = vec![0u64; batch.len() + 1];
        for (i, s) in sizes.iter().enumerate() { file_off[i + 1] = file_off[i] + s; }
        if file_off[batch.len()] > stage_cap {
            check(unsafe { hg_host_free(stage as *mut _) });
            stage_cap = file_off[batch.len()] * 5 / 4;
            check(unsafe { hg_host_alloc(stage_cap, &mut stage as *mut *mut u8 as *mut *mut _) });
        }
        let stage_addr = stage as usize;
        batch.par_iter().enumerate().for_each(|(i, f)| { // file reading stays parallel (rayon, `-t` threads)
            let dst = unsafe { std::slice::from_raw_parts_mut((stage_addr + file_off[i] as usize) as *mut u8, sizes[i] as usize) };
            std::io::Read::read_exact(&mut std::fs::File::open(f).unwrap(), dst).unwrap();
        });
        let n = batch.len();
        let (mut packed, mut bits, mut norm2, mut nh) = (vec![0u8; n * 2 * d], vec![0u8; n], vec![0i32; n], vec![0u32; n]);
        check(unsafe { hg_group_sketch_fasta_batch(group, stage, file_off.as_ptr(), n as u32, &p, std::ptr::null_mut(),
                                                   packed.as_mut_ptr(), bits.as_mut_ptr(), norm2.as_mut_ptr(), nh.as_mut_ptr()) });
        for i in 0..n {
            let nbytes = bits[i] as usize * d / 8; // hd.rs:146: hv_quant_bits * hv_d / 8 live bytes
            let row = &packed[i * 2 * d..i * 2 * d + nbytes];
            all.push(FileSketch { ksize: params.ksize, scaled: params.scaled, seed: params.seed, canonical: params.canonical,
                hv_d: d, hv_quant_bits: bits[i], hv_norm_2: norm2[i], file_str: batch[i].display().to_string(),
                hv: row.chunks_exact(2).map(|b| i16::from_le_bytes([b[0], b[1]])).collect() }); // hd.rs:155-157
        }
    }
    unsafe { hg_host_free(stage as *mut _); hg_group_destroy(group) };
    utils::dump_sketch(&all, &params.out_file); // unchanged (utils.rs:234-249)
}
