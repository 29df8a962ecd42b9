//! `dist::compute_hv_ani` + the sort / filter of `utils::dump_ani_file` under `cuda-sketch-blackwell`
//! (replaces src/dist.rs:231-294 and src/utils.rs:262-285; `dist::dist`, src/dist.rs:11-63, keeps loading the files).
//!
//! The sketch file's payload goes to the GPUs as it is (bit-packed `hv`, `hv_quant_bits`, `hv_norm_2`):
//! `decompress_file_sketch` (src/hd.rs:171-232), every pair's i32 dot and f32 ANI (src/dist.rs:139-161), the
//! `ani >= ani_th` filter and the output order run there, rows sharded over all GPUs of the box.
use crate::hypergen_b200_sys::*;
use crate::types::{FileSketch, SketchDist};

/// rows of 2 * hv_d bytes (the first hv_quant_bits * hv_d / 8 live), quant bits, norms
fn stack_packed(s: &[FileSketch], d: usize) -> (Vec<u8>, Vec<u8>, Vec<i32>) {
    let mut p = vec![0u8; s.len() * 2 * d];
    for (i, f) in s.iter().enumerate() {
        for (k, v) in f.hv.iter().enumerate() { p[i * 2 * d + 2 * k..i * 2 * d + 2 * k + 2].copy_from_slice(&v.to_le_bytes()); }
    }
    (p, s.iter().map(|f| f.hv_quant_bits).collect(), s.iter().map(|f| f.hv_norm_2).collect())
}

#[cfg(feature = "cuda-sketch-blackwell")]
pub fn compute_hv_ani_gpu(sd: &mut SketchDist, r: &[FileSketch], q: &[FileSketch], ksize: u8, sym: bool) -> Vec<u32> {
    let d = r[0].hv_d;
    let (rp, rb, rn) = stack_packed(r, d);
    let own_q = if sym { None } else { Some(stack_packed(q, d)) };
    // same file on both sides (src/dist.rs:13): pass the SAME pointers - the rows move once and only j > i is walked
    let (qp, qb, qn) = match &own_q { Some((p, b, n)) => (p, b, n), None => (&rp, &rb, &rn) };
    let mut group = std::ptr::null_mut();
    check(unsafe { hg_group_create(0, std::ptr::null(), &mut group) });
    let mut cap = 1usize << 20;
    let (hits, milli) = loop {
        let (mut hits, mut milli, mut n) = (vec![hg_hit::default(); cap], vec![0u32; cap], 0u64);
        let rc = unsafe { hg_group_dist_packed(group, rp.as_ptr(), 2 * d as u64, rb.as_ptr(), rn.as_ptr(), r.len() as u32,
                                               qp.as_ptr(), 2 * d as u64, qb.as_ptr(), qn.as_ptr(), q.len() as u32, d as u32,
                                               ksize as u32, sd.ani_threshold, sym as i32, /*sorted*/ 1, hits.as_mut_ptr(),
                                               milli.as_mut_ptr(), cap as u64, &mut n) };
        if rc == HG_E_CAPACITY { cap = n as usize; continue; }
        check(rc);
        hits.truncate(n as usize);
        milli.truncate(n as usize);
        break (hits, milli);
    };
    unsafe { hg_group_destroy(group) };
    // already in dump_ani_file's order (ANI descending, ties by descending pair index, utils.rs:262-269); only pairs with
    // ani >= ani_th are returned (utils.rs:275).  milli[t] is the `{:.3}` field in thousandths:
    //   write!(f, "{}\t{}\t{}.{:03}\n", ref, query, m / 1000, m % 1000)
    sd.file_ani = hits.iter().map(|h| ((r[h.i as usize].file_str.clone(), q[h.j as usize].file_str.clone()), h.ani)).collect();
    milli
}
