//! `extern "C"` surface of libhypergen_b200 (include/hypergen_b200.h) as bindgen emits it, written out by hand so the
//! shim reads without a build.  Takes the place of the cudarc calls of src/sketch_cuda.rs:52-60,134-156.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct hg_ctx { _private: [u8; 0] }
#[repr(C)] pub struct hg_group { _private: [u8; 0] }

/// FileSketch knobs that reach the kernels (src/types.rs:224-235; defaults src/types.rs:97-113)
#[repr(C)] #[derive(Clone, Copy)]
pub struct hg_sketch_params { pub scaled: u64, pub seed: u64, pub hv_d: u32, pub ksize: u8, pub canonical: u8, pub reserved: [u8; 2] }

/// one reported pair: indices into the ref / query records, exact i32 dot, f32 ANI of src/dist.rs:139-161
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct hg_hit { pub i: u32, pub j: u32, pub dot: i32, pub ani: f32 }

pub const HG_OK: c_int = 0;
pub const HG_E_CAPACITY: c_int = -3;

extern "C" {
    pub fn hg_last_error() -> *const c_char;
    pub fn hg_host_alloc(bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn hg_host_free(p: *mut c_void) -> c_int;
    // every GPU of the box behind one handle (the reference drives CudaDevice::new(0) only, src/sketch_cuda.rs:52)
    pub fn hg_group_create(n_devices: c_int, ordinals: *const c_int, out: *mut *mut hg_group) -> c_int;
    pub fn hg_group_destroy(g: *mut hg_group);
    pub fn hg_group_size(g: *const hg_group) -> c_int;
    pub fn hg_group_sketch_fasta_batch(g: *mut hg_group, raw: *const u8, file_off: *const u64, n_files: u32,
        p: *const hg_sketch_params, hv: *mut i16, packed: *mut u8, quant_bits: *mut u8, norm2: *mut i32,
        n_hashes: *mut u32) -> c_int;
    pub fn hg_group_dist_packed(g: *mut hg_group, ref_packed: *const u8, ref_stride: u64, ref_quant_bits: *const u8,
        ref_norm2: *const i32, n_ref: u32, qry_packed: *const u8, qry_stride: u64, qry_quant_bits: *const u8,
        qry_norm2: *const i32, n_qry: u32, hv_d: u32, ksize: u32, ani_th: f32, symmetric: c_int, sorted: c_int,
        hits: *mut hg_hit, ani_milli: *mut u32, cap: u64, n_hits: *mut u64) -> c_int;
}

/// status -> panic with the library's message, the reference's error style (`unwrap()`, Cargo.toml:67-70)
pub fn check(rc: c_int) {
    if rc != HG_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(hg_last_error()) }.to_string_lossy().into_owned();
        panic!("hypergen_b200 error {}: {}", rc, msg);
    }
}
