//! The C ABI of libcbl_gpu (include/cbl_gpu.h).  Replaces the reference's `include_cpp!` block (src/ffi.rs:7-20), which
//! bound `RankBV` and `TieredVec32` and was crossed per prefix group / per element: this boundary is crossed once per
//! sequence, batch or set operation.
#![allow(non_camel_case_types)]
use core::ffi::c_char;

#[repr(C)]
pub struct cbl_t {
    _private: [u8; 0],
}

pub const CBL_OK: i32 = 0;
pub const CBL_EINVAL: i32 = 1;
pub const CBL_ECUDA: i32 = 2;
pub const CBL_ENOMEM: i32 = 3;
pub const CBL_ENCCL: i32 = 4;
pub const CBL_EIO: i32 = 5;
pub const CBL_OP_OR: i32 = 0;
pub const CBL_OP_AND: i32 = 1;
pub const CBL_OP_SUB: i32 = 2;
pub const CBL_OP_XOR: i32 = 3;

extern "C" {
    // life cycle: CBL::new / new_canonical (src/cbl.rs:71-79), Clone, Drop
    pub fn cbl_create(k: u32, word_bits: u32, prefix_bits: u32, canonical: i32, device: i32, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_create_sharded(k: u32, word_bits: u32, prefix_bits: u32, canonical: i32, n_gpus: i32, devices: *const i32, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_destroy(h: *mut cbl_t) -> i32;
    pub fn cbl_clone(h: *mut cbl_t, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_last_error(h: *const cbl_t) -> *const c_char;
    pub fn cbl_last_global_error() -> *const c_char;
    // scalar queries (src/cbl.rs:162-177)
    pub fn cbl_count(h: *const cbl_t, out: *mut u64) -> i32;
    pub fn cbl_is_empty(h: *const cbl_t, out: *mut i32) -> i32;
    pub fn cbl_is_canonical(h: *const cbl_t, out: *mut i32) -> i32;
    pub fn cbl_num_buckets(h: *const cbl_t, out: *mut u64) -> i32;
    // sequences (src/cbl.rs:293-354)
    pub fn cbl_insert_seq(h: *mut cbl_t, seq: *const u8, len: usize) -> i32;
    pub fn cbl_remove_seq(h: *mut cbl_t, seq: *const u8, len: usize) -> i32;
    pub fn cbl_contains_seq(h: *mut cbl_t, seq: *const u8, len: usize, out: *mut u8, n_out: *mut usize) -> i32;
    pub fn cbl_contains_all(h: *mut cbl_t, seq: *const u8, len: usize, out: *mut i32) -> i32;
    // whole record loops in one call (examples/cbl.rs:160-163, 216-228)
    pub fn cbl_insert_seqs(h: *mut cbl_t, buf: *const u8, offsets: *const u64, n_seqs: usize) -> i32;
    pub fn cbl_remove_seqs(h: *mut cbl_t, buf: *const u8, offsets: *const u64, n_seqs: usize) -> i32;
    pub fn cbl_contains_seqs(h: *mut cbl_t, buf: *const u8, offsets: *const u64, n_seqs: usize, out: *mut u8) -> i32;
    pub fn cbl_count_kmers(h: *const cbl_t, offsets: *const u64, n_seqs: usize, out: *mut u64) -> i32;
    pub fn cbl_last_kmer_count(h: *const cbl_t, out: *mut u64) -> i32;
    // single k-mers (src/cbl.rs:219-235), batched; out[i] = membership BEFORE the call
    pub fn cbl_contains_kmers(h: *mut cbl_t, lo: *const u64, hi: *const u64, n: usize, out: *mut u8) -> i32;
    pub fn cbl_insert_kmers(h: *mut cbl_t, lo: *const u64, hi: *const u64, n: usize, out: *mut u8) -> i32;
    pub fn cbl_remove_kmers(h: *mut cbl_t, lo: *const u64, hi: *const u64, n: usize, out: *mut u8) -> i32;
    // set operations (src/cbl.rs:411-569, 108-124)
    pub fn cbl_setop(op: i32, a: *mut cbl_t, b: *mut cbl_t, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_setop_assign(op: i32, a: *mut cbl_t, b: *mut cbl_t) -> i32;
    pub fn cbl_merge_many(hs: *mut *mut cbl_t, n: usize, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_intersect_many(hs: *mut *mut cbl_t, n: usize, out: *mut *mut cbl_t) -> i32;
    // iteration and statistics (src/cbl.rs:358-396)
    pub fn cbl_export_words(h: *mut cbl_t, start: u64, lo: *mut u64, hi: *mut u64, cap: usize, n_out: *mut usize) -> i32;
    pub fn cbl_export_kmers(h: *mut cbl_t, start: u64, lo: *mut u64, hi: *mut u64, cap: usize, n_out: *mut usize) -> i32;
    pub fn cbl_bucket_sizes(h: *mut cbl_t, prefixes: *mut u32, sizes: *mut u32, cap: usize, n_out: *mut usize) -> i32;
    // serde (src/cbl.rs:127-160)
    pub fn cbl_serialize_size(h: *mut cbl_t, out: *mut usize) -> i32;
    pub fn cbl_serialize(h: *mut cbl_t, out: *mut u8, cap: usize, n_out: *mut usize) -> i32;
    pub fn cbl_deserialize(proto: *const cbl_t, data: *const u8, len: usize, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_save_to_file(h: *mut cbl_t, path: *const c_char) -> i32;
    pub fn cbl_load_from_file(proto: *const cbl_t, path: *const c_char, out: *mut *mut cbl_t) -> i32;
    pub fn cbl_sync(h: *mut cbl_t) -> i32;
}

/// Turns a non-zero status into the panic the reference would have raised (same message: the library words its errors
/// like src/cbl.rs:87-91, 294-299, 329-334, 422-425).
///
/// # Safety
/// `h` must be null or a live handle.
pub unsafe fn check(h: *const cbl_t, rc: i32) {
    if rc != CBL_OK {
        let p = if h.is_null() { cbl_last_global_error() } else { cbl_last_error(h) };
        let msg = std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}
