//! `CBL<K, T, PREFIX_BITS>` with the reference's public API (src/cbl.rs:40-569), every body a call into libcbl_gpu.
//! k-mers cross the boundary as `IntKmer` integers (first base most significant, A=0 C=1 T=2 G=3: src/kmer.rs:11-24), as
//! two 64-bit halves.  Where the reference panics, the library returns a status and `ffi::check` panics with the same
//! message, so callers (examples/cbl.rs, the crate's own tests) observe the same behaviour.
use crate::ffi;
use core::marker::PhantomData;
use core::ops::{BitAnd, BitAndAssign, BitOr, BitOrAssign, BitXor, BitXorAssign, Sub, SubAssign};
use std::ffi::CString;
use std::path::Path;

/// The integer types the reference instantiates `T` with (u32 / u64 / u128).
pub trait Word: Copy {
    const BITS: u32;
    fn split(self) -> (u64, u64);
    fn join(lo: u64, hi: u64) -> Self;
}
impl Word for u32 {
    const BITS: u32 = 32;
    fn split(self) -> (u64, u64) { (self as u64, 0) }
    fn join(lo: u64, _hi: u64) -> Self { lo as u32 }
}
impl Word for u64 {
    const BITS: u32 = 64;
    fn split(self) -> (u64, u64) { (self, 0) }
    fn join(lo: u64, _hi: u64) -> Self { lo }
}
impl Word for u128 {
    const BITS: u32 = 128;
    fn split(self) -> (u64, u64) { (self as u64, (self >> 64) as u64) }
    fn join(lo: u64, hi: u64) -> Self { ((hi as u128) << 64) | lo as u128 }
}

/// A fully dynamic set of k-mers resident in GPU memory.  `!Send + !Sync` like the reference (raw pointer field); a
/// handle is used through `&mut self`.
pub struct CBL<const K: usize, T: Word, const PREFIX_BITS: usize = 24> {
    h: *mut ffi::cbl_t,
    _t: PhantomData<T>,
}

impl<const K: usize, T: Word, const PREFIX_BITS: usize> CBL<K, T, PREFIX_BITS> {
    fn from_raw(h: *mut ffi::cbl_t) -> Self { Self { h, _t: PhantomData } }
    fn create(canonical: bool) -> Self {
        let mut h = core::ptr::null_mut();
        // parameter checks (src/cbl.rs:87-91, src/wordset/mod.rs:37-41) happen in the library, with the same messages
        unsafe { ffi::check(core::ptr::null(), ffi::cbl_create(K as u32, T::BITS, PREFIX_BITS as u32, canonical as i32, 0, &mut h)) };
        Self::from_raw(h)
    }
    /// src/cbl.rs:71-73
    pub fn new() -> Self { Self::create(false) }
    /// src/cbl.rs:77-79
    pub fn new_canonical() -> Self { Self::create(true) }
    /// The same set prefix-sharded over several GPUs of this process; every method below works unchanged.
    pub fn new_sharded(devices: &[i32], canonical: bool) -> Self {
        let mut h = core::ptr::null_mut();
        unsafe {
            ffi::check(core::ptr::null(), ffi::cbl_create_sharded(K as u32, T::BITS, PREFIX_BITS as u32, canonical as i32, devices.len() as i32, devices.as_ptr(), &mut h))
        };
        Self::from_raw(h)
    }

    /// src/cbl.rs:108-115 (k-way union; `canonical` taken from element 0)
    pub fn merge(cbls: Vec<&mut Self>) -> Self { Self::many(cbls, false) }
    /// src/cbl.rs:117-124
    pub fn intersect(cbls: Vec<&mut Self>) -> Self { Self::many(cbls, true) }
    fn many(cbls: Vec<&mut Self>, intersect: bool) -> Self {
        assert!(!cbls.is_empty());
        let mut hs: Vec<*mut ffi::cbl_t> = cbls.iter().map(|c| c.h).collect();
        let mut out = core::ptr::null_mut();
        unsafe {
            let rc = if intersect { ffi::cbl_intersect_many(hs.as_mut_ptr(), hs.len(), &mut out) } else { ffi::cbl_merge_many(hs.as_mut_ptr(), hs.len(), &mut out) };
            ffi::check(hs[0], rc);
        }
        Self::from_raw(out)
    }

    /// src/cbl.rs:127-141
    pub fn save_to_file<P: AsRef<Path> + Copy>(&self, path: P) {
        let p = CString::new(path.as_ref().to_str().expect("path")).unwrap();
        unsafe { ffi::check(self.h, ffi::cbl_save_to_file(self.h, p.as_ptr())) };
    }
    /// src/cbl.rs:143-160
    pub fn load_from_file<P: AsRef<Path> + Copy>(path: P) -> Self {
        let proto = Self::new();
        let p = CString::new(path.as_ref().to_str().expect("path")).unwrap();
        let mut out = core::ptr::null_mut();
        unsafe { ffi::check(proto.h, ffi::cbl_load_from_file(proto.h, p.as_ptr(), &mut out)) };
        Self::from_raw(out)
    }

    /// src/cbl.rs:162-177
    pub fn is_canonical(&self) -> bool { let mut v = 0; unsafe { ffi::check(self.h, ffi::cbl_is_canonical(self.h, &mut v)) }; v != 0 }
    pub fn count(&self) -> usize { let mut n = 0u64; unsafe { ffi::check(self.h, ffi::cbl_count(self.h, &mut n)) }; n as usize }
    pub fn is_empty(&self) -> bool { let mut v = 0; unsafe { ffi::check(self.h, ffi::cbl_is_empty(self.h, &mut v)) }; v != 0 }

    /// src/cbl.rs:219-235.  `kmer` = `IntKmer::<K, T>::to_int()`.
    pub fn contains(&mut self, kmer: T) -> bool { self.kmer_op(ffi::cbl_contains_kmers, kmer) }
    /// Returns true if the k-mer was absent.
    pub fn insert(&mut self, kmer: T) -> bool { !self.kmer_op(ffi::cbl_insert_kmers, kmer) }
    /// Returns true if the k-mer was present.
    pub fn remove(&mut self, kmer: T) -> bool { self.kmer_op(ffi::cbl_remove_kmers, kmer) }
    fn kmer_op(&mut self, f: unsafe extern "C" fn(*mut ffi::cbl_t, *const u64, *const u64, usize, *mut u8) -> i32, kmer: T) -> bool {
        let (lo, hi) = kmer.split();
        let mut was = 0u8;
        unsafe { ffi::check(self.h, f(self.h, &lo, &hi, 1, &mut was)) };
        was != 0
    }

    /// src/cbl.rs:293-309
    pub fn contains_all(&mut self, seq: &[u8]) -> bool {
        let mut v = 0;
        unsafe { ffi::check(self.h, ffi::cbl_contains_all(self.h, seq.as_ptr(), seq.len(), &mut v)) };
        v != 0
    }
    /// src/cbl.rs:311-324: one answer per k-mer, per 2048-k-mer chunk the forward-canonical k-mers first (src/cbl.rs:248-275).
    pub fn contains_seq(&mut self, seq: &[u8]) -> Vec<bool> {
        let mut out = vec![0u8; seq.len().saturating_sub(K) + 1];
        let mut n = 0usize;
        unsafe { ffi::check(self.h, ffi::cbl_contains_seq(self.h, seq.as_ptr(), seq.len(), out.as_mut_ptr(), &mut n)) };
        out.truncate(n);
        out.into_iter().map(|b| b != 0).collect()
    }
    /// src/cbl.rs:328-339
    pub fn insert_seq(&mut self, seq: &[u8]) { unsafe { ffi::check(self.h, ffi::cbl_insert_seq(self.h, seq.as_ptr(), seq.len())) } }
    /// src/cbl.rs:343-354
    pub fn remove_seq(&mut self, seq: &[u8]) { unsafe { ffi::check(self.h, ffi::cbl_remove_seq(self.h, seq.as_ptr(), seq.len())) } }
    /// A whole file's record loop in one call: `buf` = the records back to back, `offsets` = n + 1 byte offsets.
    pub fn insert_seqs(&mut self, buf: &[u8], offsets: &[u64]) {
        unsafe { ffi::check(self.h, ffi::cbl_insert_seqs(self.h, buf.as_ptr(), offsets.as_ptr(), offsets.len() - 1)) }
    }
    pub fn contains_seqs(&mut self, buf: &[u8], offsets: &[u64]) -> Vec<bool> {
        let mut n = 0u64;
        unsafe { ffi::check(self.h, ffi::cbl_count_kmers(self.h, offsets.as_ptr(), offsets.len() - 1, &mut n)) };
        let mut out = vec![0u8; n as usize + 1];
        unsafe { ffi::check(self.h, ffi::cbl_contains_seqs(self.h, buf.as_ptr(), offsets.as_ptr(), offsets.len() - 1, out.as_mut_ptr())) };
        unsafe { ffi::check(self.h, ffi::cbl_last_kmer_count(self.h, &mut n)) };
        out.truncate(n as usize);
        out.into_iter().map(|b| b != 0).collect()
    }

    /// src/cbl.rs:358-360: the stored k-mers (`IntKmer` integers) in ascending word order.
    pub fn iter(&self) -> impl Iterator<Item = T> + '_ {
        const CH: usize = 1 << 20;
        let h = self.h;
        let mut start = 0u64;
        let (mut lo, mut hi) = (vec![0u64; CH], vec![0u64; CH]);
        let (mut at, mut have) = (0usize, 0usize);
        core::iter::from_fn(move || {
            if at == have {
                let mut n = 0usize;
                unsafe { ffi::check(h, ffi::cbl_export_kmers(h, start, lo.as_mut_ptr(), hi.as_mut_ptr(), CH, &mut n)) };
                if n == 0 { return None; }
                start += n as u64;
                at = 0;
                have = n;
            }
            at += 1;
            Some(T::join(lo[at - 1], hi[at - 1]))
        })
    }

    /// src/cbl.rs:364-372
    pub fn prefix_load(&self) -> f64 {
        let mut nb = 0u64;
        unsafe { ffi::check(self.h, ffi::cbl_num_buckets(self.h, &mut nb)) };
        nb as f64 / (1u64 << PREFIX_BITS) as f64
    }
    pub fn buckets_sizes(&self) -> impl Iterator<Item = (usize, usize)> {
        let mut n = 0usize;
        unsafe { ffi::check(self.h, ffi::cbl_bucket_sizes(self.h, core::ptr::null_mut(), core::ptr::null_mut(), 0, &mut n)) };
        let (mut p, mut s) = (vec![0u32; n.max(1)], vec![0u32; n.max(1)]);
        if n > 0 { unsafe { ffi::check(self.h, ffi::cbl_bucket_sizes(self.h, p.as_mut_ptr(), s.as_mut_ptr(), n, &mut n)) }; }
        p.truncate(n);
        s.truncate(n);
        p.into_iter().zip(s).map(|(a, b)| (a as usize, b as usize))
    }
    /// src/cbl.rs:374-377: bucket size -> number of buckets of that size
    pub fn buckets_size_count(&self) -> std::collections::BTreeMap<usize, usize> {
        let mut m = std::collections::BTreeMap::new();
        for (_, s) in self.buckets_sizes() { *m.entry(s).or_insert(0) += 1; }
        m
    }
    /// src/cbl.rs:379-396: share of the elements held by buckets of each size
    pub fn buckets_load_repartition(&self) -> std::collections::BTreeMap<usize, f64> {
        let total = self.count().max(1) as f64;
        self.buckets_size_count().into_iter().map(|(size, n)| (size, (size * n) as f64 / total)).collect()
    }

    fn binary(&mut self, op: i32, other: &mut Self) -> Self {
        let mut out = core::ptr::null_mut();
        unsafe { ffi::check(self.h, ffi::cbl_setop(op, self.h, other.h, &mut out)) };   // canonical mismatch panics: src/cbl.rs:422-425
        Self::from_raw(out)
    }
    fn assign(&mut self, op: i32, other: &mut Self) { unsafe { ffi::check(self.h, ffi::cbl_setop_assign(op, self.h, other.h)) } }
}

impl<const K: usize, T: Word, const P: usize> Default for CBL<K, T, P> { fn default() -> Self { Self::new() } }
impl<const K: usize, T: Word, const P: usize> Drop for CBL<K, T, P> { fn drop(&mut self) { unsafe { ffi::cbl_destroy(self.h) }; } }
impl<const K: usize, T: Word, const P: usize> Clone for CBL<K, T, P> {
    fn clone(&self) -> Self {
        let mut out = core::ptr::null_mut();
        unsafe { ffi::check(self.h, ffi::cbl_clone(self.h, &mut out)) };
        Self::from_raw(out)
    }
}

// src/cbl.rs:411-569: | & - ^ on `&mut CBL` (out of place) and the four assign forms
macro_rules! set_op {
    ($tr:ident, $f:ident, $tra:ident, $fa:ident, $op:expr) => {
        impl<const K: usize, T: Word, const P: usize> $tr<Self> for &mut CBL<K, T, P> {
            type Output = CBL<K, T, P>;
            fn $f(self, other: Self) -> Self::Output { self.binary($op, other) }
        }
        impl<const K: usize, T: Word, const P: usize> $tra<&mut Self> for CBL<K, T, P> {
            fn $fa(&mut self, other: &mut Self) { self.assign($op, other) }
        }
    };
}
set_op!(BitOr, bitor, BitOrAssign, bitor_assign, ffi::CBL_OP_OR);
set_op!(BitAnd, bitand, BitAndAssign, bitand_assign, ffi::CBL_OP_AND);
set_op!(Sub, sub, SubAssign, sub_assign, ffi::CBL_OP_SUB);
set_op!(BitXor, bitxor, BitXorAssign, bitxor_assign, ffi::CBL_OP_XOR);
