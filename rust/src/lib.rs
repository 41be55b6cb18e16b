//! `cbl` — the reference crate's public surface (imartayan/CBL `src/lib.rs:4-14`) over the B200 library.
//!
//! `kmer` and `necklace` stay pure-Rust host utilities exactly as in the reference (they are part of the public API and
//! are not reproduced here: copy `src/kmer.rs` and `src/necklace/` of the reference next to these files); everything
//! below `CBL` — `src/ffi.rs` (autocxx bindings of sux / tiered-vector), `src/bitvector`, `src/wordset`, `src/trievec`,
//! `src/trie.rs`, `src/sliced_int.rs` — is replaced by [`ffi`], the `extern "C"` block of `include/cbl_gpu.h`.
//!
//! Not compiled in this repository's image (no Rust toolchain there); see INTEGRATION.md.
pub mod cbl;
pub mod ffi;

pub use crate::cbl::{Word, CBL};
