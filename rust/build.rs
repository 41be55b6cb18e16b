// Replaces the reference's build.rs:70-86 (autocxx_build over cxx/rank_bv.h + cxx/tiered_vec.h): nothing is compiled,
// the crate links the prebuilt CUDA library.  K / PREFIX_BITS stay compile-time parameters of `CBL<K, T, PREFIX_BITS>`
// (const generics), so the env-var plumbing of build.rs:9-57 is only needed by examples/cbl.rs and is unchanged there.
fn main() {
    let dir = std::env::var("CBL_GPU_LIB_DIR").unwrap_or_else(|_| "../cbl_b200/csrc".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cbl_gpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-changed=../include/cbl_gpu.h");
    println!("cargo:rerun-if-env-changed=CBL_GPU_LIB_DIR");
}
