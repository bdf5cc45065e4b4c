// UNVERIFIED (never compiled here).  Link against libhnsw_b200.so built by `make -C hnswlib-rs_b200/csrc`.
fn main() {
    let dir = std::env::var("HNSW_B200_LIB_DIR").unwrap_or_else(|_| "../lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=hnsw_b200");
}
