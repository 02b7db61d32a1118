// Point rustc at the shared library built by `python __graft_entry__.py` (or BLOCK_ALIGNER_B200_LIB_DIR).
fn main() {
    let dir = std::env::var("BLOCK_ALIGNER_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../block_aligner_b200", here)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=block_aligner_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
