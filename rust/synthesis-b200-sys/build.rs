// Links libsynthesis_b200.so (built by `python __graft_entry__.py`, nvcc, sm_100a).
// SYNTHESIS_B200_LIB_DIR points at the directory that holds it.
fn main() {
    let dir = std::env::var("SYNTHESIS_B200_LIB_DIR").unwrap_or_else(|_| "../../synthesis_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=synthesis_b200");
    println!("cargo:rerun-if-env-changed=SYNTHESIS_B200_LIB_DIR");
}
