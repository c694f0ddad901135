// Locates libportello_b200.so (built in-tree by `make -C portello_b200/csrc`, nvcc for sm_100a).
// PORTELLO_B200_LIB_DIR overrides the default location relative to this crate.
fn main() {
    let dir = std::env::var("PORTELLO_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../portello_b200/csrc", here)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=portello_b200");
    println!("cargo:rerun-if-env-changed=PORTELLO_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/portello_b200.h");
}
