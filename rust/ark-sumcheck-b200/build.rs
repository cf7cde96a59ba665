// Source only — not compiled in this image.  Point SUMCHECK_B200_LIB_DIR at the directory holding libsumcheck_b200.so.
fn main() {
    let dir = std::env::var("SUMCHECK_B200_LIB_DIR").unwrap_or_else(|_| "../../sumcheck_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sumcheck_b200");
    println!("cargo:rerun-if-env-changed=SUMCHECK_B200_LIB_DIR");
}
