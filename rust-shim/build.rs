// Links the prebuilt libcqb200.so (built by `make -C sha2-on-cq-halo2_b200/csrc`); set CQB200_LIB_DIR to its directory.
fn main() {
    let dir = std::env::var("CQB200_LIB_DIR").unwrap_or_else(|_| "../sha2-on-cq-halo2_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=cqb200");
    println!("cargo:rerun-if-env-changed=CQB200_LIB_DIR");
}
