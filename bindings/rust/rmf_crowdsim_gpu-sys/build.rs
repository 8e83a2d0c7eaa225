fn main() {
    // librcs.so is built by `python -m rmf_crowdsim_b200._build` (nvcc, sm_100a)
    let dir = std::env::var("RCS_LIB_DIR").expect("set RCS_LIB_DIR to the directory that holds librcs.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rcs");
    println!("cargo:rerun-if-env-changed=RCS_LIB_DIR");
}
