// Links libaero_b200.so.  AERO_B200_LIB_DIR points at the directory holding the library built by
// `python -m aero_b200.build` (aero_b200/libaero_b200.so in the aero_b200 repository).
fn main() {
    let dir = std::env::var("AERO_B200_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=aero_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=AERO_B200_LIB_DIR");
}
