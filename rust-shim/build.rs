// UNVERIFIED: never compiled.
fn main() {
    let dir = std::env::var("HMGPU_LIB_DIR").expect("set HMGPU_LIB_DIR to the directory that holds libhmgpu.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=hmgpu");
    println!("cargo:rerun-if-env-changed=HMGPU_LIB_DIR");
}
