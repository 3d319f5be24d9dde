// build.rs of backends/b200: link librfwb200.so.  Source only, never compiled here.
fn main() {
    // librfwb200.so is built by `make -C rfw_rs_b200/csrc`; point RFWB200_LIB_DIR at its directory
    let dir = std::env::var("RFWB200_LIB_DIR").expect("set RFWB200_LIB_DIR");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rfwb200");
}
