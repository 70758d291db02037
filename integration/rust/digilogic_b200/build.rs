// Links libdigilogic_b200.so (built by `python digilogic_b200/build.py` in the engine repository).
// DIGILOGIC_B200_LIB_DIR = directory holding the library.
fn main() {
    let dir = std::env::var("DIGILOGIC_B200_LIB_DIR").expect("set DIGILOGIC_B200_LIB_DIR to the directory of libdigilogic_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=digilogic_b200");
    println!("cargo:rerun-if-env-changed=DIGILOGIC_B200_LIB_DIR");
}
