// Links libzelll_b200.so.  ZELLL_B200_LIB_DIR names the directory that holds it
// (default: ../../zelll_b200, where `python -m zelll_b200.build` puts it).
fn main() {
    let dir = std::env::var("ZELLL_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../zelll_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zelll_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=ZELLL_B200_LIB_DIR");
}
