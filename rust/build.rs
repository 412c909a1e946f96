// NOT COMPILED in this repository's environment (no rustc/cargo in the image).
// Reference-side binding of libgenedex_b200.so, kept in sync with INTEGRATION.md and include/genedex_b200.h.
fn main() {
    // path of genedex_b200/csrc (contains libgenedex_b200.so, built with `make -C genedex_b200/csrc`)
    let dir = std::env::var("GENEDEX_B200_LIB_DIR").expect("set GENEDEX_B200_LIB_DIR");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=genedex_b200");
}
