// Link against libb2z.so built in-tree by `python -m zeth_b200.build` (B2Z_LIB_DIR overrides the search path).
fn main() {
    let dir = std::env::var("B2Z_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        manifest.join("../../../zeth_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b2z");
    println!("cargo:rerun-if-env-changed=B2Z_LIB_DIR");
}
