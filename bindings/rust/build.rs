// build.rs -- only acts under the `cuda` feature: link libndconv_cuda.so from NDCONV_LIB_DIR
// (the directory that holds the library built by `python __graft_entry__.py`, i.e. <repo>/ndarray-conv_b200).
fn main() {
    println!("cargo:rerun-if-env-changed=NDCONV_LIB_DIR");
    if std::env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    let dir = std::env::var("NDCONV_LIB_DIR")
        .expect("feature `cuda`: set NDCONV_LIB_DIR to the directory that contains libndconv_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ndconv_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
