fn main() {
    let dir = std::env::var("BLAZE_B200_LIB_DIR").unwrap_or_else(|_| "../../blaze_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=blaze_b200");
    println!("cargo:rerun-if-env-changed=BLAZE_B200_LIB_DIR");
}
