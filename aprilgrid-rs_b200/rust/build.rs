// Link against the in-tree C-ABI library built by `make -C aprilgrid-rs_b200`.
fn main() {
    let dir = std::env::var("APRILGRID_B200_LIB_DIR").unwrap_or_else(|_| "../lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=aprilgrid_b200");
}
