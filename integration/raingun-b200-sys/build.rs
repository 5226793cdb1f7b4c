// Links libraingun_b200.so.  RAINGUN_B200_LIB_DIR points at the directory that holds it
// (raingun_b200/ in the raingun-b200 checkout, after `make -C raingun_b200/csrc`).
use std::env;

fn main() {
    if let Ok(dir) = env::var("RAINGUN_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=raingun_b200");
    println!("cargo:rerun-if-env-changed=RAINGUN_B200_LIB_DIR");
}
