fn main() {
    // OFPS_B200_LIB_DIR = directory holding libofps_b200.so (python -m ofps_b200.build puts it in ofps_b200/)
    if let Ok(dir) = std::env::var("OFPS_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=ofps_b200");
    println!("cargo:rerun-if-env-changed=OFPS_B200_LIB_DIR");
}
