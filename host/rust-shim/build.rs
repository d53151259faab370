// Links libmapquik_b200.so; MAPQUIK_B200_LIB_DIR points at the directory that holds it (mapquik_b200/ in this repo).
fn main() {
    let dir = std::env::var("MAPQUIK_B200_LIB_DIR").expect("set MAPQUIK_B200_LIB_DIR to the directory of libmapquik_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=mapquik_b200");
    println!("cargo:rerun-if-env-changed=MAPQUIK_B200_LIB_DIR");
}
