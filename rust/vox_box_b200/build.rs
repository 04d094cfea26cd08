// Links the in-tree shared library built by `make -C vox_box.rs_b200` (or __graft_entry__.build()).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("VOXBOX_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../vox_box.rs_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=voxbox_b200");
    println!("cargo:rerun-if-env-changed=VOXBOX_B200_LIB_DIR");
}
