{
  "targets": [
    {
      "target_name": "bls381_b200",
      "sources": ["addon.cc"],
      "include_dirs": ["../include"],
      "cflags_cc": ["-std=c++17", "-O2"],
      "libraries": ["-L<(module_root_dir)/../noble_bls12_381_b200", "-lbls381_b200", "-Wl,-rpath,<(module_root_dir)/../noble_bls12_381_b200"]
    }
  ]
}
