{
  # node-gyp recipe for the addon (unexecuted here: no Node in the authoring image; see INTEGRATION.md).
  # libbpe_b200.so is built by `python -m bpe_tokenizer_b200.build` (nvcc, sm_100a) in ../../bpe_tokenizer_b200.
  "targets": [
    {
      "target_name": "bpe_b200",
      "sources": ["bpe_b200_napi.c"],
      "include_dirs": ["../../include"],
      "cflags": ["-std=c11", "-Wall", "-Wextra"],
      "libraries": [
        "-L<(module_root_dir)/../../bpe_tokenizer_b200",
        "-lbpe_b200",
        "-Wl,-rpath,<(module_root_dir)/../../bpe_tokenizer_b200"
      ]
    }
  ]
}
