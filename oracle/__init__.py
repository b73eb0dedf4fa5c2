"""CPU oracle for the RaCFormer sampling hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package. racformer_b200/ never does (tests/test_no_oracle_in_product.py enforces it).

  c_oracle        exact-arithmetic C restatement of the CUDA kernels (masks bit-exact, fp64 accumulation)
  reference_port  restatement of the reference's own PyTorch grid_sample paths (the CPU baseline)
  build_ref       recipe that compiles the reference's unmodified CUDA extension into oracle/_ref/
"""
