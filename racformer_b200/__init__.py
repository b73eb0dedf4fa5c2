"""racformer_b200 -- B200-native (sm_100a) implementation of RaCFormer's sampling hot path.

  wrapper                               msmv_sampling & friends (mirror of the reference's models/csrc/wrapper.py)
  multi_scale_deformable_attn_function  MultiScaleDeformableAttnFunction_fp32/_fp16, ext_module (mirror of the reference's
                                        models/multi_scale_deformable_attn_function.py)
  bev_pool                              bev_pool_v2, QuickCumsumCuda (mirror of models/csrc/bev_pool_v2/bev_pool.py)
  points                                fused sampling-point generation / re-layout kernels, self-attention core, box
                                        refinement, AdaptiveMixing core (inference)
  linear                                tcgen05 Linear layers with exact bf16x3 operand splitting (inference and, through
                                        TrainableSplitLinear, training)
  rowops                                row programs: the row-wise operator chains of a decoder iteration in one launch
  decoder, graphs, parallel, synthetic  decoder harness, CUDA-graph serving, multi-GPU plumbing, synthetic inputs
  build, _lib                           nvcc build of libracformer_ops.so and its ctypes binding (include/racformer_ops.h)

Importing a sub-module that binds kernels loads libracformer_ops.so and fails loudly if it is missing; there is no
CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"
