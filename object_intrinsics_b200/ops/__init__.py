"""sm_100a rewrites of the StyleGAN2 ops vendored by the reference.

Module layout mirrors the reference so that call sites read the same:
  ops.upfirdn2d  <- src/third_party/ada/torch_utils/ops/upfirdn2d.py   (upfirdn2d, upsample2d, downsample2d, ...)
                    + upfirdn2d_native_layout for src/third_party/stylesdf/op/upfirdn2d.py
  ops.bias_act   <- src/third_party/ada/torch_utils/ops/bias_act.py    (bias_act)
  ops.fused_act  <- src/third_party/stylesdf/op/fused_act.py           (fused_leaky_relu, FusedLeakyReLU)
"""
from . import bias_act, fused_act, upfirdn2d  # noqa: F401
