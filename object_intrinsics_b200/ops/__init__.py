"""sm_100a rewrites of the StyleGAN2 ops vendored by the reference (src/third_party/{stylesdf/op,ada/torch_utils/ops})."""
from .fused_act import FusedLeakyReLU, fused_bias_act, fused_leaky_relu  # noqa: F401
from .bias_act import bias_act  # noqa: F401
from .upfirdn2d import (downsample2d, filter2d, setup_filter, upfirdn2d, upfirdn2d_native_layout,  # noqa: F401
                        upsample2d)
