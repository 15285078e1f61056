"""Drop-in for the reference's `modelZoo.py`: same class names and signatures, B200 (sm_100a) kernels underneath.

    import modelZoo
    generator = getattr(modelZoo, "regressor_fcn_bn_32")()        # train_gan.py:61-67
    generator.build_net(36, 252, require_text=False)
    generator.to("cuda")
    out = generator(input_, feats_=None)
"""
import b2h_b200  # noqa: F401
from b2h_b200.modelzoo import (regressor_fcn_bn_32, regressor_fcn_bn_32_b2h, regressor_fcn_bn_32_v2,  # noqa: F401
                               regressor_fcn_bn_32_v4, regressor_fcn_bn_32_v4_deeper,
                               regressor_fcn_bn_discriminator)

__all__ = ["regressor_fcn_bn_32", "regressor_fcn_bn_32_b2h", "regressor_fcn_bn_32_v2", "regressor_fcn_bn_32_v4",
           "regressor_fcn_bn_32_v4_deeper", "regressor_fcn_bn_discriminator"]
