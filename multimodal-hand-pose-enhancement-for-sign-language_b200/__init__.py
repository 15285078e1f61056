"""b2h_b200 -- B200-native (sm_100a) implementation of the Body2Hands-style temporal pose regressor /
discriminator GAN hot path of alvaro-budria/Multimodal-Hand-Pose-Enhancement-for-Sign-Language.

Everything numerical runs in libb2h.so (hand-written CUDA behind the C ABI of include/b2h_abi.h);
this package is the host-side mirror of the reference's modelZoo / train_gan interfaces.
There is no CPU or PyTorch fallback: without the library or a B200-class GPU, calls raise.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
