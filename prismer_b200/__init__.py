"""prismer_b200 -- B200-native (sm_100a) forward / backward / generate engine for the Prismer hot path.

Drop-in surface: ``PrismerCaption`` / ``PrismerVQA`` (``.forward`` / ``.generate`` / ``state_dict`` as the reference);
every block on the path is a hand-written CUDA kernel reached through the C-ABI in ``include/prismer_sm100.h``.
"""
__all__ = ["PrismerCaption", "PrismerVQA"]


def __getattr__(name):
    if name == "PrismerCaption":
        from .prismer_caption import PrismerCaption
        return PrismerCaption
    if name == "PrismerVQA":
        from .prismer_vqa import PrismerVQA
        return PrismerVQA
    raise AttributeError(name)
