"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference modules.

Only usable in the build container (``/root/reference`` is absent on the GPU box).
It is used by ``oracle/gen_golden.py`` to generate the committed fixtures under
``tests/golden/`` and by the ``-m "not gpu"`` tests (skipped when the reference
tree is not mounted) to pin ``oracle/prismer_oracle.py`` against the real thing.

Shims (SURVEY.md section 8c -- version drift vs the pinned transformers 4.26.1):
  1. ``clip.clip._download`` is stubbed: ``model/modules/vit.py:10`` imports it, but
     it is only used by the network loader ``vit.py:179``.
  2. the LM-head weight tie (``roberta.py:352-356``) is forced explicitly, because
     transformers 5.x does not tie through ``get_output_embeddings`` after ``post_init``.
  3. ``GenerationMixin`` is mixed in (``PreTrainedModel`` dropped it in >= 4.50) so that
     ``text_decoder.generate`` (``prismer_caption.py:45``) exists.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PRISMER_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model", "modules"))


def _install_stubs():
    if "clip.clip" not in sys.modules:
        clip = types.ModuleType("clip")
        clip_clip = types.ModuleType("clip.clip")

        def _download(*a, **k):  # pragma: no cover - network loader is never used
            raise RuntimeError("network download is not available in the oracle")

        clip_clip._download = _download
        clip.clip = clip_clip
        sys.modules["clip"] = clip
        sys.modules["clip.clip"] = clip_clip
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Returns a namespace with the reference's hot-path classes."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    import importlib
    vit = importlib.import_module("model.modules.vit")
    resampler = importlib.import_module("model.modules.resampler")
    roberta = importlib.import_module("model.modules.roberta")
    utils = importlib.import_module("model.modules.utils")
    from transformers import RobertaConfig
    try:
        from transformers.generation import GenerationMixin
    except Exception:  # pragma: no cover
        from transformers import GenerationMixin

    class Decoder(roberta.RobertaForCausalLMModified, GenerationMixin):
        pass

    def build_decoder(cfg_dict):
        cfg = RobertaConfig.from_dict(dict(cfg_dict))
        dec = Decoder(cfg)
        # shim (2): explicit tie, roberta.py:352-356 + :417-419
        dec.lm_head.decoder.weight = dec.roberta.embeddings.word_embeddings.weight
        dec.lm_head.decoder.bias = dec.lm_head.bias
        return dec

    ns = types.SimpleNamespace(vit=vit, resampler=resampler, roberta=roberta, utils=utils,
                               VisionTransformer=vit.VisionTransformer, build_decoder=build_decoder,
                               RobertaConfig=RobertaConfig)
    return ns
