"""Model registry with timm's ``register_model`` / ``create_model`` call shape.

The reference registers its models with timm 0.4.12 (mem/modeling_pretrain.py:15,128,
mem/modeling_finetune.py:18,379) and builds them with ``timm.models.create_model(name, **kw)``
(mem/run_mem_pretraining.py:176-192).  timm is not a dependency of this package: the same two
functions live here, and when a real ``timm`` is importable the models are registered there as
well so that ``timm.create_model('pt_vit', ...)`` keeps working.
"""
from __future__ import annotations

_MODELS = {}


def register_model(fn):
    _MODELS[fn.__name__] = fn
    try:  # optional: mirror into timm's registry when timm exists
        from timm.models.registry import register_model as _timm_register
        _timm_register(fn)
    except Exception:
        pass
    return fn


def create_model(model_name, pretrained=False, **kwargs):
    """``timm.models.create_model`` subset: look the entry point up by name and call it.
    ``drop_block_rate`` (always None in the reference's call) is accepted and dropped."""
    if model_name not in _MODELS:
        raise RuntimeError(f"Unknown model ({model_name}); registered: {sorted(_MODELS)}")
    kwargs = dict(kwargs)
    if kwargs.get("drop_block_rate", None) is None:
        kwargs.pop("drop_block_rate", None)
    return _MODELS[model_name](pretrained=pretrained, **kwargs)


def list_models():
    return sorted(_MODELS)
