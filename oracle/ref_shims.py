"""Import shims that let the UNMODIFIED reference run on CPU in the build container.

Used only by ``oracle/make_golden.py`` (fixture generation) -- ``/root/reference`` does
not exist on the GPU box, so nothing at test/bench time imports this module.

The reference needs: ipdb (model/transformer.py:28), easydict (model/bert.py:14),
timm.models.layers (model/mico.py:19, model/swin.py:11) and a transformers 4.31 API
surface (model/bert.py:40-60).  We stub exactly those names; no reference source is
copied.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MICO_REFERENCE_ROOT", "/root/reference")


class _AttrDict(dict):
    """Minimal stand-in for easydict.EasyDict (attribute access over a dict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_AttrDict(x) if isinstance(x, dict) else x for x in v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def install():
    """Install stubs into sys.modules and chdir to the reference root."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)  # model/mico.py:102,109 use relative paths

    sys.modules.setdefault("ipdb", types.ModuleType("ipdb"))
    ed = types.ModuleType("easydict")
    ed.EasyDict = _AttrDict
    sys.modules.setdefault("easydict", ed)

    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for name in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def _no_prune(*a, **k):
            raise NotImplementedError("head pruning is not on the path")
        mu.find_pruneable_heads_and_indices = _no_prune
    import transformers.models.auto as auto
    if not hasattr(auto, "MODEL_FOR_VISION_2_SEQ_MAPPING"):
        auto.MODEL_FOR_VISION_2_SEQ_MAPPING = {}
    for modname, names in (
        ("transformers.generation.beam_constraints", ("DisjunctiveConstraint", "PhrasalConstraint", "Constraint")),
        ("transformers.generation.beam_search", ("BeamScorer", "BeamSearchScorer", "ConstrainedBeamSearchScorer")),
    ):
        try:
            __import__(modname)
        except Exception:
            m = types.ModuleType(modname)
            for n in names:
                setattr(m, n, type(n, (), {}))
            sys.modules[modname] = m
    mu.PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n

    # timm.models.layers: re-export the identical helpers the reference vendors itself
    if "timm" not in sys.modules:
        try:
            import timm.models.layers  # noqa: F401
        except Exception:
            import importlib
            eva = importlib.import_module("model.evaclip.eva_vit_model")
            timm = types.ModuleType("timm")
            models = types.ModuleType("timm.models")
            layers = types.ModuleType("timm.models.layers")
            layers.DropPath = eva.DropPath
            layers.to_2tuple = eva.to_2tuple
            layers.trunc_normal_ = eva.trunc_normal_
            timm.models = models
            models.layers = layers
            sys.modules["timm"] = timm
            sys.modules["timm.models"] = models
            sys.modules["timm.models.layers"] = layers
    return _AttrDict


def default_model_cfg(**over):
    cfg = dict(vision_encoder_type="evaclip01_giant", vision_resolution=224, checkpointing=False,
               contra_dim=512, max_vision_sample_num=8, max_audio_sample_num=3, max_depth_sample_num=1,
               beam_size=3, itm_ratio=0.1, max_omni_caption_len=70, max_caption_len=40,
               max_subtitle_len=70, frame_embedding_type="adaptive", pool_video=False)
    cfg.update(over)
    return _AttrDict(cfg)
