"""Model / task configuration objects for the VOLTA single-stream encoder step.

Mirrors the attribute surface of the reference's ``BertConfig`` / ``M3PConfig``
(reference: volta/volta/config.py:L218-413, L416-609): a JSON file whose keys
override constructor defaults (config.py:L388-394).  The encoder classes in
``clg_vqa_b200.encoders`` only read attributes, so an instance of the
reference's own config class can be passed instead of these.

Only the keys the hot path reads get defaults here; unknown JSON keys are kept
as attributes untouched (same as the reference's ``from_dict``).
"""
from __future__ import annotations

import copy
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG_DIR = os.path.join(_HERE, "configs")

_BERT_DEFAULTS = dict(
    vocab_size=30522, hidden_size=768, num_attention_heads=12, intermediate_size=3072,
    pooler_size=768, hidden_act="gelu", hidden_dropout_prob=0.1,
    attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
    pad_token_id=0, layer_norm_eps=1e-12, num_locs=5, image_embeddings="vilbert",
    initializer_range=0.02, v_feature_size=2048, v_hidden_size=768,
    v_num_attention_heads=12, v_intermediate_size=3072, v_hidden_dropout_prob=0.1,
    v_attention_probs_dropout_prob=0.1, fixed_layers=[], fusion_method="mul",
    fusion_act="relu", clf_hidden_size=1536, model="bert", norm_embeddings=False,
    tt_attn_sublayers=[], tv_attn_sublayers=[], vt_attn_sublayers=[], vv_attn_sublayers=[],
    t_ff_sublayers=[], v_ff_sublayers=[], shared_sublayers=[], single_ln_sublayers=[],
)

_M3P_DEFAULTS = dict(_BERT_DEFAULTS)
_M3P_DEFAULTS.update(n_heads=12, n_layers=12, image_embeddings="m3p", model="roberta", n_words=250002, emb_dim=768,
                     n_langs=100, refine_layers=6, pad_index=1, dropout=0.1, attention_dropout=0.1, vocab_size=250002,
                     pad_token_id=1, type_vocab_size=1, fusion_method="text", clf_hidden_size=1536, num_locs=5,
                     norm_embeddings=True)


class _Config:
    _defaults: dict = {}

    def __init__(self, **kwargs):
        for k, v in self._defaults.items():
            setattr(self, k, copy.deepcopy(v))
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_dict(cls, json_object):
        return cls(**json_object)

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.loads(reader.read()))

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def __repr__(self):
        return "%s %s" % (type(self).__name__, self.to_json_string())


class BertConfig(_Config):
    """UC2 / BERT-family config (reference: volta/volta/config.py:L218-413)."""
    _defaults = _BERT_DEFAULTS


class M3PConfig(_Config):
    """M3P config (reference: volta/volta/config.py:L416-609)."""
    _defaults = _M3P_DEFAULTS


def uc2_base(num_layers: int = 12, **overrides) -> BertConfig:
    """The shipped UC2-base config (values of volta/config/uc2_base.json); `num_layers`
    shrinks the sub-layer lists for fast parity cases (sub-layer 2l = attention,
    2l+1 = feed-forward, all shared / single-LN, as in uc2_base.json:L33-168)."""
    cfg = BertConfig.from_json_file(os.path.join(CONFIG_DIR, "uc2_base.json"))
    attn = list(range(0, 2 * num_layers, 2))
    ff = list(range(1, 2 * num_layers, 2))
    for k in ("tt_attn_sublayers", "tv_attn_sublayers", "vt_attn_sublayers", "vv_attn_sublayers"):
        setattr(cfg, k, list(attn))
    for k in ("t_ff_sublayers", "v_ff_sublayers"):
        setattr(cfg, k, list(ff))
    cfg.shared_sublayers = list(range(2 * num_layers))
    cfg.single_ln_sublayers = list(range(2 * num_layers))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def m3p_base(num_layers: int = 12, **overrides) -> M3PConfig:
    cfg = M3PConfig.from_json_file(os.path.join(CONFIG_DIR, "m3p_base.json"))
    cfg.n_layers = num_layers
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def gqa_task_cfg(max_region_num: int = 36, num_labels: int = 1842, semantic_lambda: float = 10.0,
                 task: str = "TASK15") -> dict:
    """The fields of the GQA task yaml the hot path reads
    (volta/config_tasks/iglue_trainval_tasks_boxes36.dtu.yml:L1-36)."""
    return {task: {
        "name": "GQA", "type": "VL-classifier-GQA", "num_labels": num_labels,
        "loss": "CrossEntropyLoss", "process": "normal", "task_id": 15,
        "max_seq_length": 40, "max_region_num": max_region_num, "batch_size": 256,
        "eval_batch_size": 1024, "lr": 4e-5, "num_epoch": 5,
        "semantic_lambda": semantic_lambda,
    }}


def num_encoder_layers(config) -> int:
    if hasattr(config, "tt_attn_sublayers") and config.tt_attn_sublayers:
        return len(config.tt_attn_sublayers)
    return int(getattr(config, "n_layers", 12))
