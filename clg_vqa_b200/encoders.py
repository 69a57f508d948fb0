"""Drop-in module surface of the reference's `volta/volta/encoders.py` for the fine-tune / eval hot
path: `BertForVLTasks` (UC2) and `M3PForVLTasks` (M3P).

Same constructor, `forward(input_txt, input_imgs, image_loc, task_id, token_type_ids, attention_mask,
image_attention_mask, ...)` signature and return tuple (encoders.py:L1202-1259, L1311-1353), same
`state_dict()` keys / shapes incl. the aliased `v_*` entries, same
`named_parameters()` names (they drive the weight-decay groups of train_task.py:L249-260) and the
same named sub-modules `torch.nn.utils.prune` re-parametrises in train_task_sft.py:L44-132.

The sub-modules only HOLD parameters: their own `forward`s are never called.  The whole step runs in
`clg_vqa_b200.engine.Engine`, which executes the plain post-LN BERT restatement of the gated
bi-stream encoder with hand-written sm_100a kernels through the C ABI.  There is no
PyTorch-eager or CPU fallback: calling forward without the CUDA library / off a B200 raises.
"""
from __future__ import annotations

import logging
import os

import torch
import torch.nn as nn

from . import config as cfgmod

logger = logging.getLogger(__name__)


class BertLayerNorm(nn.Module):
    """Parameter holder with the reference's names (encoders.py:L44-62; apex FusedLayerNorm or the
    python fallback there).  eps is read by the engine."""

    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps
        self.eps = eps


class GeLU(nn.Module):
    pass


def _no_call(self, *a, **k):  # pragma: no cover
    raise RuntimeError("parameter-holder sub-modules are not callable; the step runs in clg_vqa_b200.engine")


# ------------------------------------------------------------------------------------------------
# UC2
# ------------------------------------------------------------------------------------------------
class UC2Embeddings(nn.Module):
    """embeddings.py:L605-669"""

    def __init__(self, config):
        super().__init__()
        self.padding_idx = config.pad_token_id
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=self.padding_idx)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.new_token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.image_embeddings = nn.Linear(config.v_feature_size, config.v_hidden_size)
        self.image_location_embeddings = nn.Linear(config.num_locs, config.v_hidden_size)
        self.image_token_type_embeddings = self.new_token_type_embeddings
        self.image_layer_norm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.image_location_layer_norm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.v_LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.v_dropout = nn.Dropout(config.hidden_dropout_prob)


class BertGatedSelfAttention(nn.Module):
    """encoders.py:L164-359 (shared text/vision weights: v_* are the same module objects)."""

    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(config.hidden_size, config.hidden_size)
        self.value = nn.Linear(config.hidden_size, config.hidden_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self.v_query = self.query
        self.v_key = self.key
        self.v_value = self.value
        self.v_dropout = self.dropout


class BertGatedSelfOutput(nn.Module):
    """encoders.py:L362-425"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.v_dense = self.dense
        self.v_LayerNorm = self.LayerNorm
        self.v_dropout = self.dropout


class BertGatedAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention_self = BertGatedSelfAttention(config)
        self.attention_output = BertGatedSelfOutput(config)


class BertGatedIntermediate(nn.Module):
    """encoders.py:L453-502"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.v_dense = self.dense


class BertGatedOutput(nn.Module):
    """encoders.py:L505-567"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.v_dense = self.dense
        self.v_LayerNorm = self.LayerNorm
        self.v_dropout = self.dropout


class BertGatedFeedForward(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.intermediate = BertGatedIntermediate(config)
        self.output = BertGatedOutput(config)


class BertEncoder(nn.Module):
    """encoders.py:L821-892: sub-layer 2l = attention, 2l+1 = feed-forward."""

    def __init__(self, config):
        super().__init__()
        _check_uc2_structure(config)
        n = cfgmod.num_encoder_layers(config)
        layers = []
        for _ in range(n):
            layers.append(BertGatedAttention(config))
            layers.append(BertGatedFeedForward(config))
        self.layer = nn.ModuleList(layers)


class BertTextPooler(nn.Module):
    """encoders.py:L597-608"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.pooler_size)
        self.activation = nn.ReLU() if getattr(config, "fusion_act", "relu") == "relu" else nn.Tanh()


class SimpleClassifier(nn.Module):
    """encoders.py:L788-815: Linear -> GeLU -> LayerNorm -> Linear"""

    def __init__(self, in_dim, hid_dim, out_dim, layer_norm_eps=1e-12, dropout_prob=0.0):
        super().__init__()
        self.logit_fc = nn.Sequential(
            nn.Linear(in_dim, hid_dim), GeLU(), BertLayerNorm(hid_dim, eps=layer_norm_eps), nn.Linear(hid_dim, out_dim))


class BertModel(nn.Module):
    """encoders.py:L922-1021 with the shared UC2 embeddings (L954-956)."""

    def __init__(self, config):
        super().__init__()
        self.embeddings = UC2Embeddings(config)
        self.encoder = BertEncoder(config)
        self.t_pooler = BertTextPooler(config)
        self.fusion_method = config.fusion_method
        self.shared_embeddings = True


def _check_uc2_structure(config):
    """The engine implements the configuration the reference ships for UC2 (uc2_base.json): every
    sub-layer shared and single-LN, tt/tv/vt/vv attention on all attention sub-layers, text fusion.
    Anything else is a different VOLTA model and is refused rather than silently mis-computed."""
    n = cfgmod.num_encoder_layers(config)
    attn = list(range(0, 2 * n, 2))
    ff = list(range(1, 2 * n, 2))
    ok = (getattr(config, "image_embeddings", None) == "uc2" and
          all(sorted(getattr(config, k)) == attn for k in
              ("tt_attn_sublayers", "tv_attn_sublayers", "vt_attn_sublayers", "vv_attn_sublayers")) and
          sorted(config.t_ff_sublayers) == ff and sorted(config.v_ff_sublayers) == ff and
          sorted(config.shared_sublayers) == list(range(2 * n)) and
          sorted(config.single_ln_sublayers) == list(range(2 * n)) and
          config.hidden_size == config.v_hidden_size and
          config.hidden_size == 64 * config.num_attention_heads and
          getattr(config, "fusion_method", "text") == "text")
    if not ok:
        raise ValueError("clg_vqa_b200 implements the single-stream UC2 configuration (uc2_base.json layout: all "
                         "sub-layers shared + single-LN, tt/tv/vt/vv attention, 64-d heads, text fusion)")


class _VLTasksBase(nn.Module):
    base_model_prefix = "bert"
    _family = "uc2"

    def init_weights(self, module):
        """encoders.py:L908-919"""
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, (BertLayerNorm, nn.LayerNorm)):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def _make_classifiers(self, config, task_cfg, task_ids):
        task2clf = {}
        for task_id in task_ids:
            task_type = task_cfg[task_id]["type"]
            if task_type in {"VL-classifier", "VL-classifier-GQA"}:
                task2clf[task_id] = SimpleClassifier(config.pooler_size, config.clf_hidden_size,
                                                     task_cfg[task_id]["num_labels"], config.layer_norm_eps)
            elif task_type in {"VL-binary-classifier", "VL-tri-classifier", "VL-logit"} or task_type.startswith("V-logit"):
                raise ValueError("task type %s is outside the GQA hot path this package implements" % task_type)
            else:
                raise ValueError("Undefined task type: %s" % task_type)   # encoders.py:L1188
        return nn.ModuleDict(task2clf)

    # -- engine plumbing --------------------------------------------------------------------------
    def _get_engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None:
            from .engine import Engine
            eng = Engine(self, family=self._family)
            self.__dict__["_engine"] = eng
        return eng

    @property
    def engine(self):
        return self._get_engine()

    def zero_grad(self, set_to_none: bool = True):
        """nn.Module.zero_grad; after the fused optimiser step the gradient arena is already all zero (the AdamW kernel
        cleared what it read) and the next backward re-attaches it - walking 215 parameters to drop their `.grad` views
        would only cost host time (1 ms per step at small batches)."""
        eng = self.__dict__.get("_engine")
        if eng is not None and eng.grads_consumed and eng.grads_zeroed:
            return
        super().zero_grad(set_to_none=set_to_none)

    def set_precision(self, mode: str):
        """'bf16' (tcgen05 GEMMs, bf16 operands, fp32 accumulate / residual stream) or 'fp32'
        (correctness mode: fp32 FFMA GEMMs)."""
        if mode not in ("bf16", "fp32"):
            raise ValueError(mode)
        self.__dict__["_precision"] = mode
        eng = self.__dict__.get("_engine")
        if eng is not None:
            eng.set_precision(mode)

    @property
    def precision(self):
        return self.__dict__.get("_precision", "bf16")

    # -- checkpoint loading (utils.py:L361-580) ---------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        config = kwargs.pop("config", None)
        state_dict = kwargs.pop("state_dict", None)
        kwargs.pop("cache_dir", None)
        kwargs.pop("from_tf", None)
        kwargs.pop("default_gpu", None)
        kwargs.pop("from_hf", None)
        output_loading_info = kwargs.pop("output_loading_info", False)
        if config is None:
            raise ValueError("from_pretrained needs config= (the reference scripts always pass it)")
        model = cls(config, *model_args, **kwargs)
        if state_dict is None:
            path = pretrained_model_name_or_path
            if os.path.isdir(path):
                path = os.path.join(path, "pytorch_model.bin")
            state_dict = torch.load(path, map_location="cpu")
        state_dict = dict(state_dict)
        for key in list(state_dict.keys()):            # old gamma/beta names
            new_key = None
            if "gamma" in key:
                new_key = key.replace("gamma", "weight")
            if "beta" in key:
                new_key = key.replace("beta", "bias")
            if new_key:
                state_dict[new_key] = state_dict.pop(key)
        if getattr(config, "model", "bert") == "roberta":
            for k in list(state_dict.keys()):
                nk = k.replace("roberta", "bert")
                if nk != k:
                    state_dict[nk] = state_dict.pop(k)
        model_to_load, prefix = model, ""
        has_prefix = any(s.startswith(cls.base_model_prefix + ".") for s in state_dict.keys())
        if not has_prefix:
            model_to_load = getattr(model, cls.base_model_prefix)
        res = model_to_load.load_state_dict(state_dict, strict=False)
        missing, unexpected = list(res.missing_keys), list(res.unexpected_keys)
        if missing:
            logger.info("Weights of %s not initialized from pretrained model: %s", cls.__name__, missing)
        if unexpected:
            logger.info("Weights from pretrained model not used in %s: %s", cls.__name__, unexpected)
        model.eval()
        if output_loading_info:
            return model, {"missing_keys": missing, "unexpected_keys": unexpected, "error_msgs": []}
        return model


class BertForVLTasks(_VLTasksBase):
    """encoders.py:L1154-1259"""
    _family = "uc2"

    def __init__(self, config, task_cfg, task_ids, dropout_prob=0.1):
        super().__init__()
        self.config = config
        self.bert = BertModel(config)
        self.dropout = nn.Dropout(dropout_prob)
        self.task_cfg = task_cfg
        self.clfs_dict = self._make_classifiers(config, task_cfg, task_ids)
        self.fusion_method = config.fusion_method
        self.apply(self.init_weights)

    def forward(self, input_txt, input_imgs, image_loc, task_id, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        if output_all_encoded_layers or output_all_attention_masks:
            raise NotImplementedError("per-layer outputs are not materialised by the fused engine")
        if task_id not in self.clfs_dict:
            raise KeyError(task_id)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_txt)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_txt)
        if image_attention_mask is None:
            image_attention_mask = torch.ones(input_imgs.size(0), input_imgs.size(1)).type_as(input_txt)
        logits = self._get_engine().run(input_txt, input_imgs, image_loc, task_id, token_type_ids, attention_mask,
                                        image_attention_mask)
        return logits, None, None, ([None], [None])


# ------------------------------------------------------------------------------------------------
# M3P
# ------------------------------------------------------------------------------------------------
class M3PModel(nn.Module):
    """encoders.py:L1024-1042: `encoder` = M3PTransformerModel (only jointfwd is used), `pooler` is the
    SAME module object as encoder.pooled_layer (state_dict alias bert.pooler.* == bert.encoder.pooled_layer.*)."""

    def __init__(self, config):
        super().__init__()
        from . import m3p_layout
        self.encoder = m3p_layout.build_encoder(config)
        self.pooler = self.encoder.pooled_layer


class M3PForVLTasks(_VLTasksBase):
    """encoders.py:L1262-1353"""
    _family = "m3p"

    def __init__(self, config, task_cfg, task_ids, dropout_prob=0.1):
        super().__init__()
        self.config = config
        if int(getattr(config, "emb_dim", 768)) != 64 * int(getattr(config, "n_heads", 12)):
            raise ValueError("clg_vqa_b200 implements 64-d attention heads (m3p_base.json: 768 / 12)")
        self.bert = M3PModel(config)
        self.dropout = nn.Dropout(dropout_prob)
        self.task_cfg = task_cfg
        self.clfs_dict = self._make_classifiers(config, task_cfg, task_ids)
        self.fusion_method = getattr(config, "fusion_method", "text")
        self.apply(self.init_weights)

    def forward(self, input_txt, input_imgs, image_loc, task_id, token_type_ids=None, attention_mask=None,
                image_attention_mask=None, output_all_encoded_layers=False, output_all_attention_masks=False):
        if output_all_encoded_layers:
            raise NotImplementedError("output_all_encoded_layers for M3P")      # encoders.py:L1351
        if task_id not in self.clfs_dict:
            raise KeyError(task_id)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_txt)
        if image_attention_mask is None:
            image_attention_mask = torch.ones(input_imgs.size(0), input_imgs.size(1)).type_as(input_txt)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_txt)      # ignored by M3P (m3p_transformer.py:L877-964)
        logits = self._get_engine().run(input_txt, input_imgs, image_loc, task_id, token_type_ids, attention_mask,
                                        image_attention_mask)
        return logits, None, None, None
