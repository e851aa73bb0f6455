"""Drop-in mirror of the reference's model API for the MM-RCA late-fusion path.

Same class names, constructor arguments, forward(_input_ids, _attention_mask, _images, eval,
remove_image, remove_text) signature, helper methods and .pth state_dict layout as
CVPR_code/multimodal_model.py of espiriki/Garbage_Classification_RCA — but everything after the
backbones (reference :661-728) runs in the sm_100a kernels behind libmmrca.so.  The image / text
backbones stay stock torchvision / HuggingFace modules (north_star).

Deviations, all documented in SURVEY.md §0:
  * trailing ctor args are defaulted (batch_size=16, reverse=False, features_only=False,
    cross_attention_only=False) so that the reference's own 8/9-argument call sites work;
  * `pretrained=` / `compute=` / `dropout_generator=` keyword-only extensions;
  * features_only skips the attention blocks the reference computes and discards (:676-699).
"""
from __future__ import annotations

import math
import sys
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import _native as N
from . import functional as F

_TEXT_HIDDEN = {"distilbert": 768, "bert": 768, "bart": 1024}
IMAGE_FEATURES = 1280   # reference :258 (EfficientNetV2 pooled width)


# ---------------------------------------------------------------------------------------------
# backbones (stock modules; reference :11-36, :113-153)
# ---------------------------------------------------------------------------------------------
class EfficientNetV2MFullFeatureExtractor(nn.Module):
    """Re-registers slices of torchvision's EfficientNetV2 `features` under the names the reference
    uses (stem, stage1..stage6, final_conv — reference :14-23) so checkpoint keys interchange, and
    returns (stage-3 map, stage-6 map, pooled vector) like reference :25-36."""

    STAGES = ("stage1", "stage2", "stage3", "stage4", "stage5", "stage6")

    def __init__(self, model: nn.Module):
        super().__init__()
        feats = model.features
        self.stem = feats[:2]
        for i, name in enumerate(self.STAGES):
            setattr(self, name, feats[2 + i])
        self.final_conv = feats[8]
        self.avgpool = model.avgpool
        self.classifier = model.classifier   # registered but never applied (reference :23)

    def forward_maps(self, x):
        """(stage-3 map, stage-6 map, final_conv map): forward() without the avgpool + flatten, for the fused feature
        hand-off (functional.feature_handoff pools, gathers the CLS row and casts in one kernel)."""
        x = self.stem(x)
        taps = {}
        for name in self.STAGES:
            x = getattr(self, name)(x)
            taps[name] = x
        return taps["stage3"], taps["stage6"], self.final_conv(x)

    def forward(self, x):
        s3, s6, fmap = self.forward_maps(x)
        return s3, s6, torch.flatten(self.avgpool(fmap), 1)


def _freeze(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        p.requires_grad = False
    return m


def eff_net_v2(pretrained: bool = True) -> nn.Module:
    """EfficientNetV2-M feature extractor, frozen, classifier trimmed to its Dropout (reference :113-126)."""
    from torchvision.models import efficientnet_v2_m
    model = _freeze(efficientnet_v2_m(weights="IMAGENET1K_V1" if pretrained else None))
    model.classifier = nn.Sequential(model.classifier[0])
    return EfficientNetV2MFullFeatureExtractor(model)


def distilbert(pretrained: bool = True) -> nn.Module:
    from transformers import DistilBertConfig, DistilBertModel
    m = DistilBertModel.from_pretrained("distilbert-base-uncased") if pretrained else DistilBertModel(DistilBertConfig())
    return _freeze(m)


def bert(pretrained: bool = True) -> nn.Module:
    from transformers import BertConfig, BertModel
    m = BertModel.from_pretrained("bert-base-uncased") if pretrained else BertModel(BertConfig())
    return _freeze(m)


def bart(pretrained: bool = True) -> nn.Module:
    from transformers import BartConfig, BartModel
    m = BartModel.from_pretrained("facebook/bart-large") if pretrained else BartModel(BartConfig())
    return _freeze(m)


_TEXT_FACTORIES = {"bert": bert, "distilbert": distilbert, "bart": bart}


def decision(probability: float) -> bool:
    """Host-side coin flip of the modality dropout (reference :110-111)."""
    return np.random.rand(1)[0] < probability


# ---------------------------------------------------------------------------------------------
# attention blocks: parameter holders; called on their own they run one fused fp32 CUDA kernel each (the bf16
# tensor-core pipeline exists as the whole fused head, MM_RCA.forward_features)
# ---------------------------------------------------------------------------------------------
class _AttentionBase(nn.Module):
    def __init__(self, d_in_q: int, d_in_kv: int, d_out_kq: int, d_out_v: int):
        super().__init__()
        self.d_out_kq = d_out_kq
        self.W_query = nn.Linear(d_in_q, d_out_kq)
        self.W_key = nn.Linear(d_in_kv, d_out_kq)
        self.W_value = nn.Linear(d_in_kv, d_out_v)
        self.norm = nn.LayerNorm(d_out_v)
        self.relu = nn.ReLU()

    def _params(self) -> Tuple[torch.Tensor, ...]:
        return (self.W_query.weight, self.W_query.bias, self.W_key.weight, self.W_key.bias,
                self.W_value.weight, self.W_value.bias, self.norm.weight, self.norm.bias)


class SelfAttention(_AttentionBase):
    """Reference :39-68: softmax(QK^T/sqrt(d_kq)) V -> LayerNorm -> ReLU, one kernel."""

    def __init__(self, d_in, d_out_kq, d_out_v, name):
        super().__init__(d_in, d_in, d_out_kq, d_out_v)
        self.name = name

    def forward(self, x):
        return F.attention_block(x, None, self._params(), reverse=False)


class ReverseCrossAttention(_AttentionBase):
    """Reference :71-108: queries from x_1, keys/values from x_2; with `reverse` the weights become
    (1 - A) / (L - 1) (:97-98)."""

    def __init__(self, d_in_x1, d_in_x2, d_out_kq, d_out_v, reverse):
        super().__init__(d_in_x1, d_in_x2, d_out_kq, d_out_v)
        self.reverse = reverse

    def forward(self, x_1, x_2):
        if x_1.shape[1] != x_2.shape[1]:
            raise AssertionError("ReverseCrossAttention needs square attention (reference :93)")
        return F.attention_block(x_1, x_2, self._params(), reverse=bool(self.reverse))


class Hadamard2(nn.Module):
    """Reference :822-831 (dead for MM_RCA; kept because its tensors are in every checkpoint)."""

    def __init__(self, dim):
        super().__init__()
        self.kernel1 = nn.Parameter(torch.randn(dim))
        self.kernel2 = nn.Parameter(torch.randn(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x1, x2):
        return torch.tanh(x1 * self.kernel1 + x2 * self.kernel2 + self.bias)


# ---------------------------------------------------------------------------------------------
# the shared base: allocates the parameters of EVERY fusion variant so they share one state_dict
# ---------------------------------------------------------------------------------------------
class EffV2MediumAndDistilbertGated(nn.Module):
    """Shared constructor of all fusion variants (reference :156-328).  Registration order and names
    reproduce the reference state_dict (tests/golden/state_dict_layout.json)."""

    def __init__(self, n_classes, drop_ratio, image_or_text_dropout_chance, img_prob_dropout, num_neurons_fc,
                 text_model_name, batch_size=16, reverse=False, features_only=False, cross_attention_only=False,
                 *, pretrained: bool = True, compute: int = N.COMPUTE_FP32,
                 dropout_generator: Optional[torch.Generator] = None):
        super().__init__()
        self.text_model_name = text_model_name
        self.features_only = features_only
        self.cross_attention_only = cross_attention_only
        self.reverse = reverse
        self.n_classes = n_classes
        self.compute = compute
        self.dropout_generator = dropout_generator
        print("Only features:", self.features_only)
        print("Only cross attention:", self.cross_attention_only)
        if text_model_name not in _TEXT_FACTORIES:
            print("Wrong text model:", text_model_name)
            sys.exit(1)
        self.text_model = _TEXT_FACTORIES[text_model_name](pretrained)
        self.image_model = eff_net_v2(pretrained)

        self.fc_layer_neurons = num_neurons_fc
        self.image_or_text_dropout_chance = image_or_text_dropout_chance
        self.img_dropout_prob = img_prob_dropout
        self.batch_size = batch_size
        self.gated_output_hidden_size = 256
        self.num_patches = F.NUM_PATCHES
        hidden_txt = self.text_model.config.hidden_size
        print("Text model hidden size:", hidden_txt)
        input_size_txt, input_size_img = 768, IMAGE_FEATURES          # literals, reference :257-258
        self.txt_patch_size = input_size_txt // self.num_patches
        self.img_patch_size = input_size_img // self.num_patches
        print("txt patch size: ", self.txt_patch_size)
        print("img patch size: ", self.img_patch_size)
        self.modality_dim = 400
        ca_flat = F.CA_DV * self.num_patches * 2                     # 1536

        h, g = num_neurons_fc, self.gated_output_hidden_size
        table: List[Tuple[str, Optional[Callable[[], nn.Module]]]] = [
            ("drop", lambda: nn.Dropout(p=drop_ratio)),
            ("image_dropout", lambda: nn.Dropout2d(p=1.0)),
            ("text_dropout", lambda: nn.Dropout1d(p=1.0)),
            ("image_to_hidden_size", lambda: nn.Linear(IMAGE_FEATURES, h)),
            ("text_to_hidden_size", lambda: nn.Linear(hidden_txt, h)),
            ("concat_layer", lambda: nn.Linear(2 * h, h)),
            ("fc_layer", lambda: nn.Linear(h, n_classes)),
            ("hyper_tang_layer", nn.Tanh),
            ("softmax_layer", lambda: nn.Softmax(dim=1)),
            ("image_features_hidden_layer", lambda: nn.Linear(IMAGE_FEATURES, g)),
            ("text_features_hidden_layer", lambda: nn.Linear(hidden_txt, g)),
            ("z_layer", lambda: nn.Linear(2 * g, g)),
            ("fc_layer_gated", lambda: nn.Linear(g, n_classes)),
            ("clip_fc_layer", lambda: nn.Linear(batch_size, n_classes)),
            ("trans_conv", lambda: nn.ConvTranspose1d(8, 8, kernel_size=2, stride=2, padding=0, output_padding=0)),
            ("logit_scale", None),
            ("output_all_features", lambda: nn.Linear(640, 4)),
            ("self_attention_image", lambda: SelfAttention(self.img_patch_size, F.SA_DKQ, F.SA_DV, "Img block")),
            ("self_attention_text", lambda: SelfAttention(self.txt_patch_size, F.SA_DKQ, F.SA_DV, "Txt block")),
            ("cross_attention_1", lambda: ReverseCrossAttention(F.SA_DV, F.SA_DV, F.CA_DKQ, F.CA_DV, reverse)),
            ("cross_attention_2", lambda: ReverseCrossAttention(F.SA_DV, F.SA_DV, F.CA_DKQ, F.CA_DV, reverse)),
            ("final", lambda: nn.Linear(ca_flat, n_classes)),
            ("final_features_only_linear",
             (lambda: nn.Linear(IMAGE_FEATURES + 768, n_classes)) if features_only else "skip"),
            ("cross_attention_only_linear",
             (lambda: nn.Linear(ca_flat, n_classes)) if cross_attention_only else "skip"),
            ("final_with_everything", lambda: nn.Linear(ca_flat + IMAGE_FEATURES + 768, n_classes)),
            ("final_hierarchical_image", lambda: nn.Linear(1280 + 2560 + 2048, 512)),
            ("final_hierarchical_text", lambda: nn.Linear(768 * 3, 512)),
            ("final_hierarchical_all", lambda: nn.Linear(512 * 2, n_classes)),
            ("relu", nn.ReLU),
            ("gru_text", lambda: nn.GRU(self.modality_dim, self.modality_dim, batch_first=True)),
            ("gru_audio", lambda: nn.GRU(self.modality_dim, self.modality_dim, batch_first=True)),
            ("fusion", lambda: Hadamard2(self.modality_dim)),
            ("gru_bimodal", lambda: _quiet_gru(self.modality_dim, 500)),
            ("dropout1", lambda: nn.Dropout(0.86)),
            ("concat_fc", lambda: nn.Linear(self.modality_dim + 500, 450)),
            ("dropout2", lambda: nn.Dropout(0.86)),
            ("modality_image_to_dim", lambda: nn.Linear(1280, self.modality_dim)),
            ("modality_text_to_dim", lambda: nn.Linear(768, self.modality_dim)),
            ("classifier", lambda: nn.Linear(450, 4)),
        ]
        for name, factory in table:
            if factory == "skip":
                continue
            if name == "logit_scale":
                self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))   # reference :244-245
                continue
            setattr(self, name, factory())
        self._head_names = F.head_param_names(bool(features_only), bool(cross_attention_only))

    # ---- helpers the drivers call (reference :397-418) -------------------------------------------
    def get_tokenizer(self):
        from transformers import BartTokenizer, BertTokenizer, DistilBertTokenizer
        src = {"bert": (BertTokenizer, "bert-base-uncased"),
               "distilbert": (DistilBertTokenizer, "distilbert-base-uncased"),
               "bart": (BartTokenizer, "facebook/bart-large")}[self.text_model_name]
        self.tokenizer = src[0].from_pretrained(src[1])
        return self.tokenizer

    def get_image_size(self):
        return (480, 480)

    def get_max_token_size(self):
        from transformers import BartConfig, BertConfig, DistilBertConfig
        cfg = {"bert": BertConfig, "distilbert": DistilBertConfig, "bart": BartConfig}[self.text_model_name]
        self.config = cfg().max_position_embeddings
        return self.config

    def drop_modalities(self, _eval, remove_image, remove_text):
        """Host-side modality gating (reference :420-455): zero the whole image batch or the
        token ids + attention mask; in training two numpy coin flips decide."""
        zero_image = zero_text = False
        if _eval:
            zero_image, zero_text = bool(remove_image), bool(remove_text)
            if zero_image:
                print("    Eval: zero image")
            if zero_text:
                print("    Eval: zero text")
        elif decision(self.image_or_text_dropout_chance):
            if decision(self.img_dropout_prob):
                print("    Train: zeroing image\n")
                zero_image = True
            else:
                print("    Train: zeroing text\n")
                zero_text = True
        if zero_image:
            self._images = torch.zeros_like(self._images)
        if zero_text:
            self._input_ids = torch.zeros_like(self._input_ids)
            self._attention_mask = torch.zeros_like(self._attention_mask)

    # ---- shared pieces of the subclass forwards -------------------------------------------------
    def _backbone_features(self, output_hidden_states: bool = False):
        kw = {"output_hidden_states": True} if output_hidden_states else {}
        text_output = self.text_model(input_ids=self._input_ids, attention_mask=self._attention_mask, **kw)
        text_features = text_output[0][:, 0]
        image_out = self.image_model(self._images)
        return text_output, text_features, image_out

    def head_parameters(self) -> List[torch.Tensor]:
        """The tensors MM_RCA.forward reads, in MmrcaHeadParams order (looked up once: walking the 1 300 named parameters
        of the module on every forward costs more host time than the head's kernels take on the GPU)."""
        cache = self.__dict__.get("_head_param_cache")
        if cache is None:
            sd = dict(self.named_parameters())
            cache = [sd[n] for n in self._head_names]
            self.__dict__["_head_param_cache"] = cache
        return cache

    def attach_flat_grads(self, flat_params: bool = False):
        """Give the head a persistent flat gradient bucket: every head parameter's .grad becomes a view of ONE fp32
        buffer that the backward kernels accumulate into directly (no per-backward allocation, no copies) and that a
        data-parallel step all-reduces as a single collective (training.HeadDataParallel semantics for the module path).
        flat_params=True also moves the parameters themselves into one buffer (training.FlatParams) so that
        training.FusedSGD / FusedAdamW update all of them with one kernel.  Returns the FlatGrads (``.flat`` is the
        bucket).  Call after .to(device); optimizer.zero_grad(set_to_none=False) / flat.zero_() keeps the views alive."""
        from . import training as T
        params = self.head_parameters()
        self._flat_grads = F.FlatGrads([p.detach() for p in params])
        T.attach_flat_grads(params, self._flat_grads)
        self._flat_params = T.FlatParams(params, self._flat_grads) if flat_params else None
        return self._flat_grads

    def forward(self, _input_ids, _attention_mask, _images, eval=False, remove_image=False, remove_text=False):
        raise NotImplementedError(
            "the gated fusion (reference :330-396) is outside the B200-native hot path: MM_RCA, Hierarchical, "
            "EffV2MediumAndDistilbertClassic and EffV2MediumAndDistilbertNormalized are rebuilt here; the gated / CLIP / "
            "bimodal variants keep their parameters (state_dict parity) but have no forward")

    def _seeded_dropout(self) -> Tuple[float, int]:
        """(p, seed) of this forward's self.drop: the seed comes from torch's CPU generator (or `dropout_generator`),
        the keep mask itself is drawn inside the kernels (F.dropout_mask(seed, p, ...) returns it)."""
        p = float(self.drop.p)
        if not self.training or p <= 0.0:
            return 0.0, 0
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=self.dropout_generator).item())
        self.last_dropout_seed = seed
        return min(p, 1.0), seed


def _quiet_gru(inp, hid):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return nn.GRU(inp, hid, batch_first=True, dropout=0.35)   # reference :318 (warns: 1 layer + dropout)


class MM_RCA(EffV2MediumAndDistilbertGated):
    """Multimodal reverse-cross-attention classifier (reference :636-728)."""

    def _dropout_seed(self) -> Tuple[float, int]:
        """(p, seed) of this forward's dropout (self.drop, reference :190/:719).  The seed comes from torch's
        CPU generator (or `dropout_generator`), so torch.manual_seed makes training runs repeatable; the keep
        mask itself is drawn inside the kernels (F.dropout_mask(seed, p, ...) returns it)."""
        p = float(self.drop.p)
        if not self.training or p <= 0.0:
            return 0.0, 0
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=self.dropout_generator).item())
        self.last_dropout_seed = seed
        return min(p, 1.0), seed

    def forward_features(self, image_features: torch.Tensor, text_features: torch.Tensor,
                         drop_mask: Optional[torch.Tensor] = None, drop_scale: Optional[float] = None):
        """Fusion head on pooled features [B,1280] / [B,768] (reference :661-728).  `drop_mask` overrides
        the internally drawn dropout mask (parity tests pass the mask torch.nn.Dropout drew)."""
        bf16_ok = (self.compute != N.COMPUTE_FP32 and drop_mask is None and image_features.dtype == torch.bfloat16
                   and text_features.dtype == torch.bfloat16
                   and not (image_features.requires_grad or text_features.requires_grad))
        if not bf16_ok:      # bf16 features of frozen backbones go to the bf16 pipeline as they are (MMRCA_FLAG_FEATURES_BF16)
            image_features = image_features.float()
            text_features = text_features.float()
        drop_p, drop_seed = 0.0, 0
        if drop_mask is None:
            drop_p, drop_seed = self._dropout_seed()
            drop_scale = 1.0
        elif drop_scale is None:
            drop_scale = 1.0 / (1.0 - float(self.drop.p))
        return F.mmrca_head(image_features, text_features, self.head_parameters(), reverse=bool(self.reverse),
                            features_only=bool(self.features_only),
                            cross_attention_only=bool(self.cross_attention_only), n_classes=self.n_classes,
                            drop_mask=drop_mask, drop_scale=drop_scale, drop_p=drop_p, drop_seed=drop_seed,
                            compute=self.compute, grad_sink=getattr(self, "_flat_grads", None))

    def enable_feature_cache(self, n_samples: int, dtype: torch.dtype = torch.bfloat16):
        """Frozen-backbone phase only: keep every sample's pooled features on the device (training.FeatureCache) and skip
        both backbones when forward() is given `sample_ids` that are all cached.  The reference recomputes the frozen
        backbones every step (main_both.py:562-577)."""
        from .training import FeatureCache
        dev = next(self.parameters()).device
        self.feature_cache = FeatureCache(n_samples, IMAGE_FEATURES, 768, dev, dtype)
        return self.feature_cache

    def _frozen(self) -> bool:
        """Transfer-learning phase: the reference freezes / unfreezes each backbone as a whole (:113-153,
        main_both.py:687-694), so one parameter per backbone tells."""
        for m in (self.text_model, self.image_model):
            p = next(iter(m.parameters()), None)
            if p is not None and p.requires_grad:
                return False
        return True

    def backbone_features(self):
        """Pooled image / text features of the current inputs.  Frozen backbones on a GPU hand their raw outputs (last
        hidden state, final feature map) to ONE hand-off kernel (functional.feature_handoff: CLS gather + average pool +
        cast, bf16 out for the bf16 pipeline); otherwise the stock path of the reference (:651-659)."""
        fused = self._images.is_cuda and self._frozen() and isinstance(self.image_model, EfficientNetV2MFullFeatureExtractor)
        if not fused:
            _, text_features, (_, _, image_features) = self._backbone_features()
            return image_features, text_features
        with torch.no_grad():
            hidden = self.text_model(input_ids=self._input_ids, attention_mask=self._attention_mask)[0]
            _, _, fmap = self.image_model.forward_maps(self._images)
            if hidden.dtype not in (torch.float32, torch.bfloat16):
                hidden = hidden.float()
            if fmap.dtype not in (torch.float32, torch.bfloat16):
                fmap = fmap.float()
            out_dtype = torch.bfloat16 if self.compute != N.COMPUTE_FP32 else torch.float32
            return F.feature_handoff(hidden, fmap, out_dtype)

    def forward(self, _input_ids, _attention_mask, _images, eval=False, remove_image=False, remove_text=False,
                sample_ids: Optional[torch.Tensor] = None):
        self._images = _images
        self._input_ids = _input_ids
        self._attention_mask = _attention_mask
        self.drop_modalities(eval, remove_image, remove_text)
        cache = getattr(self, "feature_cache", None)
        # the cache holds the features of the UNMODIFIED inputs: not when a modality is removed (eval flags) or may be
        # dropped at random (training with image_or_text_dropout_chance > 0, reference :444-452)
        use_cache = (cache is not None and sample_ids is not None and self._frozen() and not (remove_image or remove_text)
                     and (eval or self.image_or_text_dropout_chance == 0))
        feats = cache.lookup(sample_ids) if use_cache else None
        if feats is None:
            feats = self.backbone_features()
            if use_cache:
                cache.store(sample_ids, *feats)
        return self.forward_features(*feats)


class Hierarchical(EffV2MediumAndDistilbertGated):
    """Hierarchical late fusion (reference :729-818, `--late_fusion=hierarchical`): multi-scale image features (pooled
    vector + average-pooled 160- and 512-channel stage maps) and the CLS embeddings of three text layers, six L2
    normalisations, two concats, dropout, Linear(5888 -> 512) / Linear(2304 -> 512) with ReLU, Linear(1024 -> n_classes).
    The backbones and the two AvgPool2d stay stock torch; everything after them is one call into libmmrca.so
    (bf16 tcgen05 GEMMs, frozen backbones).  Like the reference it assumes EfficientNetV2-M at 480 x 480."""

    def hier_parameters(self) -> List[torch.Tensor]:
        cache = self.__dict__.get("_hier_param_cache")
        if cache is None:
            sd = dict(self.named_parameters())
            cache = [sd[n] for n in F.HIER_PARAM_NAMES]
            self.__dict__["_hier_param_cache"] = cache
        return cache

    def forward_features(self, feats, drop_mask: Optional[torch.Tensor] = None, drop_scale: Optional[float] = None):
        """The head on the six pooled feature tensors (reference :777-816)."""
        drop_p, drop_seed = 0.0, 0
        if drop_mask is None:
            p = float(self.drop.p)
            if self.training and p > 0.0:
                drop_p = min(p, 1.0)
                drop_seed = int(torch.randint(0, 2 ** 62, (1,), generator=self.dropout_generator).item())
                self.last_dropout_seed = drop_seed
            drop_scale = 1.0
        elif drop_scale is None:
            drop_scale = 1.0 / (1.0 - float(self.drop.p))
        return F.hierarchical_head(feats, self.hier_parameters(), drop_mask=drop_mask, drop_scale=drop_scale,
                                   drop_p=drop_p, drop_seed=drop_seed)

    def forward(self, _input_ids, _attention_mask, _images, eval=False, remove_image=False, remove_text=False):
        self._images = _images
        self._input_ids = _input_ids
        self._attention_mask = _attention_mask
        self.drop_modalities(eval, remove_image, remove_text)
        text_output, text_features, (out_stage_3, out_stage_6, image_features) = self._backbone_features(True)
        hidden_states = text_output.hidden_states
        layer_2, layer_4 = hidden_states[2][:, 0, :], hidden_states[4][:, 0, :]                    # reference :756-757
        s3 = torch.nn.functional.avg_pool2d(out_stage_3, kernel_size=7, stride=7).flatten(1)       # :761-762, :769
        s6 = torch.nn.functional.avg_pool2d(out_stage_6, kernel_size=6, stride=6).flatten(1)       # :765-766, :772
        return self.forward_features((image_features, s3, s6, text_features, layer_2, layer_4))


class EffV2MediumAndDistilbertClassic(EffV2MediumAndDistilbertGated):
    """Classic late fusion (reference :489-534, `--late_fusion=classic`): Linear(1280 -> H) and Linear(768 -> H) projections
    into the shared fusion dimension H = num_neurons_FC, concat, Linear(2H -> H), dropout, Linear(H -> n_classes); the
    Normalized subclass L2-normalises the two projections first.  Everything after the backbones is libmmrca.so
    (mmrca_fusion_*).  The reference passes the image extractor's whole (stage3, stage6, pooled) tuple to
    image_to_hidden_size (:519-521), which raises a TypeError as shipped; the pooled vector is what is fed here."""

    NORMALIZED = False

    def fusion_parameters(self) -> List[torch.Tensor]:
        cache = self.__dict__.get("_fusion_param_cache")
        if cache is None:
            sd = dict(self.named_parameters())
            cache = [sd[n] for n in F.FUSION_PARAM_NAMES]
            self.__dict__["_fusion_param_cache"] = cache
        return cache

    def forward_features(self, image_features: torch.Tensor, text_features: torch.Tensor,
                         drop_mask: Optional[torch.Tensor] = None, drop_scale: Optional[float] = None):
        drop_p, drop_seed = 0.0, 0
        if drop_mask is None:
            drop_p, drop_seed = self._seeded_dropout()
            drop_scale = 1.0
        elif drop_scale is None:
            drop_scale = 1.0 / (1.0 - float(self.drop.p))
        # head_compute = "bf16" (attribute; default "fp32", the 1e-4 contract): the Linear layers on the tensor cores
        return F.fusion_head(image_features.float(), text_features.float(), self.fusion_parameters(),
                             normalized=self.NORMALIZED, drop_mask=drop_mask, drop_scale=drop_scale, drop_p=drop_p,
                             drop_seed=drop_seed, compute=getattr(self, "head_compute", "fp32"))

    def forward(self, _input_ids, _attention_mask, _images, eval=False, remove_image=False, remove_text=False):
        print("Normalized forward" if self.NORMALIZED else "Classic forward")       # reference :500, :545
        self._images = _images
        self._input_ids = _input_ids
        self._attention_mask = _attention_mask
        self.drop_modalities(eval, remove_image, remove_text)
        _, text_features, (_, _, image_features) = self._backbone_features()
        return self.forward_features(image_features, text_features)


class EffV2MediumAndDistilbertNormalized(EffV2MediumAndDistilbertClassic):
    """Normalized late fusion (reference :536-579, `--late_fusion=normalized`)."""

    NORMALIZED = True


def load_reference_state_dict(model: nn.Module, state_dict, strict: bool = True):
    """load_state_dict that also accepts checkpoints saved from an nn.DataParallel wrapper
    (`module.` prefix)."""
    if state_dict and all(k.startswith("module.") for k in state_dict):
        state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
    return model.load_state_dict(state_dict, strict=strict)
