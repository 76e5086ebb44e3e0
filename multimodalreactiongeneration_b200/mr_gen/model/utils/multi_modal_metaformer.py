"""Mirror of mr_gen/model/utils/multi_modal_metaformer.py — the multi-modal "metaformer" of lstmformer:
``gen_attention_mask`` :32-79, ``MultiModalEmbedding`` :82-128, ``IntegrateModalBlock`` :131-229,
``MultiModalMetaformerBlock`` :232-353, ``MultiModalMetaformer`` :356-509.

Per block: every modality runs through its own stack of token mixers (``embedding`` — the LSTM mixers are the hot
path, on ``B200LSTM``), the main modality then attends to each other modality (``integrator``: one masked
multi-head attention stack per other modality, concatenated and mixed by ``cat_linear``), then a feed-forward.
Attribute names are the reference's, so ``state_dict`` keys are identical.

Reference behaviour kept on purpose: Q3 — every mixer stack hands back what is LEFT of its input state list, so
the per-block ``{"emb": [...], "crm": [...]}`` dicts that ``forward`` returns only ever hold ``None`` leaves
(autoregressive generation is stateless per step).

Difference in representation only: ``gen_attention_mask`` returns the ``[B, heads, L, S]`` mask as a broadcast
VIEW of a ``[B, 1, L, S]`` tensor computed by index arithmetic (the reference materialises the tiled triangle and
``repeat``s it per head: 92 MB of bools per mask at B=256, 4 heads, T=300), and on CUDA ``Metaformer`` does not
build the tensor at all: ``gen_attention_mask_spec`` hands the fused attention kernel the rule itself."""
from typing import Any, List, Optional, Tuple, Union

import torch
from torch import nn

from ....attention import AttentionMaskSpec
from ....linear import B200Linear
from .mixer_block import FeedForward, MixerLayerdFactory, split_state
from .residual_connection import ResidualConnection
from .values import PADDING_VALUE


def check_form_modal_num(modal_num: int, sameone, msg: str = None) -> list:
    """``sameone`` must hold one entry per modality; a single entry is repeated."""
    if not isinstance(sameone, (list, tuple)):
        raise TypeError(f"must be list or tuple. but got {type(sameone)}")
    if len(sameone) != modal_num:
        if len(sameone) != 1:
            raise ValueError("" if msg is None else msg)
        sameone = sameone * modal_num
    return sameone


def gen_attention_mask_spec(main_modal: torch.Tensor, other_modal: torch.Tensor,
                            padding_value: float = PADDING_VALUE) -> AttentionMaskSpec:
    """The reference's mask as a RULE (mode, rate, per-frame padding bytes) for the fused attention kernel, which
    evaluates it on the fly and skips key tiles that are masked for a whole query tile."""
    L, S = main_modal.shape[1], other_modal.shape[1]
    if S % L != 0 and L % S != 0:
        raise ValueError(f"other_modal_len must be divisible by main_modal_len. "
                         f"main_modal_len: {L}, other_modal_len: {S}")
    mode, rate = (1, S // L) if S % L == 0 else (2, L // S)
    pad_q = (main_modal[:, :, 0] == padding_value).to(torch.uint8).contiguous()
    pad_k = (other_modal[:, :, 0] == padding_value).to(torch.uint8).contiguous()
    return AttentionMaskSpec(mode, rate, pad_q, pad_k)


def gen_attention_mask(main_modal: torch.Tensor, other_modal: torch.Tensor, head_num: int,
                       padding_value: float = PADDING_VALUE) -> torch.Tensor:
    """bool ``[B, head_num, L, S]``, True = may NOT attend.  Causal between two streams whose frame rates differ
    by an integer factor: with ``S = r*L`` query frame i sees the keys of frames ``<= i`` (``j // r <= i``); with
    ``L = r*S`` query i sees keys ``j <= i // r``.  A (query, key) pair that is padding on BOTH sides (first
    feature == ``padding_value``) is masked as well."""
    return gen_attention_mask_spec(main_modal, other_modal, padding_value).materialize(head_num)


class MultiModalEmbedding(nn.Module):
    """One ``*MixerLayerd`` per modality; returns (outputs, remaining states)."""

    def __init__(self, modal_num: int, mixer_configs):
        super().__init__()
        self.modal_num = modal_num
        self.mixer_configs = check_form_modal_num(
            modal_num, mixer_configs,
            f"modal_num must be equal to the length of mixer_configs."
            f"modal_num: {modal_num}, mixer_configs: {len(mixer_configs)}")
        factory = MixerLayerdFactory()
        self.modal_embeddings = nn.ModuleList(factory.build(*cfg) for cfg in self.mixer_configs)

    def forward(self, x: List[torch.Tensor], hx: Optional[list] = None, other: Optional[List[tuple]] = None):
        x = check_form_modal_num(self.modal_num, x,
                                 f"The length of x must be equal to modal_num."
                                 f"modal_num: {self.modal_num}, x: {len(x)}")
        hx = [None] * self.modal_num if hx is None else hx
        other = [(None,)] * self.modal_num if other is None else other
        outs, states = [], []
        # the modality stacks are independent of each other: they run side by side on the library's streams when
        # the caller has forked them (Metaformer.forward); here they are simply issued in order
        for i, embed in enumerate(self.modal_embeddings):
            y, state, _ = embed(x[i], hx[i], other[i])
            outs.append(y)
            states.append(state)
        return outs, states


class IntegrateModalBlock(nn.Module):
    """main modality (query) x every other modality (key = value) -> concat -> ``cat_linear``."""

    def __init__(self, modal_num: int, mixer_configs, output_dim: int):
        super().__init__()
        self.modal_num = modal_num
        self.mixer_configs = check_form_modal_num(
            modal_num - 1, mixer_configs,
            f"modal_num must be equal to the length + 1 of mixer_configs. "
            f"modal_num: {modal_num}, mixer_configs: {len(mixer_configs)} (+1)")
        factory = MixerLayerdFactory()
        self.integrators = nn.ModuleList()
        width = 0
        for kind, cfg in self.mixer_configs:
            if kind != "mha":
                raise TypeError("IntegrateModalBlock only supports mha mixer.")
            self.integrators.append(factory.build(kind, cfg))
            width += cfg["output_projection_size"] if cfg["output_projection"] else cfg["hidden_size"]
        self.cat_linear = B200Linear(width, output_dim)

    def check_form_input(self, other_modals, attn_mask, hx=None):
        n = self.modal_num - 1
        if isinstance(other_modals, torch.Tensor):
            other_modals = [other_modals]
        other_modals = check_form_modal_num(
            n, other_modals, f"The length of other_modals must be equal to modal_num - 1. "
                             f"modal_num: {self.modal_num}, other_modals: {len(other_modals)}")
        if attn_mask is None:
            attn_mask = [None] * n
        elif isinstance(attn_mask, torch.Tensor):
            attn_mask = [attn_mask]
        attn_mask = check_form_modal_num(
            n, attn_mask, f"The length of attn_mask must be equal to modal_num - 1. "
                          f"modal_num: {self.modal_num}, attn_mask: {len(attn_mask)}")
        return other_modals, attn_mask, ([None] * n if hx is None else hx)

    def forward(self, main_modal: torch.Tensor, other_modals, attn_mask=None, hxs=None):
        other_modals, attn_mask, hxs = self.check_form_input(other_modals, attn_mask, hxs)
        outs, states = [], []
        for i, integrator in enumerate(self.integrators):
            y, state, _ = integrator(main_modal, hxs[i], other_modals[i], other_modals[i], attn_mask[i])
            outs.append(y)
            states.append(state)
        return self.cat_linear(torch.cat(outs, dim=-1)), states


class MultiModalMetaformerBlock(nn.Module):
    def __init__(self, num_modal: int, main_modal_embedding_config, integrate_configs, feedforward_configs: dict,
                 encode_other_modal: bool = False, other_modal_embedding_config=None):
        super().__init__()
        if not encode_other_modal or other_modal_embedding_config is None:
            other_modal_embedding_config = []
        if isinstance(main_modal_embedding_config, tuple):
            main_modal_embedding_config = [main_modal_embedding_config]
        if encode_other_modal:
            other_modal_embedding_config = check_form_modal_num(
                num_modal - 1, other_modal_embedding_config,
                f"The length of other_modal_embedding_config must be equal to num_modal - 1. num_modal: "
                f"{num_modal}, other_modal_embedding_config: {len(other_modal_embedding_config)}")
        integrate_configs = check_form_modal_num(
            num_modal - 1, integrate_configs,
            f"The length of integrate_configs must be equal to num_modal - 1. "
            f"num_modal: {num_modal}, integrate_configs: {len(integrate_configs)}")
        self.num_modal = num_modal
        self.emb_num_modal = num_modal if encode_other_modal else 1
        self.main_modal_embedding_config = main_modal_embedding_config
        self.integrate_configs = integrate_configs
        self.encode_other_modal = encode_other_modal
        self.other_modal_embedding_config = other_modal_embedding_config
        self.embedding_configs = list(main_modal_embedding_config) + list(other_modal_embedding_config)
        self.emb_mixer_type = [cfg[0] for cfg in self.embedding_configs]

        self.embedding = MultiModalEmbedding(self.emb_num_modal, self.embedding_configs)
        self.integrator = IntegrateModalBlock(num_modal, integrate_configs, feedforward_configs["hidden_size"])
        self.feedforward = FeedForward(**feedforward_configs)

    def forward(self, main_modal, other_modals: List[torch.Tensor] = None, hx: Optional[list] = None,
                prev_hx: Optional[list] = None, main_modal_others: Tuple[Any] = None,
                other_modals_others: List[Tuple[Any]] = None, integrate_attn_mask: List[torch.Tensor] = None):
        if isinstance(main_modal, tuple):  # chained call: everything travels in one tuple
            (main_modal, other_modals, hx, prev_hx, main_modal_others, other_modals_others,
             integrate_attn_mask) = main_modal[:7]
        state, hx, prev_hx = split_state(hx, prev_hx)
        if state is None:
            state = {"emb": None, "crm": None}
        other_modals_others = check_form_modal_num(
            self.num_modal - 1, other_modals_others if other_modals_others else [None],
            f"The length of other_modals_others must be equal to num_modal - 1."
            f"num_modal: {self.num_modal}, other_modals_others: {len(other_modals_others or [None])}")

        streams = [main_modal] + list(other_modals) if self.encode_other_modal else [main_modal]
        modals, emb_states = self.embedding(streams, state["emb"], [main_modal_others] + list(other_modals_others))
        main_modal = modals[0]
        if self.encode_other_modal:
            other_modals = modals[1:]
        main_modal, crm_states = self.integrator.forward(main_modal, other_modals, integrate_attn_mask,
                                                         state["crm"])
        prev_hx.append({"emb": emb_states, "crm": crm_states})
        main_modal = self.feedforward(main_modal)
        return (main_modal, other_modals, hx, prev_hx, main_modal_others, other_modals_others,
                integrate_attn_mask)


class MultiModalMetaformer(nn.Module):
    def __init__(self, modal_num: int, hidden_dim: int, num_layer: int,
                 main_modal_feature_dim: Union[List[int], int], main_mixer_type, main_mixer_configs,
                 integrate_mixer_configs, feedforward_configs: dict, output_feedforward_configs: dict,
                 other_modal_feature_dim: Union[List[int], int] = None, other_mixer_type="mha",
                 other_mixer_configs=None, repeat_with_encoder: bool = False, interlayer_residual: bool = False,
                 interlayer_residual_norm: bool = True):
        super().__init__()
        n_other = modal_num - 1

        def first(v):
            return v[0] if isinstance(v, (list, tuple)) else v

        main_modal_feature_dim = first(main_modal_feature_dim)
        main_mixer_type = first(main_mixer_type)
        main_cfg = [(main_mixer_type, first(main_mixer_configs))]
        def per_other(value, name, scalar_types):
            """one entry per non-main modality: a bare value is wrapped, a single entry is repeated"""
            if isinstance(value, scalar_types):
                value = [value]
            return check_form_modal_num(
                n_other, value, f"The length of {name} must be equal to modal_num - 1."
                                f"modal_num: {modal_num}, {name}: {len(value)}")

        integrate_cfg = [("mha", c) for c in per_other(integrate_mixer_configs, "integrate_mixer_configs", dict)]
        other_modal_feature_dim = per_other(other_modal_feature_dim, "other_modal_feature_dim", int)
        other_mixer_type = per_other(other_mixer_type, "other_mixer_type", str)
        other_mixer_configs = per_other(other_mixer_configs, "other_mixer_configs", dict)
        other_cfg = [(other_mixer_type[i], other_mixer_configs[i]) for i in range(n_other)]

        self.modal_num = modal_num
        self.hidden_dim = hidden_dim
        self.num_layer = num_layer
        self.repeat_with_encoder = repeat_with_encoder
        self.interlayer_residual = interlayer_residual
        self.embedding_mixer_type = [main_mixer_type] + list(other_mixer_type)

        self.feature_embedding = nn.ModuleList(
            B200Linear(dim, hidden_dim) for dim in [main_modal_feature_dim] + list(other_modal_feature_dim))
        # only the first block encodes the other modalities unless repeat_with_encoder
        blocks = [MultiModalMetaformerBlock(modal_num, main_cfg, integrate_cfg, feedforward_configs,
                                            encode_other_modal=True, other_modal_embedding_config=other_cfg)]
        for _ in range(num_layer - 1):
            blocks.append(MultiModalMetaformerBlock(
                modal_num, main_cfg, integrate_cfg, feedforward_configs, encode_other_modal=repeat_with_encoder,
                other_modal_embedding_config=other_cfg if repeat_with_encoder else None))
        self.metaformer_blocks = nn.ModuleList(
            ResidualConnection(b, interlayer_residual_norm, hidden_dim) if interlayer_residual else b
            for b in blocks)
        self.output_feedforward = FeedForward(**output_feedforward_configs)

    def forward(self, main_modal: torch.Tensor, other_modals: List[torch.Tensor], hx: Optional[list] = None,
                main_modal_others: Tuple[Any] = None, other_modals_others: List[Tuple[Any]] = None,
                integrate_attn_mask: List[torch.Tensor] = None):
        main_modal = self.feature_embedding[0](main_modal)
        other_modals = [self.feature_embedding[i + 1](m) for i, m in enumerate(other_modals)]
        rest = [other_modals, hx, None, main_modal_others, other_modals_others, integrate_attn_mask]
        for block in self.metaformer_blocks:
            main_modal, *rest = block(main_modal, *rest)
        # rest = [other_modals, remaining input states, collected states, ...]: the reference hands back the COLLECTED
        # list (:506-509 swaps the two names) — one {"emb": [...], "crm": [...]} dict per block whose leaves are the
        # mixers' "remaining input state", i.e. None (Q3)
        return self.output_feedforward(main_modal), rest[0], rest[2]
