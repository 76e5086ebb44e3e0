"""Mirror of the LSTM half of mr_gen/model/utils/mixer_block.py — the token mixers of lstmformer:
``split_state`` :21-30, ``FeedForward`` :37-87, ``LSTMMixer`` :211-252, ``LSTMMixerBlock`` :431-507,
``LSTMMixerLayerd`` :762-843, with ``nn.LSTM`` replaced by ``B200LSTM``.

Quirk Q3 is kept: ``LSTMMixerLayerd.forward`` returns the REMAINING input state
list (``hx`` after ``split_state``), not the collected new states, so callers only ever see ``None``.
The mixers lstmformer puts around them are mirrored as well so that ``Metaformer`` is a drop-in:
``MLPMixer*`` :114-166,308-352,605-676 (plain Linear stacks), ``MHAforSequentail`` (for_sequential.py:8-50),
``MHAMixer`` :255-305, ``MHAMixerBlock`` :508-602, ``MHAMixerLayerd`` :845-963 (cross-modal integrators; their
projections run on the tcgen05 GEMM through ``B200MultiheadAttention``), and the two factories :966-1017.
``GRUMixer`` :169-208, ``GRUMixerBlock`` :355-428, ``GRUMixerLayerd`` :679-759 run on
``B200GRU`` (csrc/mrg_gru.cu) — never cuDNN."""
from collections import OrderedDict
from typing import Any, List, Optional, Tuple, Union

import torch
from torch import nn

from ....gru import B200GRU
from ....lstm import B200LSTM
from ....linear import B200Linear
from ....attention import B200MultiheadAttention
from .residual_connection import ResidualConnection

LSTMStateType = Tuple[torch.Tensor, torch.Tensor]
DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")

_ACTIVATIONS = {"relu": nn.ReLU, "swish": nn.SiLU, "tanh": nn.Tanh}  # nonlinearity.py:6-16


def set_nonlinearity(name):
    """class of the activation, or None for ``None`` / ``"none"`` (nonlinearity.py:6-16)"""
    if name in _ACTIVATIONS:
        return _ACTIVATIONS[name]
    if name is None or name == "none":
        return None
    raise ValueError(f"nonlinearity must be in {sorted(_ACTIVATIONS) + [None]}")


def split_state(state: Optional[list], prev_state: Optional[list]):
    """Pop the first per-block state off ``state``.  Like the reference, the remainder of an exhausted
    list stays ``[]`` (not ``None``) and popping from an empty list raises IndexError."""
    head, rest = (None, None) if state is None else (state[0], state[1:])
    return head, rest, ([] if prev_state is None else prev_state)


class FeedForward(nn.Module):
    """``nonlinearity in (None, "none")`` -> a single Linear named ``feedforward``; otherwise
    input / activation / output.  Optionally wrapped in ResidualConnection (+LayerNorm)."""

    def __init__(self, hidden_size: int, bottleneck_size: int = None, output_size: int = None,
                 nonlinearity=None, residual: bool = False, residual_layer_norm: bool = False,
                 bias: bool = True, device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": DEVICE if device is None else device, "dtype": dtype}
        bottleneck_size = hidden_size if bottleneck_size is None else bottleneck_size
        output_size = hidden_size if output_size is None else output_size
        if residual and hidden_size != output_size:
            raise ValueError("hidden_size must be equal to output_size when residual is True.")
        if nonlinearity is None or nonlinearity == "none":
            layers = [("feedforward", B200Linear(hidden_size, output_size, **kw))]
        else:
            layers = [("input", B200Linear(hidden_size, bottleneck_size, **kw)),
                      ("activation", set_nonlinearity(nonlinearity)()),
                      ("output", B200Linear(bottleneck_size, output_size, **kw))]
        ff = nn.Sequential(OrderedDict(layers))
        self.feed_forward = ResidualConnection(ff, residual_layer_norm, hidden_size) if residual else ff

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.feed_forward(x)


class LSTMMixer(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, bias: bool = True,
                 batch_first: bool = True, dropout: float = 0.0, bidirectional: bool = False,
                 proj_size: int = 0, device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        if num_layers < 1:
            raise ValueError("num_layers must be greater than 0.")
        if bidirectional:
            if hidden_size % 2 != 0:
                raise ValueError("hidden_size must be even number when bidirectional is True.")
            hidden_size //= 2
        self.mixer = B200LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                              batch_first=batch_first, dropout=dropout, bidirectional=bidirectional,
                              proj_size=proj_size, bias=bias, device=DEVICE if device is None else device,
                              dtype=dtype)

    def forward(self, x: torch.Tensor, hn: Optional[LSTMStateType]):
        return self.mixer(x, hn)


class LSTMMixerBlock(nn.Module):
    """ResidualConnection(LSTMMixer) followed by FeedForward; threads (x, hx, prev_hx) like the
    reference so it can sit in an nn.Sequential or an autoregressive loop."""

    def __init__(self, hidden_size: int, num_layers: int = 1, dropout: float = 0.0, batch_first: bool = True,
                 bidirectional: bool = False, proj_size: int = 0, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        core = LSTMMixer(input_size=hidden_size, hidden_size=hidden_size, num_layers=num_layers,
                         batch_first=batch_first, dropout=dropout, bidirectional=bidirectional,
                         proj_size=proj_size, **kw)
        self.mixer = ResidualConnection(core, residual_layer_norm, hidden_size) if residual else core
        if residual and residual_layer_norm and device is not None:
            self.mixer.layer_norm.to(device)
        self.feed_forward = FeedForward(hidden_size=hidden_size, bottleneck_size=bottleneck_size,
                                        nonlinearity=nonlinearity, residual=residual,
                                        residual_layer_norm=residual_layer_norm, **kw)

    def forward(self, x: Union[torch.Tensor, tuple], hx: List[LSTMStateType] = None,
                prev_hx: List[LSTMStateType] = None):
        if isinstance(x, (tuple, list)):
            x, hx, prev_hx = x
        elif not isinstance(x, torch.Tensor):
            raise TypeError(f"x must be torch.Tensor or tuple or list, but got {type(x)}.")
        state, hx, prev_hx = split_state(hx, prev_hx)
        y, state = self.mixer(x, state)
        y = self.feed_forward(y)
        prev_hx.append(state)
        return (y, hx, prev_hx)


class LSTMMixerLayerd(nn.Module):
    def __init__(self, hidden_size: int, input_projection: bool = False, input_projection_size: int = None,
                 output_projection: bool = False, output_projection_size: int = None, num_layerd: int = 1,
                 num_internal_layer: int = 1, dropout: float = 0.0, batch_first: bool = True,
                 bidirectional: bool = False, proj_size: int = 0, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        if input_projection and input_projection_size is None:
            raise ValueError("input_projection_size must be specified when input_projection is True.")
        if output_projection and output_projection_size is None:
            raise ValueError("output_projection_size must be specified when output_projection is True.")
        self.input_projection = B200Linear(input_projection_size, hidden_size, **kw) if input_projection else None
        self.output_projection = (B200Linear(output_projection_size, hidden_size, **kw)
                                  if output_projection else None)
        self.mixer = nn.ModuleList(
            LSTMMixerBlock(hidden_size=hidden_size, num_layers=num_internal_layer, dropout=dropout,
                           batch_first=batch_first, bidirectional=bidirectional, proj_size=proj_size,
                           nonlinearity=nonlinearity, residual=residual,
                           residual_layer_norm=residual_layer_norm, bottleneck_size=bottleneck_size, **kw)
            for _ in range(num_layerd))

    def forward(self, x: torch.Tensor, hx: List[LSTMStateType] = None, other=(None,)):
        if self.input_projection is not None:
            x = self.input_projection(x)
        collected = None
        for block in self.mixer:
            x, hx, collected = block(x, hx, collected)
        if self.output_projection is not None:
            x = self.output_projection(x)
        return (x, hx, other)  # hx = what is left of the INPUT list (Q3), `collected` is dropped


# ------------------------------------------------------------------------------------------------------
# MLP mixer (Linear stacks; no state)
# ------------------------------------------------------------------------------------------------------
class MLPMixer(nn.Module):
    """input Linear (+act) -> (num_layer-1) x hidden Linear (+act) -> output Linear, registered as
    ``mixer.input.input``, ``mixer.hidden[i].hidden``, ``mixer.output`` (mixer_block.py:134-162)."""

    def __init__(self, input_size: int, hidden_size: int, num_layer: int = 1, nonlinearity=None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        if num_layer < 1:
            raise ValueError("num_layer must be greater than 0.")
        kw = {"bias": bias, "device": DEVICE if device is None else device, "dtype": dtype}
        act = set_nonlinearity(nonlinearity)
        self.nonlinearity = act

        def stage(name, n_in):
            parts = [(name, B200Linear(n_in, hidden_size, **kw))]
            if act is not None:
                parts.append(("activation", act()))
            return nn.Sequential(OrderedDict(parts))

        stages = [("input", stage("input", input_size))]
        stages += [(f"hidden[{i}]", stage("hidden", hidden_size)) for i in range(1, num_layer)]
        stages.append(("output", B200Linear(hidden_size, hidden_size, **kw)))
        # the reference also keeps the sub-blocks as attributes, which duplicates their state_dict keys
        self.input_block = stages[0][1]
        if num_layer > 1:
            self.hidden_block = stages[-2][1]
        self.output_block = stages[-1][1]
        self.mixer = nn.Sequential(OrderedDict(stages))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.mixer(x)


class MLPMixerBlock(nn.Module):
    def __init__(self, hidden_size: int, num_layer: int = 1, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        core = MLPMixer(input_size=hidden_size, hidden_size=hidden_size, num_layer=num_layer,
                        nonlinearity=nonlinearity, **kw)
        self.mixer = ResidualConnection(core, residual_layer_norm, hidden_size) if residual else core
        if residual and residual_layer_norm and device is not None:
            self.mixer.layer_norm.to(device)
        self.feed_forward = FeedForward(hidden_size=hidden_size, bottleneck_size=bottleneck_size,
                                        nonlinearity=nonlinearity, residual=residual,
                                        residual_layer_norm=residual_layer_norm, **kw)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.feed_forward(self.mixer(x))


def _projections(hidden_size, input_projection, input_projection_size, output_projection, output_projection_size,
                 kw):
    if input_projection and input_projection_size is None:
        raise ValueError("input_projection_size must be specified when input_projection is True.")
    if output_projection and output_projection_size is None:
        raise ValueError("output_projection_size must be specified when output_projection is True.")
    return (B200Linear(input_projection_size, hidden_size, **kw) if input_projection else None,
            B200Linear(output_projection_size, hidden_size, **kw) if output_projection else None)


class MLPMixerLayerd(nn.Module):
    def __init__(self, hidden_size: int, input_projection: bool = False, input_projection_size: int = None,
                 output_projection: bool = False, output_projection_size: int = None, num_layerd: int = 1,
                 num_internal_layer: int = 1, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        self.input_projection, self.output_projection = _projections(
            hidden_size, input_projection, input_projection_size, output_projection, output_projection_size, kw)
        self.mixer = nn.Sequential(OrderedDict(
            (f"block[{i + 1}]", MLPMixerBlock(hidden_size=hidden_size, num_layer=num_internal_layer,
                                             nonlinearity=nonlinearity, residual=residual,
                                             residual_layer_norm=residual_layer_norm,
                                             bottleneck_size=bottleneck_size, **kw))
            for i in range(num_layerd)))

    def forward(self, x: torch.Tensor, hx=None, other=(None,)):
        if self.input_projection is not None:
            x = self.input_projection(x)
        x = self.mixer(x)
        if self.output_projection is not None:
            x = self.output_projection(x)
        return (x, hx, other)


# ------------------------------------------------------------------------------------------------------
# multi-head attention mixer (the cross-modal integrators of lstmformer)
# ------------------------------------------------------------------------------------------------------
class MHAforSequentail(nn.Module):
    """``nn.MultiheadAttention`` taking its 8 positional arguments as ONE tuple (so it can sit in a chain);
    ``need_weights`` / ``average_attn_weights`` forced off.  Returns the attention's ``(out, None)`` tuple, or
    the activated tensor when a nonlinearity is configured (for_sequential.py:40-50).  [sic] class name."""

    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0, bias: bool = True,
                 add_bias_kv: bool = False, add_zero_attn: bool = False, kdim: int = None, vdim: int = None,
                 batch_first: bool = False, nonlinearity=None, device: torch.device = None,
                 dtype: torch.dtype = None):
        super().__init__()
        self.mha = B200MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, dropout=dropout, bias=bias,
                                          add_bias_kv=add_bias_kv, add_zero_attn=add_zero_attn, kdim=kdim,
                                          vdim=vdim, batch_first=batch_first, device=device, dtype=dtype)
        act = set_nonlinearity(nonlinearity)
        self.nonlinearity = None if act is None else act()

    def forward(self, x):
        q, k, v, key_padding_mask, _, attn_mask, _, is_causal = x
        out = self.mha(q, k, v, key_padding_mask, False, attn_mask, False, is_causal)
        return out if self.nonlinearity is None else self.nonlinearity(out[0])


class MHAMixer(nn.Module):
    def __init__(self, input_size: int, num_heads: int, num_layers: int = 1, dropout: float = 0.0,
                 add_bias_kv: bool = False, add_zero_attn: bool = False, kdim: int = None, vdim: int = None,
                 batch_first: bool = False, nonlinearity=None, bias: bool = True, device: torch.device = None,
                 dtype: torch.dtype = None):
        super().__init__()
        if num_layers < 1:
            raise ValueError("num_layers must be greater than 0.")
        self.mixer = nn.ModuleList(
            MHAforSequentail(embed_dim=input_size, num_heads=num_heads, dropout=dropout, add_bias_kv=add_bias_kv,
                             add_zero_attn=add_zero_attn, kdim=kdim, vdim=vdim, batch_first=batch_first,
                             nonlinearity=nonlinearity, bias=bias, device=DEVICE if device is None else device,
                             dtype=dtype)
            for _ in range(num_layers))

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_mask: torch.Tensor = None):
        x = (q, k, v, None, False, attn_mask, False, False)
        for layer in self.mixer:  # like the reference, only well-formed for one internal layer
            x = layer(x)
        return x[0]


class MHAMixerBlock(nn.Module):
    """ResidualConnection(MHAMixer) + FeedForward.  The per-block state is the (key, value) pair; in eval mode a
    supplied state is prepended as a KV cache (dead code under Q3, kept: mixer_block.py:591-595, including its
    slice over the BATCH axis)."""

    def __init__(self, hidden_size: int, num_layers: int = 1, num_heads: int = 1, dropout: float = 0.0,
                 batch_first: bool = True, add_bias_kv: bool = False, add_zero_attn: bool = False,
                 kdim: int = None, vdim: int = None, max_context_len: int = 125, nonlinearity=None,
                 residual: bool = False, residual_layer_norm: bool = False, bottleneck_size: int = None,
                 bias: bool = True, device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        core = MHAMixer(input_size=hidden_size, num_heads=num_heads, num_layers=num_layers, dropout=dropout,
                        add_bias_kv=add_bias_kv, add_zero_attn=add_zero_attn, kdim=kdim, vdim=vdim,
                        batch_first=batch_first, nonlinearity=nonlinearity, **kw)
        self.mixer = ResidualConnection(core, residual_layer_norm, hidden_size) if residual else core
        if residual and residual_layer_norm and device is not None:
            self.mixer.layer_norm.to(device)
        self.feed_forward = FeedForward(hidden_size=hidden_size, bottleneck_size=bottleneck_size,
                                        nonlinearity=nonlinearity, residual=residual,
                                        residual_layer_norm=residual_layer_norm, **kw)
        self.max_context_len = max_context_len

    def forward(self, query, key: torch.Tensor = None, value: torch.Tensor = None,
                attn_mask: torch.Tensor = None, hx=None, prev_hx=None):
        if isinstance(query, (tuple, list)):
            query, key, value, attn_mask, hx, prev_hx = query
        elif not isinstance(query, torch.Tensor):
            raise TypeError(f"query must be torch.Tensor or tuple or list, but got {type(query)}.")
        state, hx, prev_hx = split_state(hx, prev_hx)
        if isinstance(state, (tuple, list)) and not self.training:
            key = torch.cat([state[0], key], dim=1)[-self.max_context_len:]
            value = torch.cat([state[1], value], dim=1)[-self.max_context_len:]
        x = self.feed_forward(self.mixer(query, key, value, attn_mask))
        prev_hx.append((key, value))
        return (x, key, value, attn_mask, hx, prev_hx)


class MHAMixerLayerd(nn.Module):
    def __init__(self, hidden_size: int, input_projection: bool = False, input_projection_size: int = None,
                 self_attention: bool = False, output_projection: bool = False,
                 output_projection_size: int = None, num_heads: int = 1, dropout: float = 0.0,
                 batch_first: bool = True, add_bias_kv: bool = False, add_zero_attn: bool = False,
                 kdim: int = None, vdim: int = None, max_context_len: int = 125, num_layerd: int = 1,
                 num_internal_layer: int = 1, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        self.input_projection, self.output_projection = _projections(
            hidden_size, input_projection, input_projection_size, output_projection, output_projection_size, kw)
        self.self_attention = self_attention
        self.mixer = nn.ModuleList(
            MHAMixerBlock(hidden_size=hidden_size, num_layers=num_internal_layer, num_heads=num_heads,
                          dropout=dropout, batch_first=batch_first, add_bias_kv=add_bias_kv,
                          add_zero_attn=add_zero_attn, kdim=kdim, vdim=vdim, nonlinearity=nonlinearity,
                          residual=residual, residual_layer_norm=residual_layer_norm,
                          bottleneck_size=bottleneck_size, **kw)  # max_context_len is NOT forwarded (:906-924)
            for _ in range(num_layerd))
        self.max_context_len = max_context_len

    def forward(self, x: torch.Tensor, hx=None, key=None, value: torch.Tensor = None,
                attn_mask: torch.Tensor = None):
        if isinstance(key, (tuple, list)):
            key, value, attn_mask = key
        elif not isinstance(key, torch.Tensor):
            raise TypeError(f"key must be torch.Tensor or tuple or list, but got {type(key)}.")
        query = x if self.input_projection is None else self.input_projection(x)
        if self.self_attention:
            key = value = query
        if key is None or value is None:
            raise ValueError("key and value must be specified when self_attention is False.")
        collected = None
        for block in self.mixer:
            query, _, _, _, hx, collected = block(query, key, value, attn_mask, hx, collected)
        if self.output_projection is not None:
            query = self.output_projection(query)
        return (query, hx, (key, value, attn_mask))  # hx: remaining INPUT states (Q3)


# ------------------------------------------------------------------------------------------------------
# GRU mixer (emb_mixers: "gru", mr_gen/model/lstmformer/config_gru.yaml)
# ------------------------------------------------------------------------------------------------------
class GRUMixer(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, batch_first: bool = True,
                 dropout: float = 0.0, bidirectional: bool = False, bias: bool = True, device: torch.device = None,
                 dtype: torch.dtype = None):
        super().__init__()
        if num_layers < 1:
            raise ValueError("num_layers must be greater than 0.")
        if bidirectional:
            if hidden_size % 2 != 0:
                raise ValueError("hidden_size must be even number when bidirectional is True.")
            hidden_size //= 2
        self.mixer = B200GRU(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                             batch_first=batch_first, dropout=dropout, bidirectional=bidirectional, bias=bias,
                             device=DEVICE if device is None else device, dtype=dtype)

    def forward(self, x: torch.Tensor, hn: Optional[torch.Tensor]):
        return self.mixer(x, hn)


class GRUMixerBlock(nn.Module):
    """ResidualConnection(GRUMixer) + FeedForward.  Like the reference (:412-415) a supplied state list is sliced
    with ``hx[:1]`` — a one-element LIST, which ``nn.GRU`` / ``B200GRU`` reject; under Q3 no state is ever supplied."""

    def __init__(self, hidden_size: int, num_layers: int = 1, dropout: float = 0.0, batch_first: bool = True,
                 bidirectional: bool = False, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        core = GRUMixer(input_size=hidden_size, hidden_size=hidden_size, num_layers=num_layers,
                        batch_first=batch_first, dropout=dropout, bidirectional=bidirectional, **kw)
        self.mixer = ResidualConnection(core, residual_layer_norm, hidden_size) if residual else core
        if residual and residual_layer_norm and device is not None:
            self.mixer.layer_norm.to(device)
        self.feed_forward = FeedForward(hidden_size=hidden_size, bottleneck_size=bottleneck_size,
                                        nonlinearity=nonlinearity, residual=residual,
                                        residual_layer_norm=residual_layer_norm, **kw)

    def forward(self, x, hx=None, prev_hx=None):
        if isinstance(x, tuple):
            x, hx, prev_hx = x
        elif not isinstance(x, torch.Tensor):
            raise TypeError(f"x must be torch.Tensor or tuple or list, but got {type(x)}.")
        state, hx = (None, None) if hx is None else (hx[:1], hx[1:])
        state = None if state == [] else state
        hx = None if hx == [] else hx
        prev_hx = [] if prev_hx is None else prev_hx
        y, state = self.mixer(x, state)
        y = self.feed_forward(y)
        prev_hx.append(state)
        return (y, hx, prev_hx)


class GRUMixerLayerd(nn.Module):
    def __init__(self, hidden_size: int, input_projection: bool = False, input_projection_size: int = None,
                 output_projection: bool = False, output_projection_size: int = None, num_layerd: int = 1,
                 num_internal_layer: int = 1, dropout: float = 0.0, batch_first: bool = True,
                 bidirectional: bool = False, nonlinearity=None, residual: bool = False,
                 residual_layer_norm: bool = False, bottleneck_size: int = None, bias: bool = True,
                 device: torch.device = None, dtype: torch.dtype = None):
        super().__init__()
        kw = {"bias": bias, "device": device, "dtype": dtype}
        self.input_projection, self.output_projection = _projections(
            hidden_size, input_projection, input_projection_size, output_projection, output_projection_size, kw)
        self.mixer = nn.ModuleList(
            GRUMixerBlock(hidden_size=hidden_size, num_layers=num_internal_layer, dropout=dropout,
                          batch_first=batch_first, bidirectional=bidirectional, nonlinearity=nonlinearity,
                          residual=residual, residual_layer_norm=residual_layer_norm,
                          bottleneck_size=bottleneck_size, **kw)
            for _ in range(num_layerd))

    def forward(self, x: torch.Tensor, hx=None, other=(None,)):
        if self.input_projection is not None:
            x = self.input_projection(x)
        collected = None
        for block in self.mixer:
            x, hx, collected = block(x, hx, collected)
        if self.output_projection is not None:
            x = self.output_projection(x)
        return (x, hx, other)  # hx = what is left of the INPUT list (Q3)


class MixerBlockFactory:
    _kinds = {"mlp": MLPMixerBlock, "gru": GRUMixerBlock, "lstm": LSTMMixerBlock, "mha": MHAMixerBlock}

    def build(self, mixer_type: str, configs: dict):
        if mixer_type not in self._kinds:
            raise ValueError(f"mixer_type must be in {self._kinds.keys()}.")
        return self._kinds[mixer_type](**configs)


class MixerLayerdFactory:
    _kinds = {"mlp": MLPMixerLayerd, "gru": GRUMixerLayerd, "lstm": LSTMMixerLayerd, "mha": MHAMixerLayerd}

    def build(self, mixer_type: str, configs: dict):
        if mixer_type not in self._kinds:
            raise ValueError(f"mixer_type must be in {self._kinds.keys()}.")
        return self._kinds[mixer_type](**configs)
