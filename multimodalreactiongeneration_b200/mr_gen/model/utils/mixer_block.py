"""Mirror of the LSTM half of mr_gen/model/utils/mixer_block.py — the token mixers of lstmformer:
``split_state`` :21-30, ``FeedForward`` :37-87, ``LSTMMixer`` :211-252, ``LSTMMixerBlock`` :431-507,
``LSTMMixerLayerd`` :762-843, with ``nn.LSTM`` replaced by ``B200LSTM``.

Quirk Q3 is kept: ``LSTMMixerLayerd.forward`` returns the REMAINING input state
list (``hx`` after ``split_state``), not the collected new states, so callers only ever see ``None``.
The MLP / GRU / MHA mixers of the same reference file are out of the LSTM hot path."""
from collections import OrderedDict
from typing import Any, List, Optional, Tuple, Union

import torch
from torch import nn

from ....lstm import B200LSTM
from ....linear import B200Linear
from .residual_connection import ResidualConnection

LSTMStateType = Tuple[torch.Tensor, torch.Tensor]
DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")

_ACTIVATIONS = {"relu": nn.ReLU, "swish": nn.SiLU, "tanh": nn.Tanh}  # nonlinearity.py:6-16


def set_nonlinearity(name: str):
    if name in _ACTIVATIONS:
        return _ACTIVATIONS[name]
    raise ValueError(f"nonlinearity must be in {sorted(_ACTIVATIONS)} or none")


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
