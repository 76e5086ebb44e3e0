"""Mirror of mr_gen/model/utils/lstm_block.py (LSTMModule :9-46, LSTMBlock :49-107,
LSTMLayerd :110-169) with ``nn.LSTM`` replaced by ``B200LSTM``.

Quirks kept on purpose: the operator-precedence size check of LSTMBlock (Q8),
``affine_hidden_size`` forced to the LSTM width without mixing (Q8), and ``LSTMLayerd.forward``
returning the *input* ``hxs`` (Q2) — so callers never see carried predictor state."""
from collections import OrderedDict
from typing import List, Optional, Tuple

import torch
from torch import nn

from ....lstm import B200LSTM
from ....linear import B200Linear
from .residual_connection import ResidualConnection

State = Tuple[torch.Tensor, torch.Tensor]


class LSTMModule(nn.Module):
    def __init__(self, input_size=256, hidden_size=128, num_layers=1, output_size=256, dropout=0.0,
                 bidirectional=True, use_mixing=True):
        super().__init__()
        width = hidden_size * (2 if bidirectional else 1)
        if not use_mixing and width != output_size:
            raise ValueError("lstm_out_size must be equal to output_size when use_mixing is False.")
        self.lstm_module = B200LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                                    dropout=dropout, batch_first=True, bidirectional=bidirectional)
        self.mixer = B200Linear(width, output_size) if use_mixing else None

    def forward(self, input_tensor, hx=None) -> Tuple[torch.Tensor, State]:
        hs, hx = self.lstm_module(input_tensor, hx)
        return (hs if self.mixer is None else self.mixer(hs)), hx


class LSTMBlock(nn.Module):
    def __init__(self, input_size=256, hidden_size=128, lstm_out_size=256, num_layers=1, bottleneck_size=64,
                 output_size=256, dropout=0.0, bidirectional=True, use_layer_norm=True, use_relu=True,
                 use_mixing=False, use_residual=True, use_feed_forward=True) -> None:
        super().__init__()
        # same precedence as the reference: (use_residual and a) or b
        if (use_residual and input_size != lstm_out_size) or lstm_out_size != output_size:
            raise ValueError("input_size must be equal to lstm_out_size and output_size when use_residuals.")
        self.use_feed_forward = use_feed_forward
        core = LSTMModule(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                          output_size=lstm_out_size, dropout=dropout, bidirectional=bidirectional,
                          use_mixing=use_mixing)
        self.lstm_module = ResidualConnection(core, use_layer_norm, lstm_out_size, dropout) if use_residual else core
        if use_feed_forward:
            parts = [("input", B200Linear(lstm_out_size, bottleneck_size))]
            if use_relu:
                parts.append(("relu", nn.ReLU()))
            parts.append(("mapping", B200Linear(bottleneck_size, output_size)))
            ffn = nn.Sequential(OrderedDict(parts))
            self.feed_forward_module = (ResidualConnection(ffn, use_layer_norm, output_size, dropout)
                                        if use_residual else ffn)

    def forward(self, input_tenor: torch.Tensor, hx=None) -> Tuple[torch.Tensor, State]:
        y, hx = self.lstm_module(input_tenor, hx)
        if self.use_feed_forward:
            y = self.feed_forward_module(y)
        return y, hx

    def dropout_inactive(self) -> bool:
        return not self.training or all(m.p == 0.0 for m in self.modules() if isinstance(m, nn.Dropout))

    def forward_last_step(self, input_tenor: torch.Tensor, hx=None) -> torch.Tensor:
        """``self.forward(x, hx)[0][..., -1:, :]``: the recurrence runs over the whole sequence, the per-token tail (mixer,
        residual + LayerNorm, bottleneck FFN) only on the frame that is kept.  Exact while no dropout mask is drawn (a mask
        of the sliced shape would be a different draw): callers check ``dropout_inactive()``."""
        core = self.lstm_module
        wrapped = isinstance(core, ResidualConnection)
        mod = core.module if wrapped else core
        hs, (h_n, _) = mod.lstm_module(input_tenor, hx)
        lstm = mod.lstm_module
        if not lstm.bidirectional and lstm.batch_first and input_tenor.dim() == 3:
            # one direction: the last output frame IS the final hidden state of the top layer — taking it from h_n keeps
            # the gradient out of a dense [B, T, H] zero tensor (it enters the BPTT kernel as dh_n; no dy stream at all)
            y = h_n[-1].unsqueeze(1)
        else:
            y = hs[..., -1:, :]
        if mod.mixer is not None:
            y = mod.mixer(y)
        if wrapped:
            y = core.finish(y, input_tenor[..., -1:, :])
        if self.use_feed_forward:
            y = self.feed_forward_module(y)
        return y


class LSTMLayerd(nn.Module):
    def __init__(self, input_size=256, lstm_hidden_size=128, affine_hidden_size=256, bottleneck_size=64,
                 num_layers=2, num_layers_per_block=1, output_size=256, dropout=0.0, bidirectional=True,
                 use_layer_norm=True, use_relu=True, use_mixing=False, use_residual=True,
                 use_feed_forward=True):
        super().__init__()
        lstm_width = lstm_hidden_size * (2 if bidirectional else 1)
        inner = affine_hidden_size if use_mixing else lstm_width
        blocks = []
        for i in range(num_layers):
            blocks.append(LSTMBlock(
                input_size=input_size if i == 0 else inner, hidden_size=lstm_hidden_size,
                lstm_out_size=inner, num_layers=num_layers_per_block, bottleneck_size=bottleneck_size,
                output_size=output_size if i == num_layers - 1 else inner, dropout=dropout,
                bidirectional=bidirectional, use_layer_norm=use_layer_norm, use_relu=use_relu,
                use_mixing=use_mixing, use_residual=use_residual, use_feed_forward=use_feed_forward))
        self.lstm_layered = nn.ModuleList(blocks)

    def forward(self, input_tensor: torch.Tensor, hxs: Optional[List[State]] = None):
        x = input_tensor
        for i, block in enumerate(self.lstm_layered):
            x, _ = block(x, None if hxs is None else hxs[i])
        # the reference collects the new states but hands back its argument (lstm_block.py:164-169)
        return x, hxs

    def forward_last_step(self, input_tensor: torch.Tensor, hxs: Optional[List[State]] = None) -> torch.Tensor:
        """``self.forward(x, hxs)[0][..., -1:, :]`` — what the reference's MotionDecoder keeps of its decoder stack
        (``seq_reshape``, simple_lstm.py:168-170) — without running the per-token tail of the LAST block on the T - 1
        frames nobody reads (mixer Linear, residual + LayerNorm, bottleneck FFN and their backward GEMMs over [B, T, F]).
        Per-token operators commute with the slice, so the kept frame and every gradient are the same numbers."""
        blocks = self.lstm_layered
        last = blocks[len(blocks) - 1]
        if not last.dropout_inactive():
            return self.forward(input_tensor, hxs)[0][..., -1:, :]
        x = input_tensor
        for i in range(len(blocks) - 1):
            x, _ = blocks[i](x, None if hxs is None else hxs[i])
        return last.forward_last_step(x, None if hxs is None else hxs[len(blocks) - 1])
