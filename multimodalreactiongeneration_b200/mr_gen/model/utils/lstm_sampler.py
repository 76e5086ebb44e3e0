"""Mirror of mr_gen/model/utils/lstm_sampler.py:6-34: an LSTM over acoustic frames whose output is
kept at every ``decline_rate``-th step (100 fps acoustic -> 12.5 fps motion in the reference)."""
from typing import Optional, Tuple

import torch
from torch import nn

from ....lstm import B200LSTM


class LSTMSampler(nn.Module):
    def __init__(self, hiddn_size: int, num_layers: int, dropout: float, decline_rate: int,
                 bidirectional=False):
        super().__init__()
        # attribute name fixes the checkpoint keys: sampling_lstm.sampler.weight_ih_l0 ...
        self.sampler = B200LSTM(input_size=hiddn_size, hidden_size=hiddn_size, num_layers=num_layers,
                                dropout=dropout, bidirectional=bidirectional, batch_first=True)
        self.decline_rate = decline_rate

    def forward(self, x: torch.Tensor, hx: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        h, hx = self.sampler(x, hx)
        r = self.decline_rate
        return h[:, r - 1::r, :].contiguous(), hx
