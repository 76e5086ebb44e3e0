"""B200-native (sm_100a) LSTM hot path of MultimodalReactionGeneration.

Public surface:
  * ``B200LSTM``            — drop-in for ``torch.nn.LSTM`` at the reference's three seams
  * ``B200GRU``             — drop-in for ``torch.nn.GRU`` at lstmformer's GRU mixer
  * ``mr_gen``              — host-side mirror of the reference's module tree (same class names,
                              constructor / forward signatures, state_dict keys)
  * ``_cabi``               — ctypes binding of the C-ABI in ``include/mrg_lstm.h``

There is no CPU fallback: calling the LSTM path without the compiled extension or without a
compute-capability-10.x device raises.
"""
from .lstm import B200LSTM, lstm_layer, set_precision  # noqa: F401
from .linear import B200Linear  # noqa: F401
from .attention import B200MultiheadAttention  # noqa: F401
from .gru import B200GRU  # noqa: F401

__all__ = ["B200LSTM", "B200GRU", "B200Linear", "B200MultiheadAttention", "lstm_layer", "set_precision"]
