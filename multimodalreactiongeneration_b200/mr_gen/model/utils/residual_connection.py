"""Mirror of mr_gen/model/utils/residual_connection.py:5-37 (``y = Dropout(LN(module(x) + x))``)."""
from torch import nn

from .... import layernorm as _fused_ln


class ResidualConnection(nn.Module):
    """Wraps ``module``; extra tuple outputs (LSTM states) pass through untouched.

    The reference derives from ``pl.LightningModule`` only as a base class; ``nn.Module`` gives the
    same parameters and ``state_dict`` keys (``module.*``, ``layer_norm.*``)."""

    def __init__(self, module: nn.Module, use_layer_norm=True, num_nodes: int = -1, dropout=0.0):
        super().__init__()
        if use_layer_norm and num_nodes == -1:
            raise ValueError("num_nodes must be specified when use_layer_norm is set to True.")
        self.module = module
        self.use_layer_norm = use_layer_norm
        self.layer_norm = nn.LayerNorm(num_nodes) if use_layer_norm else None
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, *args, **kwargs):
        out = self.module(x, *args, **kwargs)
        extras = None
        if isinstance(out, (tuple, list)):
            out, extras = out[0], tuple(out[1:])
        out = self.finish(out, x)
        return out if extras is None else (out, *extras)

    def finish(self, out, x):
        """Dropout(LN(out + x)): everything of ``forward`` behind the wrapped module (per-token)."""
        ln = self.layer_norm
        if ln is not None and ln.elementwise_affine and ln.bias is not None and _fused_ln.supported(out, x):
            # one pass over the [B, T, H] stream: add + LayerNorm fused, strides consumed in place
            out = _fused_ln.residual_layer_norm(out, x, ln.weight, ln.bias, ln.eps)
        else:
            out = out + x
            if ln is not None:
                out = ln(out)
        return self.dropout(out)
