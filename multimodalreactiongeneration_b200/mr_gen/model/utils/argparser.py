"""Mirror of mr_gen/model/utils/argparser.py: pick, out of one shared hyper-parameter set, the keyword
arguments each mixer stack's constructor takes (``mixer_layerd_argments_select`` :324-435,
``feedforward_block_argments`` :300-321, the per-type ``*_mixer_layerd_argments`` :132-297).  [sic] spelling of
the function names is the reference's public API."""

_COMMON = ("hidden_size", "input_projection", "input_projection_size", "output_projection",
           "output_projection_size", "num_layerd", "num_internal_layer", "nonlinearity", "residual",
           "residual_layer_norm", "bottleneck_size", "bias", "device", "dtype")
_LAYERD_KEYS = {
    "mlp": _COMMON,
    "gru": _COMMON + ("dropout", "batch_first", "bidirectional"),
    "lstm": _COMMON + ("dropout", "batch_first", "bidirectional", "proj_size"),
    "mha": _COMMON + ("self_attention", "num_heads", "dropout", "batch_first", "add_bias_kv", "add_zero_attn",
                      "kdim", "vdim", "max_context_len"),
}
_DEFAULTS = dict(
    input_projection=False, input_projection_size=None, self_attention=False, output_projection=False,
    output_projection_size=None, num_heads=1, dropout=0.0, batch_first=True, bidirectional=False, proj_size=0,
    add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None, max_context_len=125, num_layerd=1,
    num_internal_layer=1, nonlinearity=None, residual=False, residual_layer_norm=False, bottleneck_size=None,
    bias=True, device=None, dtype=None)


def mixer_layerd_argments_select(mixer_type: str, hidden_size: int, **hyper) -> dict:
    """kwargs of ``{MLP,GRU,LSTM,MHA}MixerLayerd`` taken from the shared set; like the reference an unknown
    ``mixer_type`` yields ``None`` and an unknown hyper-parameter is a ``TypeError``."""
    unknown = set(hyper) - set(_DEFAULTS)
    if unknown:
        raise TypeError(f"mixer_layerd_argments_select() got an unexpected keyword argument '{sorted(unknown)[0]}'")
    if mixer_type not in _LAYERD_KEYS:
        return None
    merged = {**_DEFAULTS, **hyper, "hidden_size": hidden_size}
    return {k: merged[k] for k in _LAYERD_KEYS[mixer_type]}


def mlp_mixer_layerd_argments(hidden_size: int, **hyper) -> dict:
    return mixer_layerd_argments_select("mlp", hidden_size, **hyper)


def gru_mixer_layerd_argments(hidden_size: int, **hyper) -> dict:
    return mixer_layerd_argments_select("gru", hidden_size, **hyper)


def lstm_mixer_layerd_argments(hidden_size: int, **hyper) -> dict:
    return mixer_layerd_argments_select("lstm", hidden_size, **hyper)


def mha_mixer_layerd_argments(hidden_size: int, **hyper) -> dict:
    return mixer_layerd_argments_select("mha", hidden_size, **hyper)


def feedforward_block_argments(hidden_size: int, bottleneck_size: int = None, output_size: int = None,
                               nonlinearity=None, residual: bool = False, residual_layer_norm: bool = False,
                               bias: bool = True, device=None, dtype=None) -> dict:
    return dict(hidden_size=hidden_size, bottleneck_size=bottleneck_size, output_size=output_size,
                nonlinearity=nonlinearity, residual=residual, residual_layer_norm=residual_layer_norm, bias=bias,
                device=device, dtype=dtype)
