"""Mirror of mr_gen/model/model_loader.py:13-26: build a model by name and load ``ckpt["state_dict"]``."""
import torch

from .lstm_with_sampling.lstm_with_sample import LSTMwithSample
from .simple_lstm.simple_lstm import SimpleLSTM

_MODELS = {"simple_lstm": SimpleLSTM, "lstm_with_sampling": LSTMwithSample}


def load_model(model_type: str, model_path: str, cfg):
    if model_type == "lstmformer":
        from .lstmformer.lstmformer import Metaformer
        cls = Metaformer
    elif model_type in _MODELS:
        cls = _MODELS[model_type]
    else:
        raise ValueError(f"invalid model type: {model_type}")
    model = cls(cfg.model, cfg.optim, cfg.metrics)
    # the reference's checkpoints are Lightning .ckpt files (pickled non-tensor objects next to the state_dict): they are
    # the user's own training output, loaded as the reference does (torch >= 2.6 defaults to weights_only=True)
    model.load_state_dict(torch.load(model_path, map_location="cpu", weights_only=False)["state_dict"])
    return model
