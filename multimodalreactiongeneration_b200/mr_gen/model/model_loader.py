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
    model.load_state_dict(torch.load(model_path, map_location="cpu")["state_dict"])
    return model
