"""Stand-in for the slice of ``pytorch_lightning.LightningModule`` the reference's models touch
(``current_epoch``, ``device``, ``log``, ``log_dict``); the training loop lives in mr_gen/tainer."""
import torch
from torch import nn


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.current_epoch = 0
        self.logged = {}

    @property
    def device(self) -> torch.device:
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, name, value, **_):
        # detached: a logged loss must not keep the step's autograd graph alive
        self.logged[name] = value.detach() if torch.is_tensor(value) else value

    def log_dict(self, values, **_):
        if isinstance(values, dict):
            for k, v in values.items():
                self.log(k, v)
