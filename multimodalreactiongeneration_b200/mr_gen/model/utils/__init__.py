from .residual_connection import ResidualConnection  # noqa: F401
from .lstm_sampler import LSTMSampler  # noqa: F401
