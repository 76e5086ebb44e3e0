"""Host-side mirror of the reference's ``mr_gen`` package, restricted to the LSTM hot path.

Same class names, constructor / ``forward`` signatures, attribute names (hence ``state_dict`` /
``.ckpt`` keys) and error behaviour as the reference; every ``torch.nn.LSTM``
is a ``B200LSTM``.  pytorch_lightning / omegaconf / hydra / torchmetrics are not required.
"""
