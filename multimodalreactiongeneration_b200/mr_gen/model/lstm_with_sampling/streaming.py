"""Frame-by-frame batched generation with device-resident state (BASELINE.json configs[4]: 1024 concurrent
dyads at 30 fps).  This is the batched, persistent form of the reference's ``generate_one_step``
(mr_gen/model/lstm_with_sampling/lstm_with_sample.py:410-433) as it is driven by
``visualize_metaformer.py:116-127``: the sampler LSTM state is carried (it is the only carried state, Q2),
the predictor blocks restart from zero each frame, the previous pose is the model's own last prediction
(free running) unless the caller supplies one.

All per-frame work — acoustic projection, sampler LSTM step(s) with carried (h, c), feature projection,
two stateless LSTM blocks + LayerNorm, bottleneck FFN — is captured ONCE in a CUDA graph over static
buffers; a frame is: copy inputs in, replay, read the pose out."""
from __future__ import annotations

from typing import Optional

import torch


class StreamingGenerator:
    def __init__(self, model, batch_size: int, use_cuda_graph: bool = True):
        self.model = model.eval()
        self.B = batch_size
        dev = model.device
        m = model.model
        self.ratio = model.ratio
        self.A = (m.nmels + 1) * (m.delta_order + 1)
        self.P = (m.use_centroid + m.use_angle) * 3 * (m.delta_order + 1)
        L, Hs = m.sampler_num_layers, m.sampler_hidden_size
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.audio = z(batch_size, self.ratio, self.A)       # static input buffers
        self.partner = z(batch_size, 1, self.P)
        self.prev = z(batch_size, 1, self.P)                  # fed-back pose (state)
        self.h, self.c = z(L, batch_size, Hs), z(L, batch_size, Hs)
        self.out = z(batch_size, 1, self.P)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._use_graph = use_cuda_graph

    @torch.no_grad()
    def reset(self, lead_audio=None, lead_partner=None, lead_self=None) -> None:
        """Zero state, or warm the sampler state up on a leading segment (reference: ``warmup_model``)."""
        self.h.zero_(); self.c.zero_(); self.prev.zero_()
        if lead_audio is not None and lead_audio.shape[1] > 0:
            a = self.model.acoustic_projection(lead_audio.to(self.audio.device))
            _, (h, c) = self.model.sampling_lstm(a, None)
            self.h.copy_(h); self.c.copy_(c)
        if lead_self is not None and lead_self.shape[1] > 0:
            self.prev.copy_(lead_self[:, -1:].to(self.prev.device))

    @torch.no_grad()
    def _frame(self) -> None:
        mdl = self.model
        a = mdl.acoustic_projection(self.audio)
        s, (h, c) = mdl.sampling_lstm(a, (self.h, self.c))
        feat = mdl.feature_projection(torch.cat([s, self.partner, self.prev], dim=-1))
        hid, _ = mdl.layerd_lstm(feat, None)
        y = mdl.feed_forward(hid)
        self.h.copy_(h); self.c.copy_(c)
        self.out.copy_(y)
        self.prev.copy_(y)

    def _capture(self) -> None:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved = [t.clone() for t in (self.h, self.c, self.prev)]
        with torch.cuda.stream(side):
            for _ in range(3):
                self._frame()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._frame()
        for dst, src in zip((self.h, self.c, self.prev), saved):  # warm-up frames must not advance the state
            dst.copy_(src)

    @torch.no_grad()
    def step(self, audio: torch.Tensor, partner: torch.Tensor, previous: Optional[torch.Tensor] = None):
        """audio [B, ratio, A], partner [B, P] or [B, 1, P] (host or device) -> pose [B, P] (device view)."""
        self.audio.copy_(audio.view_as(self.audio), non_blocking=True)
        self.partner.copy_(partner.reshape(self.B, 1, self.P), non_blocking=True)
        if previous is not None:
            self.prev.copy_(previous.reshape(self.B, 1, self.P), non_blocking=True)
        if self._use_graph:
            if self._graph is None:
                self._capture()
            self._graph.replay()
        else:
            self._frame()
        return self.out[:, 0]
