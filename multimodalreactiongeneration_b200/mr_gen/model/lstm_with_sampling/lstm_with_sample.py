"""Mirror of mr_gen/model/lstm_with_sampling/lstm_with_sample.py (LSTMwithSample :59) on the B200 path.

Structure (forward :151-232): acoustic Linear -> LSTMSampler (every ``ratio``-th output) ->
cat[audio | partner motion | own motion] -> Linear -> LSTMLayerd (residual LSTM blocks) -> bottleneck FFN.

Reference behaviour reproduced on purpose: Q2 predictor state never carried
between ``forward`` calls, Q4 one scheduled-sampling decision per time step shared by the batch (drawn
with ``torch.rand`` when no mask is supplied), Q5 one-frame lag of teacher forcing in the rollout,
Q6 gradients flow through fed-back predictions, Q7 loss averaged over padded positions too.

Extensions (BASELINE.json north_star): ``sampling_mask`` may be supplied ([T] or [T, B] bool) — e.g.
from the counter-based Philox generator (``philox_sampling_mask``), bit-exact for a given seed.

Rollout schedules (``model.rollout``):
``"kernel"`` (default): the whole T loop of the predictor — feedback select, feature projection of the fed-back pose,
the zero-state LSTM blocks + LayerNorm, the bottleneck FFN — runs inside ONE persistent cluster kernel per direction
(``multimodalreactiongeneration_b200.rollout``, C-ABI ``mrg_rollout_forward/backward``); the sampler LSTM, which does not
depend on the feedback, runs once over lead + sequence in the recurrent kernel and enters as a time-parallel GEMM.  No
host read, no per-step launch, the launch count does not depend on T or on the sampling rate.  Configurations the
kernel is not built for (mixing Linear inside the blocks, no LayerNorm, more than 2 blocks, ...) take the wavefront
schedule below, which runs on the same library kernels.

``"wavefront"``: device-resident re-scheduling with the per-layer kernels.  Because of Q2 the predictor
restarts from zero state at every step, so step t depends on step t-1 ONLY through the fed-back 6..18-d
pose, and only where the mask says "feed the prediction back".  The rollout is therefore re-scheduled
without changing any arithmetic: the sampler LSTM runs ONCE over lead+sequence (identical recurrence to the
reference's carried-state calls), the audio/partner part of ``feature_projection`` is time-parallel, and the
positions are processed in waves of equal feedback depth (depth[t] = depth[t-1]+1 if mask[t-1] else 0):
wave d needs only wave d-1.  There is no per-step host loop — the number of launches is proportional to
the longest run of fed-back steps, not to T — and autograd carries Q6 through the gathers."""
from collections import OrderedDict
from typing import List, Optional, Tuple

import torch
from torch import nn

from ....linear import B200Linear, _LinearFn
from .... import rollout as _rollout

from ....import _cabi
from ...utils.lightning_shim import LightningModule
from ...utils.metrics import MultiTargetMetrics, gen_target_dict
from ..simple_lstm.simple_lstm import _configure_optimizers
from ..utils import LSTMSampler
from ..utils.lstm_block import LSTMLayerd
from ..utils.values import PADDING_VALUE

InputTypes = Tuple[torch.Tensor, torch.Tensor]


def _linear(x, weight, bias):
    """``x @ weight.T + bias`` with a SLICE of a registered weight, on the library's GEMM (no cuBLAS)."""
    return _LinearFn.apply(x, weight, bias)


def philox_sampling_mask(seed: int, offset: int, prob: float, length: int, batch: int, device,
                         shared: bool = False) -> torch.Tensor:
    """bool [T, B] mask from Philox4x32-10 (counter = (offset+t, b), key = seed), generated on the
    device by ``mrg_philox_mask``; ``shared=True`` repeats one decision per step over the batch."""
    out = torch.empty((length, batch), dtype=torch.uint8, device=device)
    if length * batch:
        with torch.cuda.device(out.device):
            st = _cabi.lib().mrg_philox_mask(seed, offset, float(prob), length, batch, int(shared),
                                             out.data_ptr(), torch.cuda.current_stream(out.device).cuda_stream)
        _cabi.check(st, "mrg_philox_mask")
    return out.bool()


class LSTMwithSample(LightningModule):
    def __init__(self, model, optim, metrics):
        super().__init__()
        self.model, self.optim, self.metrics = model, optim, metrics
        self.max_epochs = model.max_epochs
        self.use_scheduled_sampling = model.use_scheduled_sampling
        if model.loss_type not in ("mse", "mae", "huber", "smoothl1"):
            raise ValueError("invalid loss type")
        self.huber_delta = model.get("huber_delta", 1.0)
        self.smoothl1_beta = model.get("smoothl1_beta", 1.0)

        # acoustic frames per predicted motion frame
        self.ratio = int((model.sampling_rate / model.shift) / model.pred_fps)
        pose = (model.use_centroid + model.use_angle) * 3 * (model.delta_order + 1)
        acoustic_in = (model.nmels + 1) * (model.delta_order + 1)

        self.acoustic_projection = B200Linear(acoustic_in, model.sampler_hidden_size)
        self.sampling_lstm = LSTMSampler(model.sampler_hidden_size, model.sampler_num_layers,
                                         model.sampler_dropout_rate, self.ratio, bidirectional=False)
        self.use_device = "cuda" if torch.cuda.is_available() else "cpu"
        self.feature_projection = B200Linear(2 * pose + model.sampler_hidden_size, model.hidden_size)
        self.layerd_lstm = LSTMLayerd(
            input_size=model.hidden_size, lstm_hidden_size=model.hidden_size,
            affine_hidden_size=model.hidden_size, bottleneck_size=model.bottleneck_size,
            num_layers=model.num_layers, num_layers_per_block=model.num_lstm, output_size=model.hidden_size,
            dropout=model.dropout_rate, bidirectional=False, use_layer_norm=model.use_layer_norm,
            use_mixing=model.use_mixing, use_residual=model.use_residual, use_feed_forward=False)
        head = [("input", B200Linear(model.hidden_size, model.bottleneck_size))]
        if model.use_relu:
            head.append(("relu", nn.ReLU()))
        head.append(("mapping", B200Linear(model.bottleneck_size, pose)))
        self.feed_forward = nn.Sequential(OrderedDict(head))

        ranges = gen_target_dict(metrics)
        self.train_metrics = MultiTargetMetrics(target_range=ranges, prefix="train_")
        self.valid_metrics = MultiTargetMetrics(target_range=ranges, prefix="valid_")
        self.genrt_metrics = MultiTargetMetrics(target_range=ranges, prefix="genrt_")
        self.optimizer = None
        self.lr_scheduler = None
        self.delta_loss_scale = model.get("delta_loss_scale", 1.0)
        self.all_static = model.get("all_static", False)
        self.delta_order = metrics.delta_order
        # counter-based scheduled sampling (off unless configured): seed + running offset
        self.sampling_seed: Optional[int] = model.get("sampling_seed", None)
        self.sampling_offset = 0
        self.sampling_per_sample = model.get("sampling_per_sample", True)
        # "kernel": one persistent kernel per direction; "wavefront": re-scheduled rollout on the per-layer kernels;
        # "stepwise": the reference's Python time loop (cross-check)
        self.rollout = model.get("rollout", "kernel")

    # ------------------------------------------------------------------------------------------
    def forward(self, acoustic_partner: InputTypes, motion_partner: InputTypes, motion_self: InputTypes,
                leading_acoustic_partner: InputTypes, leading_motion_partner: InputTypes,
                leading_motion_self: InputTypes, cell_state=None):
        dev = self.use_device
        a, mp, (ms, ms_len) = acoustic_partner[0].to(dev), motion_partner[0].to(dev), motion_self
        ms = ms.to(dev)
        la, lp, ls = (t[0].to(dev) for t in (leading_acoustic_partner, leading_motion_partner,
                                             leading_motion_self))
        hx_sampler, hxs = (None, None) if cell_state is None else cell_state

        lead_len = lp.shape[1]
        audio = torch.cat([la, a], dim=1)
        partner = torch.cat([lp, mp], dim=1)
        own = torch.cat([ls, ms], dim=1)
        motion_len = partner.shape[1]

        sampled, hx_sampler = self.sampling_lstm(self.acoustic_projection(audio), hx_sampler)
        if not (sampled.shape[1] == partner.shape[1] == own.shape[1]):
            raise RuntimeError(
                f"acoustic: {tuple(audio.shape)} -> {tuple(sampled.shape)}: {la.shape[1]} + {a.shape[1]}\n"
                f"motion_p: {lead_len} + {mp.shape[1]}\n"
                f"motion_s: {ls.shape[1]} + {ms.shape[1]}\n"
                f"ratio: {self.ratio}")
        features = self.feature_projection(torch.cat([sampled, partner, own], dim=-1))
        h, hxs = self.layerd_lstm(features, hxs)
        return self.feed_forward(h), (lead_len, motion_len, ms_len), (hx_sampler, hxs)

    # ------------------------------------------------------------------------------------------
    def lossfun(self):
        m = self.model
        if m.loss_type == "mse":
            return nn.MSELoss(reduction=m.loss_reduction)
        if m.loss_type == "mae":
            return nn.L1Loss(reduction=m.loss_reduction)
        if m.loss_type == "huber":
            return nn.HuberLoss(reduction=m.loss_reduction, delta=self.huber_delta)
        return nn.SmoothL1Loss(reduction=m.loss_reduction, beta=self.smoothl1_beta)

    def configure_optimizers(self):
        return _configure_optimizers(self)

    @staticmethod
    def _mask_padding(y, target):
        keep = (target != PADDING_VALUE).int()
        return y * keep, target * keep

    def training_step(self, batch: List[InputTypes], *args):
        if self.use_scheduled_sampling:
            self.log("scheduled_sampling_rate", self.current_epoch / self.max_epochs, logger=True)
            y, target = self.prediction(batch, use_scheduled_sampling=True)
        else:
            y, (lead_len, _, _), _ = self.forward(*batch[:-1])
            target = batch[-1][0].to(y.device)
            y = y[:, lead_len:]
        y, target = self._mask_padding(y, target)
        scaler = torch.ones_like(y)
        scaler[:, :, y.shape[2] // (self.delta_order + 1):] = float(self.delta_loss_scale) ** 0.5
        loss = self.lossfun()(y * scaler, target * scaler)
        self.log("train_loss", loss, prog_bar=True, logger=True)
        self.log_dict(self.train_metrics(y * scaler, target * scaler), logger=True, on_epoch=True, on_step=True)
        return {"loss": loss}

    def validation_step(self, batch: List[InputTypes], *args):
        y, (lead_len, _, _), _ = self.forward(*batch[:-1])
        y, target = self._mask_padding(y[:, lead_len:], batch[-1][0].to(y.device))
        loss = self.lossfun()(y, target)
        self.log("val_loss", loss, prog_bar=True, logger=True)
        self.log_dict(self.valid_metrics(y, target), logger=True, on_epoch=True, on_step=True)
        return {"loss": loss, "gen_loss": self.generation_step(batch)["loss"]}

    def generation_step(self, batch: List[InputTypes]):
        pred, target = self.prediction(batch)
        pred, target = self._mask_padding(pred, target)
        loss = self.lossfun()(pred, target)
        self.log("genrt_loss", loss, prog_bar=False, logger=True)
        self.log_dict(self.genrt_metrics(pred, target), logger=True, on_epoch=True, on_step=True)
        return {"loss": loss}

    # ------------------------------------------------------------------------------------------
    # autoregressive rollout (reference :339-463)
    # ------------------------------------------------------------------------------------------
    def prediction(self, batch: List[InputTypes], use_scheduled_sampling: bool = False,
                   full_generation: bool = False, sampling_mask: Optional[torch.Tensor] = None):
        if self.rollout == "kernel" and self._kernel_rollout_layers() is not None:
            return self._prediction_kernel(batch, use_scheduled_sampling, full_generation, sampling_mask)
        if self.rollout in ("kernel", "wavefront"):
            return self._prediction_wavefront(batch, use_scheduled_sampling, full_generation, sampling_mask)
        formed, dummy, length = self.batch_forming(batch)
        target = batch[-1][0].to(self.device)
        state = self.warmup_model(dummy, batch)
        pred = self.head_motion_generation(formed, dummy, length, state, use_scheduled_sampling,
                                           full_generation, sampling_mask)
        return pred, target

    def _resolve_mask(self, length, batch_size, use_scheduled_sampling, full_generation, sampling_mask):
        if sampling_mask is None:
            if use_scheduled_sampling:
                sampling_mask = self.draw_sampling_mask(length, batch_size)
            else:
                sampling_mask = torch.full((length,), bool(full_generation), dtype=torch.bool)
        return sampling_mask

    def _kernel_rollout_layers(self):
        """[(w_ih, w_hh, b_ih, b_hh, ln_weight, ln_bias)] of the predictor blocks when the persistent rollout kernel is
        built for this configuration (residual + LayerNorm blocks of one uni-directional LSTM layer, no mixing
        Linear, no block FFN, no active dropout), else None."""
        from ..utils.residual_connection import ResidualConnection
        layers = []
        for block in self.layerd_lstm.lstm_layered:
            rc = block.lstm_module
            if block.use_feed_forward or not isinstance(rc, ResidualConnection) or rc.layer_norm is None:
                return None
            ln, core = rc.layer_norm, rc.module
            lstm = core.lstm_module
            if (core.mixer is not None or lstm.num_layers != 1 or lstm.bidirectional or not ln.elementwise_affine
                    or ln.bias is None or lstm.input_size != lstm.hidden_size
                    or (self.training and (rc.dropout.p > 0 or lstm.dropout > 0))):
                return None
            layers.append((lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0 if lstm.bias else None,
                           lstm.bias_hh_l0 if lstm.bias else None, ln.weight, ln.bias))
        ff = self.feed_forward
        H, P, FB = self.feature_projection.out_features, ff.mapping.out_features, ff.input.out_features
        if not layers or len({float(b.lstm_module.layer_norm.eps) for b in self.layerd_lstm.lstm_layered}) != 1:
            return None
        if not self.feature_projection.weight.is_cuda or not _rollout.supported(H, len(layers), P, FB):
            return None
        return layers

    def _prediction_kernel(self, batch, use_scheduled_sampling, full_generation, sampling_mask):
        dev = self.device
        (a, _), (mp, _), (ms, _), (la, _) = batch[:4]
        a, mp, ms, la = a.to(dev), mp.to(dev), ms.to(dev), la.to(dev)
        target = batch[-1][0].to(dev)
        B, T, P = mp.shape
        if T == 0:
            return ms.new_zeros((B, 0, P)), target
        if sampling_mask is None and not use_scheduled_sampling:
            # step-wise teacher forcing (never feed back) / free running (always): no draw, nothing from the host
            mask = torch.ones((T, B), dtype=torch.uint8, device=dev) if full_generation else None
        else:
            mask = self._resolve_mask(T, B, use_scheduled_sampling, full_generation, sampling_mask).to(dev)
            if mask.dim() == 1:
                mask = mask.view(T, 1).expand(T, B)
            mask = mask.to(torch.uint8).contiguous()
        layers = self._kernel_rollout_layers()
        # time-parallel part: sampler recurrence over lead + sequence (the carried state), projections
        lead_frames = la.shape[1] // self.ratio
        sampled, _ = self.sampling_lstm(self.acoustic_projection(torch.cat([la, a], dim=1)), None)
        sampled = sampled[:, lead_frames:]                                          # [B, T, Hs]
        Hs = sampled.shape[-1]
        W, bias = self.feature_projection.weight, self.feature_projection.bias
        feats = torch.cat([sampled, mp], dim=-1).transpose(0, 1).contiguous()       # [T, B, Hs + P]
        base = _linear(feats, W[:, :Hs + P], bias)                                   # [T, B, Hd]
        # ground-truth previous frame with the reference's one-frame lag (Q5): ms[0], ms[0], ms[1], ...
        gt_prev = torch.cat([ms[:, :1], ms[:, :-1]], dim=1).transpose(0, 1).contiguous()
        ff = self.feed_forward
        eps = self.layerd_lstm.lstm_layered[0].lstm_module.layer_norm.eps
        pred = _rollout.rollout(base, gt_prev, mask, W[:, Hs + P:], layers, ff.input.weight, ff.input.bias,
                                ff.mapping.weight, ff.mapping.bias, relu=hasattr(ff, "relu"), eps=eps)
        return pred.transpose(0, 1).contiguous(), target

    def _prediction_wavefront(self, batch, use_scheduled_sampling, full_generation, sampling_mask):
        dev = self.device
        (a, _), (mp, _), (ms, _), (la, _), (lp, _), (ls, _) = batch[:6]
        a, mp, ms, la = a.to(dev), mp.to(dev), ms.to(dev), la.to(dev)
        target = batch[-1][0].to(dev)
        B, T, P = mp.shape
        mask = self._resolve_mask(T, B, use_scheduled_sampling, full_generation, sampling_mask)
        if T == 0:
            return ms.new_zeros((B, 0, P)), target
        mask = mask.to(dev)
        if mask.dim() == 1:
            mask = mask.view(T, 1).expand(T, B)
        # --- time-parallel part -----------------------------------------------------------------
        lead_frames = la.shape[1] // self.ratio
        audio = self.acoustic_projection(torch.cat([la, a], dim=1))
        sampled, _ = self.sampling_lstm(audio, None)          # one recurrence over lead + sequence
        sampled = sampled[:, lead_frames:]                     # [B, T, Hs]
        Hs = sampled.shape[-1]
        W, bias = self.feature_projection.weight, self.feature_projection.bias
        base = _linear(torch.cat([sampled, mp], dim=-1), W[:, :Hs + P], bias)  # [B,T,Hd]
        W_prev = W[:, Hs + P:]                                 # [Hd, P]
        base = base.reshape(B * T, -1)
        # --- feedback depth of every (b, t) -----------------------------------------------------
        t_idx = torch.arange(T, device=dev).view(T, 1)
        last_false = torch.where(mask, torch.full_like(t_idx, -1), t_idx).cummax(dim=0).values  # [T,B]
        depth = torch.zeros((T, B), dtype=torch.long, device=dev)
        depth[1:] = (t_idx - last_false)[:-1]
        depth = depth.t().reshape(-1)                           # position p = b*T + t
        order = torch.argsort(depth, stable=True)
        counts = torch.bincount(depth).tolist()                 # the one host read of the rollout
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        slot = torch.empty_like(order)
        slot[order] = torch.arange(B * T, device=dev) - torch.tensor(starts[:-1], device=dev)[depth[order]]
        # ground-truth previous frame with the reference's one-frame lag (Q5): ms[0], ms[0], ms[1], ...
        gt_prev = torch.cat([ms[:, :1], ms[:, :-1]], dim=1).reshape(B * T, P)
        outs = []
        for d, n in enumerate(counts):
            if n == 0:
                outs.append(base.new_zeros((0, P)))
                continue
            idx = order[starts[d]:starts[d + 1]]
            prev = gt_prev.index_select(0, idx) if d == 0 else outs[d - 1].index_select(0, slot[idx - 1])
            x = base.index_select(0, idx) + _linear(prev, W_prev, None)
            h, _ = self.layerd_lstm(x.unsqueeze(1), None)       # stateless predictor blocks (Q2), T = 1
            outs.append(self.feed_forward(h).squeeze(1))
        inverse = torch.empty_like(order)
        inverse[order] = torch.arange(B * T, device=dev)
        pred = torch.cat(outs, dim=0).index_select(0, inverse).view(B, T, P)
        return pred.contiguous(), target

    def batch_forming(self, batch):
        formed, length = self.form_generation_init(batch)
        return formed, self.gen_dummy_input(batch), length

    def warmup_model(self, dummy_input, batch):
        """run the leading segment only; returns (sampler state, None)  (Q2)"""
        return self.forward(*dummy_input[:3], *batch[3:6], cell_state=None)[2]

    def draw_sampling_mask(self, length: int, batch_size: int) -> torch.Tensor:
        """Scheduled-sampling decisions for one rollout.  Without a seed: the reference's draw
        (``torch.rand(length) < epoch/max_epochs``, global CPU generator, shared by the batch).  With
        ``sampling_seed``: Philox4x32-10 per (step, sample), advancing ``sampling_offset`` by ``length``."""
        rate = self.current_epoch / self.max_epochs
        if self.sampling_seed is None:
            return torch.rand(length) < rate
        mask = philox_sampling_mask(self.sampling_seed, self.sampling_offset, rate, length, batch_size,
                                    self.device, shared=not self.sampling_per_sample)
        self.sampling_offset += length
        return mask

    def head_motion_generation(self, formed_batch, dummy_input, length, cell_state=None,
                               use_scheduled_sampling: bool = False, full_generation: bool = False,
                               sampling_mask: Optional[torch.Tensor] = None):
        motion_s = formed_batch[2][0]
        batch_size = motion_s.shape[1]
        sampling_mask = self._resolve_mask(length, batch_size, use_scheduled_sampling, full_generation,
                                           sampling_mask)
        per_sample = sampling_mask.dim() == 2
        if per_sample:
            sampling_mask = sampling_mask.to(motion_s.device)
        else:
            sampling_mask = sampling_mask.cpu()
        y = motion_s[0]
        outs = []
        for step in range(length):
            y, cell_state = self.generate_one_step(step, formed_batch, y, dummy_input, cell_state)
            outs.append(y)
            if per_sample:
                y = torch.where(sampling_mask[step].view(-1, 1, 1), y, motion_s[step])
            elif not bool(sampling_mask[step]):
                y = motion_s[step]
        if not outs:
            return motion_s.new_zeros((batch_size, 0, motion_s.shape[-1]))
        return torch.cat(outs, dim=1).contiguous()

    def generate_one_step(self, step, formed_batch, previous, dummy_input, cell_state):
        fbank, motion_p = formed_batch[0], formed_batch[1]
        ones = torch.ones(motion_p[0][step].shape[0], dtype=torch.long, device=self.device)
        y, _, cell_state = self.forward((fbank[0][step], fbank[1]), (motion_p[0][step], motion_p[1]),
                                        (previous, ones), dummy_input[3], dummy_input[4], dummy_input[5],
                                        cell_state)
        return y, cell_state

    def form_generation_init(self, batch):
        (fbank, lf), (motion_p, lp), (motion_s, ls) = batch[0], batch[1], batch[2]
        bsz, length = motion_p.shape[0], motion_p.shape[1]
        dev = self.device
        # [B, T*ratio, F] -> [T, B, ratio, F];  [B, T, P] -> [T, B, 1, P]
        fbank = fbank.to(dev).view(bsz, length, self.ratio, fbank.shape[-1]).transpose(0, 1).contiguous()
        motion_p = motion_p.to(dev).transpose(0, 1).unsqueeze(2).contiguous()
        motion_s = motion_s.to(dev).transpose(0, 1).unsqueeze(2).contiguous()
        return [(fbank, lf), (motion_p, lp), (motion_s, ls)], length

    def gen_dummy_input(self, batch):
        """zero-length stand-ins (the step-wise forward has no leading segment)"""
        return [(torch.empty((m.shape[0], 0, m.shape[2]), dtype=m.dtype, device=self.device), l)
                for (m, l) in batch]
