"""Mirror of mr_gen/model/lstmformer/lstmformer.py (``Metaformer`` :70-559) on the B200 path.

Three modalities — own head motion (main), partner audio, partner head motion — each embedded by a stack of
token mixers (LSTM mixers by default: 1 + 5 + 5 in the first block, 1 per later block = 15 recurrences, all on
``B200LSTM``), the main stream attending causally to the two others in every block (masked multi-head attention,
projections on the tcgen05 GEMM), then feed-forwards; output = next-frame head motion for every frame.

Reference behaviour reproduced on purpose: Q3 no recurrent / KV state survives between
``forward`` calls, so the autoregressive rollout is stateless per step; Q4 one scheduled-sampling decision per
time step shared by the batch (``torch.rand`` when no mask is supplied); Q5 one-frame lag of teacher forcing;
Q6 gradients flow through fed-back predictions; Q7 mean loss over padded positions; Q11 padded frames are
zeroed before the rollout (all three modalities) and before the teacher-forced forward (own motion only);
``prediction`` multiplies the ``[B,T,P]`` target by the ``[T,B,1,P]`` padding mask of the time-major own motion,
which BROADCASTS to ``[T,B,T,P]`` (:413-414) — kept, because the scheduled-sampling loss is computed on it.

Device-resident rollout (``rollout="wavefront"``, default): because of Q3 the prediction of frame t is a pure
function of (audio frames of t, partner pose of t, previous own pose), and the previous own pose is either ground
truth or the prediction of t-1.  All (sample, frame) positions of equal feedback depth are therefore evaluated
together as one batch of length-1 sequences: the number of model evaluations is the longest run of fed-back steps
+ 1, not T, with identical arithmetic per position.  ``rollout="stepwise"`` is the reference's Python time loop."""
from typing import List, Optional, Tuple

import torch
from torch import nn

from ...utils.lightning_shim import LightningModule
from ...utils.metrics import MultiTargetMetrics, gen_target_dict
from ..lstm_with_sampling.lstm_with_sample import philox_sampling_mask
from ..simple_lstm.simple_lstm import _configure_optimizers
from ..utils.argparser import feedforward_block_argments, mixer_layerd_argments_select
from ..utils.multi_modal_metaformer import MultiModalMetaformer, gen_attention_mask, gen_attention_mask_spec
from ..utils.values import PADDING_VALUE

InputTypes = Tuple[torch.Tensor, torch.Tensor]


class Metaformer(LightningModule):
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
        pred_fps = model.pred_fps
        acoustic_fps = model.sampling_rate / model.shift
        ratio = acoustic_fps / pred_fps
        if ratio != int(ratio):
            raise ValueError("pred_fps must be a divisor of acoustic_fps",
                             f"pred_fps: {pred_fps}, acoustic_fps: {acoustic_fps}, ratio: {ratio}")
        self.ratio = int(ratio)

        # context length of each integrator, in frames of the modality it attends to
        self.modalities = list(model.modalities)
        self.other_modalities = list(model.modalities)
        self.other_modalities.pop(model.main_modal_idx)
        self.max_context_len = model.max_context_len
        rates = {"audio": acoustic_fps, "motion": pred_fps}
        if any(m not in rates for m in self.other_modalities):
            raise ValueError("invalid modality")
        self.context_len = [self.max_context_len * rates[m] for m in self.other_modalities]

        self.acoustic_input_size = (model.nmels + 1) * (model.delta_order + 1)
        self.motion_base_size = (model.use_centroid + model.use_angle) * 3
        self.motion_input_size = self.motion_base_size * (model.delta_order + 1)

        self.modal_num = len(model.modalities)
        self.hidden_dim = model.hidden_size
        self.num_block = model.num_block
        self.num_heads = model.num_heads
        self.main_mixer_type: str = model.emb_mixers[model.main_modal_idx]
        self.other_mixer_type: list = list(model.emb_mixers)
        self.other_mixer_type.pop(model.main_modal_idx)
        self.repeat_with_encoder = model.repeat_with_encoder
        self.interlayer_residual = model.interlayer_residual
        self.interlayer_residual_norm = model.interlayer_residual_norm
        self.use_device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

        self.common_configs = dict(
            hidden_size=self.hidden_dim, input_projection=False, output_projection=False, self_attention=True,
            num_heads=model.num_heads, dropout=model.dropout, batch_first=True, bidirectional=False,
            proj_size=model.proj_size, add_bias_kv=model.add_bias_kv, add_zero_attn=model.add_zero_attn,
            kdim=self.hidden_dim, vdim=self.hidden_dim, max_context_len=125, num_layerd=model.num_layerd,
            num_internal_layer=model.num_internal_layer, nonlinearity=model.nonlinearity,
            bottleneck_size=model.bottleneck_size, residual=model.residual,
            residual_layer_norm=model.residual_layer_norm, bias=model.bias, device=self.use_device)
        self.main_mixer_configs = mixer_layerd_argments_select(self.main_mixer_type, **self.common_configs)
        encoder = dict(self.common_configs, num_layerd=model.encoder_num_layer)
        self.other_mixer_configs = [mixer_layerd_argments_select(kind, **encoder) for kind in self.other_mixer_type]
        cross = dict(self.common_configs, self_attention=False)
        self.integrate_mixer_configs = [dict(mixer_layerd_argments_select("mha", **cross), max_context_len=n)
                                        for n in self.context_len]
        self.feedforward_configs = feedforward_block_argments(
            hidden_size=self.hidden_dim, bottleneck_size=model.bottleneck_size, nonlinearity=model.ffn_nonlinearity,
            residual=model.residual, residual_layer_norm=model.residual_layer_norm, bias=model.bias,
            device=self.use_device)
        self.output_feedforward_configs = feedforward_block_argments(
            hidden_size=self.hidden_dim, bottleneck_size=model.bottleneck_size, output_size=self.motion_input_size,
            nonlinearity=model.ffn_nonlinearity, residual=False, bias=model.bias, device=self.use_device)
        self.metaformer = MultiModalMetaformer(
            modal_num=self.modal_num, hidden_dim=self.hidden_dim, num_layer=self.num_block,
            main_modal_feature_dim=self.motion_input_size, main_mixer_type=self.main_mixer_type,
            main_mixer_configs=self.main_mixer_configs, integrate_mixer_configs=self.integrate_mixer_configs,
            feedforward_configs=self.feedforward_configs,
            output_feedforward_configs=self.output_feedforward_configs,
            other_modal_feature_dim=[self.acoustic_input_size, self.motion_input_size],
            other_mixer_type=self.other_mixer_type, other_mixer_configs=self.other_mixer_configs,
            repeat_with_encoder=self.repeat_with_encoder, interlayer_residual=self.interlayer_residual,
            interlayer_residual_norm=self.interlayer_residual_norm)

        ranges = gen_target_dict(metrics)
        self.train_metrics = MultiTargetMetrics(target_range=ranges, prefix="train_")
        self.valid_metrics = MultiTargetMetrics(target_range=ranges, prefix="valid_")
        self.genrt_metrics = MultiTargetMetrics(target_range=ranges, prefix="genrt_")
        self.optimizer = None
        self.lr_scheduler = None
        self.delta_loss_scale = model.get("delta_loss_scale", 1.0)
        self.delta_order = metrics.delta_order
        # extensions shared with LSTMwithSample: counter-based sampling masks, rollout schedule
        self.sampling_seed: Optional[int] = model.get("sampling_seed", None)
        self.sampling_offset = 0
        self.sampling_per_sample = model.get("sampling_per_sample", True)
        self.rollout = model.get("rollout", "wavefront")

    # ------------------------------------------------------------------------------------------
    def forward(self, acoustic_partner: InputTypes, motion_partner: InputTypes, motion_self: InputTypes,
                leading_acoustic_partner: InputTypes, leading_motion_partner: InputTypes,
                leading_motion_self: InputTypes, hxs=None):
        dev = self.device
        audio = torch.cat([leading_acoustic_partner[0].to(dev), acoustic_partner[0].to(dev)], dim=1)
        partner = torch.cat([leading_motion_partner[0].to(dev), motion_partner[0].to(dev)], dim=1)
        own = torch.cat([leading_motion_self[0].to(dev), motion_self[0].to(dev)], dim=1)

        def mask(q, k):
            if q.is_cuda:   # the rule itself: the fused attention kernel evaluates it, no mask tensor exists
                return gen_attention_mask_spec(q, k, PADDING_VALUE)
            m = gen_attention_mask(q, k, self.num_heads, PADDING_VALUE)
            return m.reshape(-1, m.shape[2], m.shape[3])   # nn.MultiheadAttention's form, as the reference (:269-294)

        kinds = [self.main_mixer_type] + self.other_mixer_type
        self_masks = [mask(s, s) if kind == "mha" else None for kind, s in zip(kinds, (own, audio, partner))]
        y, _, hxs = self.metaformer(
            own, [audio, partner], hxs, (None, None, self_masks[0]),
            [(None, None, self_masks[1]), (None, None, self_masks[2])], [mask(own, audio), mask(own, partner)])
        return y, hxs

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
            lead_len = batch[4][0].shape[1]
            own, own_len = batch[2]
            batch[2] = (own * (own != PADDING_VALUE).int(), own_len)   # in place on the caller's list, like :365-366
            y, _ = self.forward(*batch[:-1])
            y = y[:, lead_len:]
            target = batch[-1][0].to(y.device)
        y, target = self._mask_padding(y, target)
        scaler = torch.ones_like(y)
        scaler[..., y.shape[2] // (self.delta_order + 1):] = float(self.delta_loss_scale) ** 0.5
        loss = self.lossfun()(y * scaler, target * scaler)
        self.log("train_loss", loss, prog_bar=True, logger=True)
        self.log_dict(self.train_metrics(y * scaler, target * scaler), logger=True, on_epoch=True, on_step=True)
        return {"loss": loss}

    def validation_step(self, batch: List[InputTypes], *args):
        lead_len = batch[4][0].shape[1]
        y, _ = self.forward(*batch[:-1])
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
    # autoregressive rollout (reference :399-559)
    # ------------------------------------------------------------------------------------------
    def prediction(self, batch: List[InputTypes], use_scheduled_sampling: bool = False,
                   full_generation: bool = False, sampling_mask: Optional[torch.Tensor] = None):
        formed, dummy, length, pad_mask = self.batch_forming(batch)
        target = batch[-1][0].to(self.device) * pad_mask            # [B,T,P] x [T,B,1,P] -> [T,B,T,P] (sic)
        if self.rollout == "wavefront":
            pred = self._generation_wavefront(formed, dummy, length, use_scheduled_sampling, full_generation,
                                              sampling_mask)
        else:
            state = self.warmup_model(dummy, batch)
            pred = self.head_motion_generation(formed, dummy, length, state, use_scheduled_sampling,
                                               full_generation, sampling_mask)
        return pred, target

    def batch_forming(self, batch):
        formed, length, pad_mask = self.form_generation_init(batch)
        return formed, self.gen_dummy_input(batch), length, pad_mask

    def warmup_model(self, dummy_input, batch):
        """leading segment only; what comes back holds no state (Q3)"""
        return self.forward(*dummy_input[:3], *batch[3:6], hxs=None)[1]

    def draw_sampling_mask(self, length: int, batch_size: int) -> torch.Tensor:
        """``torch.rand(length) < epoch/max_epochs`` (:476) unless ``sampling_seed`` selects the Philox stream."""
        rate = self.current_epoch / self.max_epochs
        if self.sampling_seed is None:
            return torch.rand(length) < rate
        mask = philox_sampling_mask(self.sampling_seed, self.sampling_offset, rate, length, batch_size,
                                    self.device, shared=not self.sampling_per_sample)
        self.sampling_offset += length
        return mask

    def _resolve_mask(self, length, batch_size, use_scheduled_sampling, full_generation, sampling_mask):
        if sampling_mask is not None:
            return sampling_mask
        if use_scheduled_sampling:
            return self.draw_sampling_mask(length, batch_size)
        return torch.full((length,), bool(full_generation), dtype=torch.bool)

    def head_motion_generation(self, formed_batch, dummy_input, length, cell_state=None,
                               use_scheduled_sampling: bool = False, full_generation: bool = False,
                               sampling_mask: Optional[torch.Tensor] = None):
        motion_s = formed_batch[2][0]
        batch_size = motion_s.shape[1]
        sampling_mask = self._resolve_mask(length, batch_size, use_scheduled_sampling, full_generation,
                                           sampling_mask)
        per_sample = sampling_mask.dim() == 2
        sampling_mask = sampling_mask.to(motion_s.device) if per_sample else sampling_mask.cpu()
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

    def _generation_wavefront(self, formed_batch, dummy_input, length, use_scheduled_sampling, full_generation,
                              sampling_mask):
        fbank, motion_p, motion_s = formed_batch[0][0], formed_batch[1][0], formed_batch[2][0]
        T, B = length, motion_s.shape[1]
        P = motion_s.shape[-1]
        dev = motion_s.device
        mask = self._resolve_mask(T, B, use_scheduled_sampling, full_generation, sampling_mask)
        if T == 0:
            return motion_s.new_zeros((B, 0, P))
        mask = mask.to(dev)
        if mask.dim() == 1:
            mask = mask.view(T, 1).expand(T, B)
        # feedback depth of position (t, b): number of consecutive fed-back steps that end at t
        t_idx = torch.arange(T, device=dev).view(T, 1)
        last_false = torch.where(mask, torch.full_like(t_idx, -1), t_idx).cummax(dim=0).values   # [T, B]
        depth = torch.zeros((T, B), dtype=torch.long, device=dev)
        depth[1:] = (t_idx - last_false)[:-1]
        depth = depth.reshape(-1)                                # position p = t*B + b (time-major, like `formed`)
        order = torch.argsort(depth, stable=True)
        counts = torch.bincount(depth).tolist()                  # the one host read of the rollout
        starts = [0]
        for c in counts:
            starts.append(starts[-1] + c)
        slot = torch.empty_like(order)
        slot[order] = torch.arange(T * B, device=dev) - torch.tensor(starts[:-1], device=dev)[depth[order]]
        audio = fbank.reshape(T * B, self.ratio, fbank.shape[-1])
        partner = motion_p.reshape(T * B, 1, P)
        # ground-truth previous frame with the one-frame lag (Q5): ms[0], ms[0], ms[1], ...
        gt_prev = torch.cat([motion_s[:1], motion_s[:-1]], dim=0).reshape(T * B, 1, P)
        outs = []
        for d, n in enumerate(counts):
            if n == 0:
                outs.append(motion_s.new_zeros((0, 1, P)))
                continue
            idx = order[starts[d]:starts[d + 1]]
            prev = gt_prev.index_select(0, idx) if d == 0 else outs[d - 1].index_select(0, slot[idx - B])
            lens = torch.ones(n, dtype=torch.long, device=dev)
            lead = [(t[0].new_empty((n, 0, t[0].shape[2])), None) for t in dummy_input[3:6]]
            y, _ = self.forward((audio.index_select(0, idx), lens), (partner.index_select(0, idx), lens),
                                (prev, lens), *lead, None)
            outs.append(y)
        inverse = torch.empty_like(order)
        inverse[order] = torch.arange(T * B, device=dev)
        pred = torch.cat(outs, dim=0).index_select(0, inverse).view(T, B, P)
        return pred.transpose(0, 1).contiguous()

    def generate_one_step(self, step, formed_batch, previous, dummy_input, cell_state):
        fbank, motion_p = formed_batch[0], formed_batch[1]
        ones = torch.ones(motion_p[0][step].shape[0], dtype=torch.long, device=self.device)
        return self.forward((fbank[0][step], fbank[1]), (motion_p[0][step], motion_p[1]), (previous, ones),
                            dummy_input[3], dummy_input[4], dummy_input[5], cell_state)

    def form_generation_init(self, batch):
        (fbank, lf), (motion_p, lp), (motion_s, ls) = batch[0], batch[1], batch[2]
        bsz, length = motion_p.shape[0], motion_p.shape[1]
        dev = self.device
        # [B, T*ratio, F] -> [T, B, ratio, F];  [B, T, P] -> [T, B, 1, P];  padding zeroed (Q11)
        fbank = fbank.to(dev).view(bsz, length, self.ratio, fbank.shape[-1]).transpose(0, 1).contiguous()
        motion_p = motion_p.to(dev).transpose(0, 1).unsqueeze(2).contiguous()
        motion_s = motion_s.to(dev).transpose(0, 1).unsqueeze(2).contiguous()
        keep_s = (motion_s != PADDING_VALUE).int()
        fbank = fbank * (fbank != PADDING_VALUE).int()
        motion_p = motion_p * (motion_p != PADDING_VALUE).int()
        motion_s = motion_s * keep_s
        return [(fbank, lf), (motion_p, lp), (motion_s, ls)], length, keep_s

    def gen_dummy_input(self, batch):
        """zero-length stand-ins (the step-wise forward has no leading segment)"""
        return [(torch.empty((m.shape[0], 0, m.shape[2]), dtype=m.dtype, device=self.device), l)
                for (m, l) in batch]
