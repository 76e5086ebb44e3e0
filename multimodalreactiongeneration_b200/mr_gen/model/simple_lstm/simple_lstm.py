"""Mirror of mr_gen/model/simple_lstm/simple_lstm.py (AcousticEncoder :48, MotionEncoder :74,
MotionDecoder :100, SimpleLSTM :146) on the B200 LSTM path.

Documented deviation: at the reference's HEAD the three sub-modules pass the
``(tensor, states)`` tuple returned by ``LSTMLayerd`` on as if it were a tensor and ``forward`` raises;
here element ``[0]`` is taken, which is what the code did before its ``hx`` refactor."""
import os
from collections import OrderedDict
from typing import Dict

import torch
from torch import nn

from ....linear import B200Linear

from ...utils.lightning_shim import LightningModule
from ...utils.metrics import MultiTargetMetrics, gen_target_dict
from ..utils.lstm_block import LSTMLayerd
from ..utils.multi_modal_att import MultimodalAttention


class AcousticEncoder(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.embed_layer = B200Linear(cfg.acostic_feat_size, cfg.acostic_affine_size)
        self.acostic_lstm = LSTMLayerd(
            input_size=cfg.acostic_affine_size, lstm_hidden_size=cfg.acostic_lstm_size,
            affine_hidden_size=cfg.acostic_affine_size, num_layers=cfg.acostic_num_layers,
            num_layers_per_block=cfg.acostic_num_lstm, output_size=cfg.acostic_output_size,
            dropout=cfg.dropout_rate, bidirectional=cfg.bidirectional, use_layer_norm=cfg.use_layer_norm,
            use_relu=cfg.use_relu, use_mixing=cfg.use_mixing, use_residual=cfg.use_residual)

    def forward(self, acoustic_feature: torch.Tensor) -> torch.Tensor:
        return self.acostic_lstm(self.embed_layer(acoustic_feature))[0]


class MotionEncoder(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.embed_layer = B200Linear(cfg.motion_feat_size, cfg.motion_affine_size)
        self.motion_lstm = LSTMLayerd(
            input_size=cfg.motion_affine_size, lstm_hidden_size=cfg.motion_lstm_size,
            affine_hidden_size=cfg.motion_affine_size, num_layers=cfg.motion_num_layers,
            num_layers_per_block=cfg.motion_num_lstm, output_size=cfg.motion_output_size,
            dropout=cfg.dropout_rate, bidirectional=cfg.bidirectional, use_layer_norm=cfg.use_layer_norm,
            use_relu=cfg.use_relu, use_mixing=cfg.use_mixing, use_residual=cfg.use_residual)

    def forward(self, head_feature: torch.Tensor) -> torch.Tensor:
        return self.motion_lstm(self.embed_layer(head_feature))[0]


class MotionDecoder(nn.Module):
    def __init__(self, cfg) -> None:
        super().__init__()
        self.decoder_lstm = LSTMLayerd(
            input_size=cfg.motion_output_size, lstm_hidden_size=cfg.decoder_lstm_size,
            affine_hidden_size=cfg.decoder_affine_size, bottleneck_size=cfg.decoder_bottleneck_size,
            num_layers=cfg.decoder_num_layers, num_layers_per_block=cfg.decoder_num_lstm,
            output_size=cfg.decoder_output_size, dropout=cfg.dropout_rate,
            bidirectional=cfg.decoder_bidirectional, use_layer_norm=cfg.decoder_use_layer_norm,
            use_relu=cfg.decoder_use_relu, use_mixing=cfg.decoder_use_mixing,
            use_residual=cfg.decoder_use_residual)
        head = [("input", B200Linear(cfg.decoder_output_size, cfg.decoder_mapping_size))]
        if cfg.decoder_use_relu:
            head.append(("relu", nn.ReLU()))
        head.append(("output", B200Linear(cfg.decoder_mapping_size, cfg.output_size)))
        self.mapping = nn.Sequential(OrderedDict(head))

    def seq_reshape(self, x: torch.Tensor) -> torch.Tensor:
        """keep the last time step: [..., T, F] -> [..., 1, F]"""
        return x[..., -1:, :]

    def forward(self, att_embedded: torch.Tensor) -> torch.Tensor:
        # reference: mapping(seq_reshape(decoder_lstm(x)[0])) — only the last frame of the decoder output is read, so the
        # last block's per-token tail runs on that frame alone (same numbers, LSTMLayerd.forward_last_step)
        return self.mapping(self.decoder_lstm.forward_last_step(att_embedded))


class SimpleLSTM(LightningModule):
    # sub-modules whose backward completes first, in that order: the trainer all-reduces their gradient slice while
    # the encoders' BPTT still runs (each is used once per forward)
    ddp_overlap_children = ("motion_decoder", "multimodal_att")

    def __init__(self, cfg, optim, metrics):
        super().__init__()
        self.cfg, self.optim, self.metrics = cfg, optim, metrics
        self.acoustic_encoder = AcousticEncoder(cfg)
        self.motion_encoder = MotionEncoder(cfg)
        self.multimodal_att = MultimodalAttention(
            modal1_feat_size=cfg.acostic_output_size, modal2_feat_size=cfg.motion_output_size,
            num_head=cfg.att_heads, num_layers=cfg.att_num_layers, dropout=cfg.dropout_rate,
            use_residual=cfg.att_use_residual, use_layer_norm=cfg.att_use_layer_norm)
        self.motion_decoder = MotionDecoder(cfg)
        ranges = gen_target_dict(metrics)
        self.train_metrics = MultiTargetMetrics(target_range=ranges, prefix="train_")
        self.valid_metrics = MultiTargetMetrics(target_range=ranges, prefix="valid_")
        self.optimizer = None
        self.lr_scheduler = None
        self.delta_loss_scale = cfg.get("delta_loss_scale", 1.0)
        self.all_static = cfg.get("all_static", False)
        self.delta_order = metrics.delta_order

    def forward(self, acoustic_feature: torch.Tensor, motion_feature: torch.Tensor) -> torch.Tensor:
        if acoustic_feature.is_cuda and os.environ.get("MRG_TWO_STREAMS", "1") != "0":
            audio, motion = self._encode_two_streams(acoustic_feature, motion_feature)
        else:
            audio = self.acoustic_encoder(acoustic_feature)
            motion = self.motion_encoder(motion_feature)
        fused = self.multimodal_att(motion, audio)
        return self.motion_decoder(fused)

    def _encode_two_streams(self, acoustic_feature, motion_feature):
        """The two encoders are independent until the cross-modal attention: run them on two CUDA streams with
        half of the co-resident clusters each.  A persistent recurrent kernel is latency-bound per timestep, so two of
        them side by side (more rows per cluster, fewer clusters) finish sooner than one after the other, and the
        projection GEMMs of one stack fill the SMs the other's recurrence leaves idle.  Autograd replays each
        backward node on the stream of its forward, so the BPTT kernels overlap the same way."""
        from ....lstm import cluster_budget
        from .... import _cabi
        main = torch.cuda.current_stream(acoustic_feature.device)
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != acoustic_feature.device:
            self._side_stream = torch.cuda.Stream(device=acoustic_feature.device)
            info = [__import__("ctypes").c_int() for _ in range(5)]
            _cabi.lib().mrg_device_info(*[__import__("ctypes").byref(v) for v in info])
            self._half_clusters = max(1, info[1].value // 2)
        side = self._side_stream
        side.wait_stream(main)
        with cluster_budget(self._half_clusters):
            with torch.cuda.stream(side):
                motion = self.motion_encoder(motion_feature)
            audio = self.acoustic_encoder(acoustic_feature)
        main.wait_stream(side)
        motion.record_stream(main)
        return audio, motion

    def lossfun(self):
        return nn.MSELoss(reduction="mean")

    def configure_optimizers(self) -> Dict:
        return _configure_optimizers(self)

    def split_and_form(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Re-derive velocity / acceleration of the prediction from the last context frame
        (simple_lstm.py:223-237)."""
        if self.delta_order == 0:
            return y
        size = (self.metrics.use_centroid + self.metrics.use_angle) * 3
        last = x[:, -1:, :]
        pos = y[..., :size]
        vel = pos - last[..., :size]
        if self.delta_order == 1:
            return torch.cat([pos, vel], dim=-1)
        return torch.cat([pos, vel, vel - last[..., size:2 * size]], dim=-1)

    def _scaled(self, y, target):
        scaler = torch.ones_like(y)
        start = y.shape[2] // (self.delta_order + 1)
        scaler[:, :, start:] = float(self.delta_loss_scale) ** 0.5
        return y * scaler, target * scaler

    def training_step(self, batch, *args):
        acoustic_feature, motion_feature, motion_target = batch
        y = self.forward(acoustic_feature, motion_feature)
        if self.all_static:
            y = self.split_and_form(motion_feature, y)
        ys, ts = self._scaled(y, motion_target)
        loss = self.lossfun()(ys, ts)
        self.log("train_loss", loss, prog_bar=True, logger=True)
        self.log_dict(self.train_metrics(ys, ts), logger=True, on_epoch=True, on_step=True)
        return {"loss": loss}

    def validation_step(self, batch, *args):
        acoustic_feature, motion_feature, motion_target = batch
        y = self.forward(acoustic_feature, motion_feature)
        if self.all_static:
            y = self.split_and_form(motion_feature, y)
        loss = self.lossfun()(y, motion_target)
        self.log("val_loss", loss, prog_bar=True, logger=True)
        self.log_dict(self.valid_metrics(y, motion_target), logger=True, on_epoch=True, on_step=True)
        return {"loss": loss}


def _configure_optimizers(module) -> Dict:
    """AdamW / SGD + optional CosineAnnealingLR, keyed like the reference (simple_lstm.py:193-221)."""
    o = module.optim
    if o.use_optimizer == "adam":
        module.optimizer = torch.optim.AdamW(module.parameters(), lr=o.lr, weight_decay=o.weight_decay)
    elif o.use_optimizer == "sgd":
        module.optimizer = torch.optim.SGD(module.parameters(), lr=o.lr, momentum=o.momentum,
                                           weight_decay=o.weight_decay)
    else:
        raise ValueError("invalid optimizer type")
    out = {"optimizer": module.optimizer}
    if o.use_lr_sched:
        module.lr_scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(module.optimizer, T_max=o.max_epochs)
        out["lr_scheduler"] = {"scheduler": module.lr_scheduler, "monitor": "val_loss"}
    return out
