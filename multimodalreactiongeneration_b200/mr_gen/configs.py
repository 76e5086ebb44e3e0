"""Model configurations of the BASELINE.json workloads (synthetic shapes: stereo 2x40-d log-mel = 80
acoustic features, 6-d head pose, 30 fps for both streams so ratio 1, hidden 256, 2 layers)."""
from .utils.config import DictConfig


def simple_lstm_cfg(hidden=256, layers=2, bidirectional=False, acoustic=80, pose=6, heads=8, att_layers=3):
    lstm = hidden // 2 if bidirectional else hidden
    model = DictConfig(
        acostic_feat_size=acoustic, motion_feat_size=pose,
        motion_num_lstm=1, acostic_num_lstm=1, acostic_num_layers=layers, motion_num_layers=layers,
        acostic_lstm_size=lstm, motion_lstm_size=lstm, acostic_affine_size=hidden, motion_affine_size=hidden,
        acostic_output_size=hidden, motion_output_size=hidden,
        att_heads=heads, att_num_layers=att_layers, att_use_residual=True, att_use_layer_norm=True,
        dropout_rate=0, output_size=pose, bidirectional=bidirectional,
        use_layer_norm=True, use_relu=True, use_mixing=True, use_residual=True,
        decoder_num_layers=layers, decoder_num_lstm=1, decoder_lstm_size=lstm, decoder_affine_size=hidden,
        decoder_bottleneck_size=64, decoder_output_size=hidden, decoder_mapping_size=64,
        decoder_bidirectional=bidirectional, decoder_use_layer_norm=True, decoder_use_relu=True,
        decoder_use_mixing=True, decoder_use_residual=True, all_static=False, delta_loss_scale=1.0)
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100,
                       momentum=0.9)
    metrics = DictConfig(use_centroid=True, use_angle=True, delta_order=0)
    return model, optim, metrics


def lstm_with_sampling_cfg(hidden=256, layers=2, sampler_hidden=128, sampler_layers=2, acoustic=80,
                           scheduled=True, max_epochs=100, ratio=1, seed=None):
    model = DictConfig(
        nmels=acoustic - 1, delta_order=0, use_centroid=True, use_angle=True,
        sampler_hidden_size=sampler_hidden, sampler_num_layers=sampler_layers, sampler_dropout_rate=0,
        sampling_rate=48000, shift=1600 // ratio, fps=30, pred_fps=30.0,
        hidden_size=hidden, bottleneck_size=64, num_layers=layers, num_lstm=1, dropout_rate=0.0,
        use_layer_norm=True, use_relu=True, use_mixing=False, use_residual=True,
        delta_loss_scale=1, loss_type="huber", loss_reduction="mean", huber_delta=1.0, smoothl1_beta=1.0,
        use_scheduled_sampling=scheduled, max_epochs=max_epochs)
    if seed is not None:
        model["sampling_seed"] = seed
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100,
                       momentum=0.9)
    metrics = DictConfig(use_centroid=True, use_angle=True, delta_order=0)
    return model, optim, metrics


def metaformer_cfg(hidden=256, blocks=5, encoder_layers=5, bottleneck=64, heads=4, acoustic=80, ratio=1,
                   scheduled=False, max_epochs=100, mixers=("lstm", "lstm", "lstm"), seed=None):
    """lstmformer (``Metaformer``): the reference's mr_gen/model/lstmformer/config.yaml model block on the synthetic
    shapes — modalities (audio, partner motion, own motion), own motion is the main stream."""
    model = DictConfig(
        main_modal_idx=2, hidden_size=hidden, dropout=0.0, num_block=blocks, num_layerd=1,
        encoder_num_layer=encoder_layers, num_internal_layer=1, residual=True, residual_layer_norm=True, bias=True,
        emb_mixers=list(mixers), bottleneck_size=bottleneck, nonlinearity="none", ffn_nonlinearity="relu",
        proj_size=0, num_heads=heads, add_bias_kv=False, add_zero_attn=False, max_context_len=10,
        repeat_with_encoder=False, interlayer_residual=False, interlayer_residual_norm=True,
        sampling_rate=48000, shift=1600 // ratio, pred_fps=30.0,
        modalities=["audio", "motion", "motion"], use_centroid=True, use_angle=True, nmels=acoustic - 1,
        delta_order=0, loss_type="huber", loss_reduction="mean", huber_delta=1.0, smoothl1_beta=1.0,
        delta_loss_scale=1, use_scheduled_sampling=scheduled, max_epochs=max_epochs)
    if seed is not None:
        model["sampling_seed"] = seed
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100,
                       momentum=0.9)
    metrics = DictConfig(use_centroid=True, use_angle=True, delta_order=0)
    return model, optim, metrics
