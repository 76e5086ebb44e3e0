"""Mirror of mr_gen/model/utils/multi_modal_att.py — cross-modal nn.MultiheadAttention stack.
The projections around the attention run on the library's tcgen05 GEMM (``B200MultiheadAttention``,
``B200Linear``); softmax(QK^T)V runs on the fused attention kernels (``mrg_attention_forward / _backward``)."""
from torch import nn

from ....attention import B200MultiheadAttention
from ....linear import B200Linear

from .residual_connection import ResidualConnection


class MultiModalAttentionBlockSequential(nn.Module):
    def __init__(self, modal1_feat_size=256, modal2_feat_size=256, num_head=1, dropout=0.0) -> None:
        super().__init__()
        self.cross_modal_att = B200MultiheadAttention(embed_dim=modal1_feat_size, num_heads=num_head,
                                                     dropout=dropout, batch_first=True,
                                                     kdim=modal2_feat_size, vdim=modal2_feat_size)
        self.projection = B200Linear(modal1_feat_size, modal1_feat_size)

    def forward(self, modal1, modal2):
        att, _ = self.cross_modal_att(query=modal1, key=modal2, value=modal2, need_weights=False)
        return self.projection(att)


class MultimodalAttentionBlock(nn.Module):
    def __init__(self, modal1_feat_size=256, modal2_feat_size=256, num_head=1, dropout=0.0,
                 use_residual=True, use_layer_norm=True) -> None:
        super().__init__()
        inner = MultiModalAttentionBlockSequential(modal1_feat_size, modal2_feat_size, num_head, dropout)
        self.att_module = ResidualConnection(inner, use_layer_norm, modal1_feat_size) if use_residual else inner

    def forward(self, modal1, modal2):
        return self.att_module(modal1, modal2)


class MultimodalAttention(nn.Module):
    def __init__(self, modal1_feat_size=256, modal2_feat_size=256, num_head=1, num_layers=1, dropout=0.0,
                 use_residual=True, use_layer_norm=True) -> None:
        super().__init__()
        self.att_layers = nn.ModuleList(
            MultimodalAttentionBlock(modal1_feat_size, modal2_feat_size, num_head, dropout, use_residual,
                                     use_layer_norm) for _ in range(num_layers))

    def forward(self, modal1, modal2):
        for layer in self.att_layers:
            modal1 = layer(modal1, modal2)
        return modal1
