"""``B200LSTM`` — the drop-in for ``torch.nn.LSTM`` at the reference's three seams.

Reference call sites (all ``batch_first=True``, ``proj_size=0``, fp32):
  mr_gen/model/utils/lstm_block.py:21,41     LSTMModule.lstm_module
  mr_gen/model/utils/lstm_sampler.py:16,29   LSTMSampler.sampler
  mr_gen/model/utils/mixer_block.py:237,251  LSTMMixer.mixer

The class keeps ``nn.LSTM``'s constructor, parameter names / shapes / init order (so ``state_dict``s
and seeded initialisations are interchangeable, torch/nn/modules/rnn.py:935-956, 308-311) and replaces
the arithmetic: ``forward`` queues the hand-written sm_100a kernels through the C-ABI in
``include/mrg_lstm.h``.  There is no cuDNN call and no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch
from torch import nn
from torch.nn import functional as F

from . import _cabi

_WORKSPACES = {}
_F_INFER = 1 << 30   # host-only flag bit of _LSTMLayerFn (never reaches the C-ABI): the call runs under torch.no_grad()
_PACK_CACHE = {}   # (weight storages, shape, flags) -> (weights, versions, w_pack, dedicated workspace): see _LSTMLayerFn.forward


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _default_flags() -> int:
    f = 0
    if os.environ.get("MRG_GENERIC_REC", "0") == "1":
        f |= _cabi.F_GENERIC_REC
    if os.environ.get("MRG_SIMT_GEMM", "0") == "1":
        f |= _cabi.F_SIMT_GEMM
    if _PRECISION["mode"] == "tf32":
        f |= _cabi.F_TF32
    elif _PRECISION["mode"] == "bf16":
        f |= _cabi.F_TF32 | _cabi.F_BF16
    f |= (_BUDGET["clusters"] & 0xFF) << 16
    return f


_PRECISION = {"mode": os.environ.get("MRG_PRECISION", "fp32")}
_BUDGET = {"clusters": 0}


class cluster_budget:
    """Context manager: the recurrent kernels launched inside use at most ``n`` thread-block clusters, so that two
    independent LSTM stacks running on two CUDA streams (e.g. the audio and the motion encoder of SimpleLSTM) fit on the
    GPU side by side instead of queueing behind each other.  The budget is stored with the autograd node, so the
    BPTT kernels of those layers use it too."""

    def __init__(self, n: int):
        self.n = int(n)

    def __enter__(self):
        self.prev = _BUDGET["clusters"]
        _BUDGET["clusters"] = self.n
        return self

    def __exit__(self, *exc):
        _BUDGET["clusters"] = self.prev
        return False


def set_precision(mode: str) -> None:
    """"fp32" (default): 3xTF32 tensor-core GEMMs, fp32-grade (parity <= 1e-5 / 1e-4).
    "tf32": one tensor-core pass per GEMM (10-bit mantissa, at least bf16 precision) — the reduced
    precision mode of BASELINE.json's north_star; stated bound: states 2e-2, loss/gradients 5e-2
    norm-relative (measured ~1e-3).
    "bf16": "tf32" plus bfloat16 STORAGE of the LSTM reserve — the x-projection, the saved gates and the
    d(pre-activations) that feed the backward GEMMs are 4 x bf16 per hidden unit (8 bytes instead of 16), the GEMMs
    next to them write / read bfloat16 directly (bf16 values are exact tf32 operands) — for the cluster kernels
    (H in {128, 256}, T > 1); other shapes run as in "tf32".  Same stated bound: states 2e-2, loss / gradients 5e-2
    norm-relative.  In both reduced modes an H = 256 layer also runs the recurrent products h W_hh^T / dpre W_hh as one tf32
    pass on the warp-level tensor cores (csrc/mrg_rec_fwd3.cu, mrg_rec_bwd3.cu); accumulation (TMEM / MMA accumulators),
    cell state and hidden states stay fp32 in every mode, and the "fp32" recurrence is exact fp32 (FFMA2)."""
    if mode not in ("fp32", "tf32", "bf16"):
        raise ValueError("precision must be 'fp32', 'tf32' or 'bf16'")
    _PRECISION["mode"] = mode
    if os.path.exists(_cabi.LIB_PATH):   # the attention kernels follow: 3xTF32 for fp32, one tf32 pass otherwise
        _cabi.check(_cabi.lib().mrg_attention_set_mode(0 if mode == "fp32" else 1), "mrg_attention_set_mode")


# Weight-gradient overlap: inside ``with wgrad_overlap():`` (the trainer wraps ``loss.backward()`` in it) a layer's backward
# is issued in two phases — BPTT + bias sums + dX on the node's own stream (what the previous layer's BPTT waits for), the
# two weight-gradient GEMMs on a side stream, where they share the GPU with that next BPTT kernel (a cluster kernel leaves
# 28 SMs idle) instead of delaying it.  Only when the gradients go straight into the trainer's flat bucket (nobody reads
# them before the join at the end of the context).
_WGRAD = {"on": False, "pending": [], "streams": {}, "keepalive": []}


def _wgrad_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    st = _WGRAD["streams"].get(key)
    if st is None:
        st = _WGRAD["streams"][key] = torch.cuda.Stream(device=dev)
    return st


def note_fused_write(dev: torch.device) -> None:
    """A kernel that adds into the trainer's flat gradient bucket behind autograd's back is being queued on the
    current stream of ``dev`` (``linear.fused_grad_target``): record the stream for ``join_wgrad_streams``."""
    st = torch.cuda.current_stream(dev)
    if all(st != s for s in _WGRAD["pending"]):
        _WGRAD["pending"].append(st)


def join_wgrad_streams(final: bool = False) -> None:
    """The current stream waits for every stream that has queued a write into the flat gradient bucket so far:
    the side streams of the weight-gradient GEMMs and any stream a fused weight-gradient kernel ran on (e.g. the
    second encoder stream of SimpleLSTM) — an explicit join, not the autograd engine's leaf-stream sync."""
    for side in _WGRAD["pending"]:
        cur = torch.cuda.current_stream(side.device)
        if cur != side:
            cur.wait_stream(side)
    _WGRAD["pending"] = []
    if final:   # end of the backward: no gradient accumulation of autograd is outstanding any more
        _WGRAD["keepalive"] = []


class wgrad_overlap:
    def __enter__(self):
        self.prev = _WGRAD["on"]
        _WGRAD["on"] = os.environ.get("MRG_WGRAD_OVERLAP", "1") != "0"
        return self

    def __exit__(self, *exc):
        _WGRAD["on"] = self.prev
        join_wgrad_streams(final=True)
        return False


class _LSTMLayerFn(torch.autograd.Function):
    """One nn.LSTM layer (1 or 2 directions), time-major, through the C-ABI."""

    @staticmethod
    def forward(ctx, x, h0, c0, D, H, flags, *weights):
        # weights: per direction (w_ih, w_hh, b_ih | None, b_hh | None)
        L = _cabi.lib()
        if not x.is_cuda:
            raise RuntimeError("B200LSTM has no CPU path: inputs must live on a B200 (sm_100a) device")
        if x.dtype != torch.float32:
            raise TypeError(f"B200LSTM computes in fp32; got {x.dtype}")
        x = _cabi.contiguous3(x)
        T, B, I = x.shape
        dev = x.device
        # needs_input_grad is True for the parameters even under torch.no_grad(); the caller (B200LSTM.forward) reads the grad
        # mode where it is visible and passes it as a host-only flag bit
        infer = bool(flags & _F_INFER)
        flags &= ~_F_INFER
        need_grad = any(ctx.needs_input_grad) and not infer
        opts = dict(dtype=torch.float32, device=dev)
        if flags & _cabi.F_BF16 and not (T > 1 and H in (128, 256) and not flags & _cabi.F_GENERIC_REC):
            flags &= ~_cabi.F_BF16      # bf16 reserve: cluster kernels only; other shapes keep the fp32 reserve
        gates = torch.empty((D, T, B, H, 4), dtype=torch.bfloat16 if flags & _cabi.F_BF16 else torch.float32, device=dev)
        y_ext = torch.empty((D, T + 1, B, H), **opts)
        c_ext = torch.empty((D, T + 1, B, H), **opts)
        nbytes = L.mrg_lstm_workspace_bytes(T, B, I, H, D)
        # Inference with frozen weights (no gradient wanted, plain LSTM parameters): the weight packs of a (layer, shape) are
        # written once into buffers this cache owns and reused until a weight changes (version counter / storage) —
        # frame-by-frame generation otherwise re-packs every layer on every frame (StreamingGenerator: 5 launches of 3 us).
        pack_key = pack_hit = None
        if not need_grad and not (flags & _cabi.F_GRU):
            pack_key = (tuple(0 if t is None else t.data_ptr() for t in weights), T, B, I, H, D, flags, dev.index)
            ent = _PACK_CACHE.get(pack_key)
            # same storages (the key; the entry keeps its tensors alive, so an address cannot be recycled) and unchanged
            # version counters (shared by every alias of a parameter: optimizer steps and load_state_dict bump them)
            if ent is not None and all(a is None or a._version == v for a, v in zip(weights, ent[1])):
                pack_hit = ent
        if pack_hit is not None:
            w_pack, ws = pack_hit[2], pack_hit[3]
            flags |= _cabi.F_PACK_VALID
        elif pack_key is not None:
            w_pack = torch.empty((L.mrg_lstm_pack_floats(I, H, D),), **opts)
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
            if len(_PACK_CACHE) >= 64:
                _PACK_CACHE.clear()
            _PACK_CACHE[pack_key] = (tuple(weights), tuple(0 if t is None else t._version for t in weights), w_pack, ws)
        else:
            w_pack = torch.empty((L.mrg_lstm_pack_floats(I, H, D),), **opts)   # W_ih: raw | tf32 hi | lo planes
            ws = _workspace(dev, nbytes)
        dw = (_cabi.DirWeights * D)()
        keep = []
        for d in range(D):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
            for t in (w_ih, w_hh, b_ih, b_hh):
                if t is not None and (not t.is_contiguous() or t.dtype != torch.float32 or t.device != dev):
                    raise ValueError("B200LSTM weights must be contiguous fp32 tensors on the input's device")
            h0d = None if h0 is None else h0[d].contiguous()
            c0d = None if c0 is None else c0[d].contiguous()
            keep += [h0d, c0d]
            dw[d] = _cabi.DirWeights(_cabi.ptr(w_ih), _cabi.ptr(w_hh), _cabi.ptr(b_ih), _cabi.ptr(b_hh),
                                     _cabi.ptr(h0d), _cabi.ptr(c0d))
        if h0 is None and c0 is None:
            flags |= _cabi.F_ZERO_STATE
        fl = flags | (_cabi.F_TRAIN if need_grad else 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            st = L.mrg_lstm_layer_forward(x.data_ptr(), dw, w_pack.data_ptr(), gates.data_ptr(),
                                          y_ext.data_ptr(), c_ext.data_ptr(), ws.data_ptr(), ws.numel(),
                                          T, B, I, H, D, fl, stream)
        _cabi.check(st, "mrg_lstm_layer_forward")
        if D == 1:
            y = y_ext[0, 1:]
            h_n = y_ext[:, T]
            c_n = c_ext[:, T]
        else:
            y = torch.cat([y_ext[0, 1:], y_ext[1, :T]], dim=-1)
            h_n = torch.stack([y_ext[0, T], y_ext[1, 0]])
            c_n = torch.stack([c_ext[0, T], c_ext[1, 0]])
        if need_grad:
            ctx.saved = (x, gates, y_ext, c_ext, w_pack, weights, h0 is not None, c0 is not None)
            ctx.dims = (T, B, I, H, D, flags)
            ctx.consumed = False
            # version-counter guard: with one direction y / h_n / c_n are VIEWS of the buffers the backward reads (y_ext,
            # c_ext) — an in-place edit of an output (or of the input) between forward and backward must raise, not
            # silently corrupt the gradients
            ctx.save_for_backward(*((x, y, c_n) if D == 1 else (x,)))
        return y, h_n, c_n

    @staticmethod
    def backward(ctx, dy, dh_n, dc_n):
        if ctx.consumed:
            raise RuntimeError("B200LSTM backward ran twice on the same graph: the reserve (gates) is "
                               "overwritten in place by d(pre-activations); retain_graph is not supported")
        ctx.consumed = True
        _ = ctx.saved_tensors   # raises if x / y / c_n were modified in place since the forward
        L = _cabi.lib()
        x, gates, y_ext, c_ext, w_pack, weights, has_h0, has_c0 = ctx.saved
        T, B, I, H, D, flags = ctx.dims
        dev = x.device
        opts = dict(dtype=torch.float32, device=dev)
        dy = None if dy is None else _cabi.contiguous3(dy)
        dh_n = None if dh_n is None else dh_n.contiguous()
        dc_n = None if dc_n is None else dc_n.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dh0 = torch.empty((D, B, H), **opts) if has_h0 and ctx.needs_input_grad[1] else None
        dc0 = torch.empty((D, B, H), **opts) if has_c0 and ctx.needs_input_grad[2] else None
        dw = (_cabi.DirWeights * D)()
        dg = (_cabi.DirGrads * D)()
        grads = []
        for d in range(D):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
            dw[d] = _cabi.DirWeights(_cabi.ptr(w_ih), _cabi.ptr(w_hh), _cabi.ptr(b_ih), _cabi.ptr(b_hh),
                                     None, None)
            g_ih = torch.empty_like(w_ih)
            g_hh = torch.empty_like(w_hh)
            g_b = torch.empty((4 * H,), **opts) if (b_ih is not None or b_hh is not None) else None
            dg[d] = _cabi.DirGrads(g_ih.data_ptr(), g_hh.data_ptr(), _cabi.ptr(g_b),
                                   None if dh0 is None else dh0[d].data_ptr(),
                                   None if dc0 is None else dc0[d].data_ptr())
            grads += [g_ih, g_hh, g_b if b_ih is not None else None, g_b if b_hh is not None else None]
        # weight gradients straight into the trainer's flat bucket (linear.fused_grad_target) when every weight
        # matrix of the layer allows it: one launch per parameter less, no temporaries
        from .linear import fused_grad_target
        tg = [fused_grad_target(weights[4 * d + i]) for d in range(D) for i in (0, 1)]
        fused = all(t is not None for t in tg)
        if fused:
            for d in range(D):
                dg[d].dw_ih, dg[d].dw_hh = tg[2 * d].data_ptr(), tg[2 * d + 1].data_ptr()
                grads[4 * d], grads[4 * d + 1] = None, None
            flags |= _cabi.F_ACC_WEIGHTS
        nbytes = L.mrg_lstm_workspace_bytes(T, B, I, H, D)
        ws = _workspace(dev, nbytes)
        main = torch.cuda.current_stream(dev)
        overlap = fused and _WGRAD["on"] and T > 1
        with torch.cuda.device(dev):
            st = L.mrg_lstm_layer_backward(x.data_ptr(), dw, w_pack.data_ptr(), _cabi.ptr(dy),
                                           _cabi.ptr(dh_n), _cabi.ptr(dc_n), gates.data_ptr(),
                                           y_ext.data_ptr(), c_ext.data_ptr(), _cabi.ptr(dx), dg,
                                           ws.data_ptr(), ws.numel(), T, B, I, H, D,
                                           flags | (_cabi.F_BWD_NO_WGRAD if overlap else 0), main.cuda_stream)
        _cabi.check(st, "mrg_lstm_layer_backward")
        if overlap:
            side = _wgrad_stream(dev)
            side.wait_stream(main)
            with torch.cuda.device(dev), torch.cuda.stream(side):
                ws2 = _workspace(dev, nbytes)   # keyed by the current (side) stream: its own split-K partials
                st = L.mrg_lstm_layer_backward(x.data_ptr(), dw, w_pack.data_ptr(), None, None, None, gates.data_ptr(),
                                               y_ext.data_ptr(), c_ext.data_ptr(), None, dg, ws2.data_ptr(), ws2.numel(),
                                               T, B, I, H, D, flags | _cabi.F_BWD_WGRAD_ONLY, side.cuda_stream)
            _cabi.check(st, "mrg_lstm_layer_backward (weight gradients)")
            for t in (x, gates, y_ext):   # the allocator must not hand these out again before the side stream is done
                t.record_stream(side)
            if side not in _WGRAD["pending"]:
                _WGRAD["pending"].append(side)
        ctx.saved = None
        return (dx, dh0, dc0, None, None, None, *grads)


def lstm_layer(x_tm: torch.Tensor, weights, hidden_size: int, directions: int = 1,
               h0: Optional[torch.Tensor] = None, c0: Optional[torch.Tensor] = None,
               flags: Optional[int] = None):
    """Functional form: x_tm [T,B,I] time-major -> (y [T,B,D*H], h_n [D,B,H], c_n [D,B,H])."""
    if flags is None:
        flags = _default_flags()
    return _LSTMLayerFn.apply(x_tm, h0, c0, directions, hidden_size, flags, *weights)


class B200LSTM(nn.LSTM):
    """``torch.nn.LSTM`` with its arithmetic replaced by the sm_100a kernels.

    Subclassing keeps the constructor signature, parameter registration order (hence seeded init),
    ``state_dict`` keys and ``isinstance(m, nn.LSTM)`` identical to the reference's module; ``forward``
    and ``flatten_parameters`` (a cuDNN weight-layout call) are replaced."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.proj_size != 0:
            raise NotImplementedError("B200LSTM: proj_size != 0 is not used by the reference and not built")

    def flatten_parameters(self) -> None:  # no cuDNN weight buffer
        return

    def _layer_weights(self, layer: int):
        out = []
        for d in range(2 if self.bidirectional else 1):
            sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
            out += [getattr(self, "weight_ih" + sfx), getattr(self, "weight_hh" + sfx),
                    getattr(self, "bias_ih" + sfx) if self.bias else None,
                    getattr(self, "bias_hh" + sfx) if self.bias else None]
        return out

    def forward(self, input, hx=None):  # noqa: A002 - nn.LSTM's argument name
        if isinstance(input, nn.utils.rnn.PackedSequence):
            raise NotImplementedError("B200LSTM: PackedSequence input is not used by the reference")
        if input.dim() not in (2, 3):
            raise ValueError(f"LSTM: Expected input to be 2D or 3D, got {input.dim()}D instead")
        batched = input.dim() == 3
        if not batched:
            input = input.unsqueeze(1 if not self.batch_first else 0)
            if hx is not None:
                hx = (hx[0].unsqueeze(1), hx[1].unsqueeze(1))
        x = input.transpose(0, 1) if self.batch_first else input
        if x.shape[-1] != self.input_size:
            raise RuntimeError(f"input.size(-1) must be equal to input_size. Expected {self.input_size}, "
                               f"got {x.shape[-1]}")
        D = 2 if self.bidirectional else 1
        Hs = self.hidden_size
        B = x.shape[1]
        if hx is not None:
            h0, c0 = hx
            want = (self.num_layers * D, B, Hs)
            if tuple(h0.shape) != want or tuple(c0.shape) != want:
                raise RuntimeError(f"Expected hidden size {want}, got {tuple(h0.shape)} / {tuple(c0.shape)}")
        flags = _default_flags()
        if not torch.is_grad_enabled():
            flags |= _F_INFER     # inference: no reserve for a backward, cached weight packs
        hs, cs = [], []
        for layer in range(self.num_layers):
            h0l = None if hx is None else hx[0][layer * D:(layer + 1) * D]
            c0l = None if hx is None else hx[1][layer * D:(layer + 1) * D]
            x, h_n, c_n = _LSTMLayerFn.apply(x, h0l, c0l, D, Hs, flags, *self._layer_weights(layer))
            hs.append(h_n)
            cs.append(c_n)
            if self.dropout > 0 and self.training and layer + 1 < self.num_layers:
                x = F.dropout(x, self.dropout, True)
        h_n = torch.cat(hs, dim=0) if len(hs) > 1 else hs[0]
        c_n = torch.cat(cs, dim=0) if len(cs) > 1 else cs[0]
        out = x.transpose(0, 1) if self.batch_first else x
        if not batched:
            out = out.squeeze(1 if not self.batch_first else 0)
            h_n, c_n = h_n.squeeze(1), c_n.squeeze(1)
        return out, (h_n, c_n)
