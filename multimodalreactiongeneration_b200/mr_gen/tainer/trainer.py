"""Lightning-free training runtime for the LSTM models (reference: mr_gen/model/<m>/trainer.py and the
work-in-progress mr_gen/tainer/trainer.py — hydra main -> ``pl.Trainer(strategy="ddp").fit``).

What Lightning did for the reference and what replaces it here:
  * one process per GPU + DDP gradient averaging  -> ``FlatGradBucket``: every ``param.grad`` is a view
    into ONE flat fp32 buffer, so autograd accumulates straight into it and a single
    ``all_reduce`` (NCCL over NVLink on GPU boxes, gloo in the CPU tests) averages all gradients;
  * ``training_step`` / ``configure_optimizers`` hooks -> called directly (same names on the models);
  * ModelCheckpoint -> ``save_checkpoint`` / ``load_checkpoint`` writing ``{"state_dict": ...}`` with the
    reference's keys (mr_gen/model/model_loader.py:23-24 reads exactly that).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, Optional

import torch
import torch.distributed as dist
from torch import nn


_ALIGN = 64  # floats: every parameter starts on a 256-byte boundary (TMA needs 16-byte aligned weights)


class FlatGradBucket:
    """All gradients of ``module`` in one contiguous fp32 buffer (the only collective of the path).

    On CUDA the PARAMETERS are re-homed into a second flat buffer with the same offsets, so that the optimizer
    step can run as one streamed kernel over (param, grad, exp_avg, exp_avg_sq) — ``FlatAdamW``.  The
    ``nn.Parameter`` objects, their names and shapes are untouched (``state_dict`` keys stay the reference's)."""

    def __init__(self, module: nn.Module, flatten_params: bool = True):
        self.params = [p for p in module.parameters() if p.requires_grad]
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_params = None
        if flatten_params and dev.type == "cuda" and all(p.dtype == torch.float32 for p in self.params):
            self.flat_params = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            p.grad = self.flat[o:o + n].view_as(p)
            if self.flat_params is not None:
                dst = self.flat_params[o:o + n].view_as(p)
                dst.copy_(p.data)
                p.data = dst
                # the optimizer kernel clears the bucket every step: weight-gradient kernels may add into it
                p._mrg_grad_fused = os.environ.get("MRG_FUSED_WGRAD", "1") != "0"

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce_sum(self) -> int:
        """SUM over the ranks; returns the world size (the mean's 1/world is folded into the optimizer)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            return dist.get_world_size()
        return 1

    def child_range(self, module: nn.Module, child_name: str):
        """[start, end) of the flat bucket that holds the gradients of ``module.<child_name>`` (parameters are laid
        out in registration order, so a sub-module owns one contiguous range)."""
        ids = {id(p) for p in getattr(module, child_name).parameters() if p.requires_grad}
        idx = [i for i, p in enumerate(self.params) if id(p) in ids]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            return None
        last = idx[-1]
        end = self.offsets[last + 1] if last + 1 < len(self.offsets) else self.flat.numel()
        return self.offsets[idx[0]], end

    def all_reduce_mean(self) -> None:
        world = self.all_reduce_sum()
        if world > 1:
            self.flat.mul_(1.0 / world)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


class FlatAdamW:
    """``torch.optim.AdamW`` arithmetic as ONE kernel over the flat buckets (C-ABI ``mrg_adamw_flat``).

    Hyper-parameters are read from the torch optimizer that the model's ``configure_optimizers`` returned
    (``param_groups[0]``) at every step, so LR schedulers that mutate it keep working; the step counter and the
    bias corrections live on the device (CUDA-graph replay advances them)."""

    def __init__(self, bucket: FlatGradBucket, optimizer: torch.optim.Optimizer):
        from ... import _cabi
        self._cabi = _cabi
        self.bucket = bucket
        self.optimizer = optimizer
        n = bucket.flat.numel()
        dev = bucket.flat.device
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(3, dtype=torch.float32, device=dev)

    @staticmethod
    def applicable(bucket: FlatGradBucket, optimizer) -> bool:
        if bucket.flat_params is None or type(optimizer) is not torch.optim.AdamW or len(optimizer.param_groups) != 1:
            return False
        g = optimizer.param_groups[0]
        same = len(g["params"]) == len(bucket.params) and all(a is b for a, b in zip(g["params"], bucket.params))
        return same and not g.get("amsgrad", False) and not g.get("maximize", False)

    def step(self, grad_scale: float = 1.0, zero_grad: bool = True) -> None:
        g = self.optimizer.param_groups[0]
        lr = g["lr"]
        lr_dev = lr.data_ptr() if torch.is_tensor(lr) else None
        b = self.bucket
        dev = b.flat.device
        with torch.cuda.device(dev):
            st = self._cabi.lib().mrg_adamw_flat(
                b.flat_params.data_ptr(), b.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                b.flat.numel(), lr_dev, 0.0 if lr_dev else float(lr), float(g["betas"][0]), float(g["betas"][1]),
                float(g["eps"]), float(g["weight_decay"]), self.state.data_ptr(), float(grad_scale),
                1 if zero_grad else 0, torch.cuda.current_stream(dev).cuda_stream)
        self._cabi.check(st, "mrg_adamw_flat")


def broadcast_parameters(module: nn.Module, src: int = 0) -> None:
    """DDP's constructor broadcast: every rank starts from rank ``src``'s weights."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


def build_optimizer(model: nn.Module):
    """``configure_optimizers`` of the reference models returns {"optimizer", "lr_scheduler": {...}}."""
    cfg = model.configure_optimizers()
    sched = cfg.get("lr_scheduler", {}).get("scheduler") if isinstance(cfg, dict) else None
    opt = cfg["optimizer"]
    on_gpu = any(p.is_cuda for g in opt.param_groups for p in g["params"])
    if on_gpu:  # keep the step counters on the device from the first step on (CUDA-graph capturable)
        for group in opt.param_groups:
            if "capturable" in group:
                group["capturable"] = True
    return opt, sched


class Trainer:
    """fit loop: zero bucket -> training_step -> backward -> one all-reduce -> optimizer step."""

    def __init__(self, model: nn.Module, max_epochs: int = 1, log_every: int = 0,
                 ckpt_dir: Optional[str] = None):
        self.model = model
        self.max_epochs = max_epochs
        self.log_every = log_every
        self.ckpt_dir = ckpt_dir
        broadcast_parameters(model)
        # the bucket must own the .grad tensors before the optimizer is created
        self.bucket = FlatGradBucket(model)
        self.optimizer, self.scheduler = build_optimizer(model)
        # AdamW on CUDA: one fused kernel over the flat buckets; anything else (SGD, CPU tests) steps through torch
        self.flat_opt = FlatAdamW(self.bucket, self.optimizer) if FlatAdamW.applicable(self.bucket, self.optimizer) \
            else None
        self.global_step = 0
        # Bucketed, overlapped gradient all-reduce: a model may name the sub-modules whose backward finishes early
        # (`ddp_overlap_children`, each used exactly ONCE per forward); their slice of the flat bucket is all-reduced
        # asynchronously as soon as their backward is done, while the rest of the backward still runs.  Everything
        # else goes out in one all-reduce after the backward, as before.
        self._early = []          # [(start, end)] in firing order
        self._early_done = []
        self._handles = []
        self._in_backward = False
        self.n_early_all_reduces = 0   # early (overlapped) all-reduces issued so far
        if _world() > 1 and os.environ.get("MRG_DDP_OVERLAP", "1") != "0":
            for name in getattr(model, "ddp_overlap_children", ()):
                rng = self.bucket.child_range(model, name)
                if rng is None:
                    continue
                k = len(self._early)
                self._early.append(rng)
                self._early_done.append(False)
                getattr(model, name).register_full_backward_hook(
                    lambda m, gin, gout, k=k: self._child_backward_done(k, gin))

    def _child_backward_done(self, k: int, grad_input) -> None:
        # fires when the gradients w.r.t. the sub-module's INPUTS exist, i.e. after every node inside it has run
        if not self._in_backward or self._early_done[k] or not any(g is not None for g in grad_input):
            return
        a, b = self._early[k]
        if self.bucket.flat.is_cuda:   # weight gradients of this slice may still be in flight on a side stream
            from ...lstm import join_wgrad_streams
            join_wgrad_streams()
        self._handles.append(dist.all_reduce(self.bucket.flat[a:b], op=dist.ReduceOp.SUM, async_op=True))
        self._early_done[k] = True
        self.n_early_all_reduces += 1

    def _all_reduce_rest(self) -> int:
        """SUM all-reduce of whatever the early hooks did not cover; returns the world size."""
        world = _world()
        if world <= 1:
            return 1
        flat = self.bucket.flat
        done = sorted(r for r, d in zip(self._early, self._early_done) if d)
        pos = 0
        for a, b in done + [(flat.numel(), flat.numel())]:
            if a > pos:
                self._handles.append(dist.all_reduce(flat[pos:a], op=dist.ReduceOp.SUM, async_op=True))
            pos = max(pos, b)
        for h in self._handles:
            h.wait()
        self._handles = []
        self._early_done = [False] * len(self._early)
        return world

    def forward_backward(self, batch) -> torch.Tensor:
        """``training_step`` + backward into the flat bucket (this rank's gradients, not yet averaged).  On CUDA the
        weight-gradient GEMMs of the LSTM layers run on side streams; every stream that wrote into the bucket
        (those side streams, the second encoder stream of SimpleLSTM) is joined explicitly before returning."""
        if self.flat_opt is None:
            self.bucket.zero()
        loss = self.model.training_step(batch)["loss"]
        self._in_backward = True
        try:
            if loss.is_cuda:
                from ...lstm import wgrad_overlap
                with wgrad_overlap():   # joins the side streams and every fused-write stream on exit
                    loss.backward()
            else:
                loss.backward()
        finally:
            self._in_backward = False
        return loss.detach()

    def train_step(self, batch) -> torch.Tensor:
        loss = self.forward_backward(batch)
        world = self._all_reduce_rest()
        if self.flat_opt is not None:   # mean = SUM all-reduce, 1/world folded into the step; grads cleared by it
            self.flat_opt.step(grad_scale=1.0 / world, zero_grad=True)
        else:
            if world > 1:
                self.bucket.flat.mul_(1.0 / world)
            self.optimizer.step()
        self.global_step += 1
        return loss

    # ------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole step (zero bucket -> fwd -> bwd -> all-reduce -> optimizer).
    # The step launches several hundred small kernels; replaying one captured graph removes the host
    # enqueue time (~19 ms at the bench shape) from the critical path.  Shapes must be static: the
    # batch is copied into fixed device buffers first.  Not usable for steps with host reads inside
    # (the wavefront rollout reads the wave sizes once per step).
    # ------------------------------------------------------------------------------------------
    def enable_cuda_graph(self, example_batch, warmup: int = 3) -> None:
        assert all(torch.is_tensor(t) and t.is_cuda for t in example_batch), "graphed step takes CUDA tensors"
        for group in self.optimizer.param_groups:      # the learning rate is read from device memory
            if not torch.is_tensor(group["lr"]):
                group["lr"] = torch.tensor(float(group["lr"]), dtype=torch.float32,
                                           device=example_batch[0].device)
        self._static_batch = tuple(torch.empty_like(t) for t in example_batch)
        for dst, src in zip(self._static_batch, example_batch):
            dst.copy_(src)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._static_loss = self.train_step(self._static_batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self.train_step(self._static_batch)
        self.global_step -= 1  # the capture pass did not execute

    def release_cuda_graph(self) -> None:
        """Drop the captured step.  NCCL keeps a reference on every communicator a live CUDA graph has captured
        collectives of, and tearing the process group down with such a graph alive can block (seen at 8 ranks)."""
        torch.cuda.synchronize()
        self._graph = None
        self._static_loss = None
        self._static_batch = None

    def close(self, destroy_process_group: bool = True, timeout_s: float = 30.0) -> None:
        """Orderly shutdown: drop the captured step (NCCL keeps communicators alive for a live graph), drain the
        device, meet the other ranks, then tear the process group down.  The teardown has been seen to block at 8
        ranks after graph-captured collectives; a watchdog ends the process (exit code 0, all results are out) if it
        does not return within ``timeout_s``."""
        if getattr(self, "_graph", None) is not None:
            self.release_cuda_graph()
        if not (dist.is_available() and dist.is_initialized()):
            return
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier()
        if not destroy_process_group:
            return
        import sys
        import threading
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(timeout_s, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()

    def train_step_graphed(self, batch) -> torch.Tensor:
        for dst, src in zip(self._static_batch, batch):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self.global_step += 1
        return self._static_loss

    @torch.no_grad()
    def validate(self, batches: Iterable) -> float:
        self.model.eval()
        losses = [float(self.model.validation_step(b)["loss"]) for b in batches]
        self.model.train()
        return sum(losses) / max(1, len(losses))

    def fit(self, train_batches: Callable[[int], Iterable], val_batches: Optional[Callable[[int], Iterable]] = None):
        history = []
        for epoch in range(self.max_epochs):
            self.model.current_epoch = epoch
            last = None
            for batch in train_batches(epoch):
                last = self.train_step(batch)
                if self.log_every and self.global_step % self.log_every == 0 and _rank() == 0:
                    print(f"epoch {epoch} step {self.global_step} train_loss {float(last):.6f}", flush=True)
            if self.scheduler is not None:
                self.scheduler.step()
            rec = {"epoch": epoch, "train_loss": None if last is None else float(last)}
            if val_batches is not None:
                rec["val_loss"] = self.validate(val_batches(epoch))
            history.append(rec)
            if self.ckpt_dir and _rank() == 0:
                save_checkpoint(self.model, os.path.join(self.ckpt_dir, "last.ckpt"), epoch, self.global_step, trainer=self)
        return history


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def save_checkpoint(model: nn.Module, path: str, epoch: int = 0, global_step: int = 0, trainer=None) -> None:
    """Lightning-shaped ``.ckpt``: a dict whose ``"state_dict"`` has the reference's keys.  With ``trainer`` the
    optimizer / LR-scheduler state is stored as Lightning does (``optimizer_states``, ``lr_schedulers``); the moments of
    the fused flat AdamW (which torch's ``optimizer.state_dict()`` does not see) go under ``flat_adamw``."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    ckpt = {"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "epoch": epoch, "global_step": global_step}
    if trainer is not None:
        ckpt["optimizer_states"] = [trainer.optimizer.state_dict()]
        ckpt["lr_schedulers"] = [trainer.scheduler.state_dict()] if trainer.scheduler is not None else []
        if trainer.flat_opt is not None:
            fo = trainer.flat_opt
            ckpt["flat_adamw"] = {"exp_avg": fo.exp_avg.detach().cpu(), "exp_avg_sq": fo.exp_avg_sq.detach().cpu(),
                                  "state": fo.state.detach().cpu(),
                                  "names": [n for n, _ in model.named_parameters() if _.requires_grad]}
    torch.save(ckpt, path)


def load_checkpoint(model: nn.Module, path: str, map_location="cpu", trainer=None, trusted: bool = True) -> Dict:
    """``trusted=True`` (a checkpoint you wrote, or a Lightning ``.ckpt`` of the reference, which pickles non-tensor
    objects) loads with ``weights_only=False``; pass ``trusted=False`` for files of unknown origin."""
    ckpt = torch.load(path, map_location=map_location, weights_only=not trusted)
    model.load_state_dict(ckpt["state_dict"])
    if trainer is not None:
        if ckpt.get("optimizer_states"):
            trainer.optimizer.load_state_dict(ckpt["optimizer_states"][0])
        if ckpt.get("lr_schedulers") and trainer.scheduler is not None:
            trainer.scheduler.load_state_dict(ckpt["lr_schedulers"][0])
        fa = ckpt.get("flat_adamw")
        if fa is not None and trainer.flat_opt is not None and fa["exp_avg"].numel() == trainer.flat_opt.exp_avg.numel():
            trainer.flat_opt.exp_avg.copy_(fa["exp_avg"])
            trainer.flat_opt.exp_avg_sq.copy_(fa["exp_avg_sq"])
            trainer.flat_opt.state.copy_(fa["state"])
        trainer.global_step = int(ckpt.get("global_step", trainer.global_step))
    return ckpt


def init_distributed(backend: Optional[str] = None) -> int:
    """One process per GPU, rendezvous from the torchrun environment."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return world
