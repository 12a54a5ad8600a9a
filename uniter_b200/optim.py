"""Fused multi-tensor AdamW with fp32 master weights on libub200.

Replaces, for 16-bit models on a B200, what the reference assembles from three pieces:

* ``optim/adamw.py:43-103`` — the AdamW update itself (bias-corrected step size; decoupled weight
  decay ``p -= lr * wd * p`` applied AFTER the Adam update);
* apex ``amp`` O2 (``train_vqa.py:152,190-192``) — fp32 master copies of the fp16 parameters,
  master-gradient copy + unscale before the step, master -> model copy after it;
* ``torch.nn.utils.clip_grad_norm_(amp.master_params(optimizer), grad_norm)``
  (``train_vqa.py:223-226``).

Here: ONE kernel computes the global gradient norm (``ub200_grad_sumsq``) and ONE kernel does
unscale + clip + Adam + decay + 16-bit refresh for every parameter (``ub200_adamw_step``); the clip
coefficient is read from device memory, so a step never synchronises with the host.  The
encoder-layer gradients are read in place from the flat gradient arena (right where the NCCL
all-reduce left them).

Same surface as the reference optimizer: ``param_groups`` (list of dicts with ``params``, ``lr``,
``weight_decay``, ``betas``, ``eps``, ``correct_bias``) that the training loop mutates
(``train_vqa.py:207-214``), ``step()``, ``zero_grad()``, ``state_dict()`` / ``load_state_dict()``;
``build_optimizer`` groups parameters by name exactly as ``optim/misc.py:14-22``.
"""
import ctypes as C
import math

import torch

from . import _lib


def build_optimizer(model, opts):
    """optim/misc.py:12-37: no weight decay for names containing 'bias', 'LayerNorm.bias',
    'LayerNorm.weight'; betas / lr from opts (only the 'adamw' branch exists here)."""
    param_optimizer = list(model.named_parameters())
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = [
        {"params": [p for n, p in param_optimizer if not any(nd in n for nd in no_decay)],
         "weight_decay": opts.weight_decay},
        {"params": [p for n, p in param_optimizer if any(nd in n for nd in no_decay)],
         "weight_decay": 0.0},
    ]
    if getattr(opts, "optim", "adamw") != "adamw":
        raise ValueError("invalid optimizer (libub200 implements adamw)")
    return FusedAdamW(groups, lr=opts.learning_rate, betas=tuple(opts.betas))


class FusedAdamW(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0,
                 correct_bias=True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{"params": params}]
        self.param_groups = []
        seen = set()
        for g in params:
            g = dict(g)
            g["params"] = [p for p in g["params"] if not (id(p) in seen or seen.add(id(p)))]
            for k, v in defaults.items():
                g.setdefault(k, v)
            self.param_groups.append(g)
        self.state = {}          # id(param) -> dict(step, master, exp_avg, exp_avg_sq)
        self._step_count = 0
        self._dev_tables = None  # (segs uint8 tensor, blk_start int32 tensor) on the device
        self._sumsq = None
        self.last_sumsq = None   # device scalar: sum of squares of the (scaled) gradients
        self._ring = _lib.PinnedRing(8)

    # ------------------------------------------------------------------ state
    def _init_state(self, p):
        st = self.state.get(id(p))
        if st is None:
            st = dict(step=0, master=p.detach().float().clone(),
                      exp_avg=torch.zeros(p.shape, device=p.device, dtype=torch.float32),
                      exp_avg_sq=torch.zeros(p.shape, device=p.device, dtype=torch.float32))
            self.state[id(p)] = st
        return st

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def state_dict(self):
        packed, idx = {}, 0
        groups = []
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                st = self.state.get(id(p))
                if st is not None:
                    packed[idx] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                ids.append(idx)
                idx += 1
            groups.append({k: (ids if k == "params" else v) for k, v in g.items()})
        return {"state": packed, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
            for p in g["params"]:
                st = sd["state"].get(idx)
                if st is not None:
                    self.state[id(p)] = {k: (v.to(p.device).clone() if torch.is_tensor(v) else v)
                                         for k, v in st.items()}
                idx += 1

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, grad_scale=1.0, max_grad_norm=-1.0):
        """One optimizer step over every parameter that has a gradient.

        grad_scale: the loss scale the gradients carry (they are multiplied by 1 / grad_scale);
        max_grad_norm > 0: clip the global norm of the unscaled gradients like
        ``clip_grad_norm_`` (the norm itself stays on the device: ``self.last_sumsq``)."""
        lib = _lib.load()
        chunk = lib.ub200_adam_chunk()
        segs, starts = [], [0]
        dev = None
        betas = eps = None
        for g in self.param_groups:
            b1, b2 = g["betas"]
            if betas is None:
                betas, eps = (b1, b2), g["eps"]
            elif (b1, b2) != betas or g["eps"] != eps:
                raise ValueError("FusedAdamW: betas / eps must be the same in every param group")
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdamW runs on CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters")
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self._init_state(p)
                st["step"] += 1
                step_size = g["lr"]
                if g["correct_bias"]:                                   # optim/adamw.py:82-86
                    bc1 = 1.0 - b1 ** st["step"]
                    bc2 = 1.0 - b2 ** st["step"]
                    step_size = step_size * math.sqrt(bc2) / bc1
                lr_wd = g["lr"] * g["weight_decay"] if g["weight_decay"] > 0.0 else 0.0
                segs.append(_lib.AdamSegment(
                    grad=grad.data_ptr(), master=st["master"].data_ptr(), exp_avg=st["exp_avg"].data_ptr(),
                    exp_avg_sq=st["exp_avg_sq"].data_ptr(), model=p.data_ptr(), n=p.numel(),
                    step_size=step_size, lr_wd=lr_wd,
                    grad_dtype=_lib.dtype_code(grad.dtype, allow_f32=True),
                    model_dtype=_lib.dtype_code(p.dtype, allow_f32=True)))
                segs[-1]._keep = grad
                starts.append(starts[-1] + (p.numel() + chunk - 1) // chunk)
                dev = p.device
        if not segs:
            return None
        nseg, nblocks = len(segs), starts[-1]
        arr = (_lib.AdamSegment * nseg)(*segs)
        seg_bytes = C.sizeof(arr)
        host = torch.frombuffer(bytearray(C.string_at(C.addressof(arr), seg_bytes)), dtype=torch.uint8)
        # pinned staging: a pageable H2D copy here would drain the stream (a sync per step)
        segs_dev = self._ring.upload(host, dev)
        starts_dev = self._ring.upload(torch.tensor(starts, dtype=torch.int32), dev)
        stream = _lib.current_stream()
        sumsq_ptr = None
        if max_grad_norm is not None and max_grad_norm > 0:
            self.last_sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
            _lib.check(lib.ub200_grad_sumsq(segs_dev.data_ptr(), starts_dev.data_ptr(), nseg, nblocks,
                                            self.last_sumsq.data_ptr(), stream))
            sumsq_ptr = self.last_sumsq.data_ptr()
        _lib.check(lib.ub200_adamw_step(segs_dev.data_ptr(), starts_dev.data_ptr(), nseg, nblocks,
                                        betas[0], betas[1], eps, 1.0 / float(grad_scale),
                                        float(max_grad_norm) if sumsq_ptr else -1.0, sumsq_ptr, stream))
        self._dev_tables = (segs_dev, starts_dev)   # keep alive until the kernels have run
        self._step_count += 1
        return None
