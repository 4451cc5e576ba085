"""Python handle around the C ABI: holds PyTorch tensors, calls the CUDA library.

``DecodeEngine`` is what the plug-in classes (segmentors/ddp.py, depth/ddp.py) drive.  PyTorch is
used for device memory (inputs, outputs, workspace), the current stream and torch.distributed —
no tensor math happens here.
"""
import ctypes
from typing import Mapping, Optional

import torch

from . import _lib as L
from . import schedule as S

EMBED = 256


def _fptr(seq):
    return (ctypes.c_float * len(seq))(*seq)


class DecodeEngine:
    """One configured decode loop (seg or depth) bound to one CUDA device."""

    def __init__(self, task="seg", num_classes=19, timesteps=3, time_difference=1, sample_range=(0, 0.999),
                 noise_schedule="cosine", diffusion="ddim", accumulation=False, bit_scale=0.01,
                 learned_sinusoidal_dim=16, num_layers=6, min_depth=1e-3, max_depth=10.0,
                 gemm_mode="fp32", device=None, host_schedule=True):
        if noise_schedule not in ("cosine", "linear"):
            raise ValueError(f"invalid noise schedule {noise_schedule}")          # ddp.py:90
        if diffusion not in ("ddim", "ddpm"):
            raise NotImplementedError(diffusion)                                   # ddp.py:123
        if gemm_mode not in L.GEMM_MODES:
            raise ValueError(f"gemm_mode must be one of {sorted(L.GEMM_MODES)}")
        if not torch.cuda.is_available():
            raise RuntimeError("ddp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self.task = task
        self.num_classes = num_classes if task == "seg" else 1
        self.timesteps = timesteps
        self.time_difference = time_difference
        self.sample_range = tuple(sample_range)
        self.noise_schedule = noise_schedule
        self.accumulation = bool(accumulation)
        self.diffusion = diffusion
        self.cin = EMBED if task == "seg" else 1
        self.cfg = L.DDPConfig(
            abi_version=L.ABI_VERSION, task=L.TASK_SEG if task == "seg" else L.TASK_DEPTH,
            num_classes=num_classes, timesteps=timesteps, time_difference=time_difference,
            noise_schedule=L.SCHEDULE_COSINE if noise_schedule == "cosine" else L.SCHEDULE_LINEAR,
            diffusion=L.DIFFUSION_DDIM if diffusion == "ddim" else L.DIFFUSION_DDPM,
            accumulation=int(bool(accumulation)), learned_sinusoidal_dim=learned_sinusoidal_dim,
            num_layers=num_layers, gemm_mode=L.GEMM_MODES[gemm_mode], sample_range_lo=float(sample_range[0]),
            bit_scale=float(bit_scale), min_depth=float(min_depth), max_depth=float(max_depth))
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.ddp_create(ctypes.byref(self.cfg), ctypes.byref(self._h))
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_last_error(None).decode())
        self._plan = None
        self._ws = None
        self._keep = []
        if host_schedule:
            self.use_reference_schedule()

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.ddp_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def weight_names(self):
        out = {}
        n = ctypes.c_int64()
        for i in range(self.lib.ddp_weight_count(self._h)):
            name = self.lib.ddp_weight_name(self._h, i, ctypes.byref(n))
            out[name.decode()] = n.value
        return out

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]):
        """Take the hot-path entries of a reference checkpoint's state dict (extra keys are ignored)."""
        need = self.weight_names()
        missing = [k for k in need if k not in sd]
        if missing:
            raise KeyError(f"state dict lacks hot-path weights: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k, numel in need.items():
            t = sd[k].detach().to("cpu", torch.float32).contiguous()
            if t.numel() != numel:
                raise ValueError(f"{k}: expected {numel} elements, got {tuple(t.shape)}")
            self._check(self.lib.ddp_set_weight(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), numel))
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_commit_weights(self._h))
        self._plan = None

    def use_reference_schedule(self):
        """Hand the library schedule scalars computed with the reference's own torch ops (see schedule.py)."""
        T = self.timesteps
        if self.task == "seg":
            l, a, s, an, sn = S.seg_schedule(T, self.time_difference, self.sample_range, self.noise_schedule)
            self._check(self.lib.ddp_set_schedule(self._h, T, _fptr(l), _fptr(a), _fptr(s), _fptr(an), _fptr(sn)))
            if self.diffusion == "ddpm":
                omc, c, std, on = S.seg_ddpm_schedule(T, self.time_difference, self.sample_range, self.noise_schedule)
                self._check(self.lib.ddp_set_ddpm_schedule(self._h, T, _fptr(omc), _fptr(c), _fptr(std),
                                                           (ctypes.c_int32 * T)(*on)))
        else:
            t, g, gn = S.depth_schedule(T, self.time_difference)
            self._check(self.lib.ddp_set_schedule(self._h, T, _fptr(t), _fptr(g), None, _fptr(gn), None))

    def get_schedule(self):
        T = self.timesteps
        arrs = [(ctypes.c_float * T)() for _ in range(5)]
        self._check(self.lib.ddp_get_schedule(self._h, *arrs))
        return [list(a) for a in arrs]

    def plan(self, B, R, h, w):
        key = (B, R, h, w)
        if self._plan == key:
            return
        nbytes = ctypes.c_size_t()
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_plan(self._h, B, R, h, w, ctypes.byref(nbytes)))
            if self._ws is None or self._ws.numel() < nbytes.value + 256:      # a smaller plan reuses the workspace
                self._ws = None
                self._ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
        self._ws_bytes = nbytes.value
        self._plan = key

    def _ws_ptr(self):
        p = self._ws.data_ptr()
        return (p + 255) // 256 * 256

    # ------------------------------------------------------------------ the hot path
    def sample(self, x: torch.Tensor, noise: torch.Tensor, return_cls=False, step_noise: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, return_uncertainty=False):
        """x (B,256,h,w), noise (B,R,Cin,h,w): CUDA fp32 tensors -> out (B,C,h,w) [, cls (B,h,w) int32].
        diffusion='ddpm' also needs step_noise (T,B,R,256,h,w): what the reference draws with randn_like every step.
        `out`: optional preallocated result buffer (stable addresses are what the DDP_B200_GRAPH=1 latency mode keys on).

        return_uncertainty: also return {"changes": (B,h,w) int32 (seg), "spread": (B,h,w) fp32} — per-pixel class-change
        counts over steps and samples, and the disagreement of the R samples at the last step (include/ddp_b200.h).

        Asynchronous on torch's current stream."""
        if self.diffusion == "ddpm":
            if step_noise is None:
                raise ValueError("diffusion='ddpm' needs step_noise (T,B,R,256,h,w)")
            step_noise = step_noise.contiguous()
            assert step_noise.is_cuda and tuple(step_noise.shape) == (self.timesteps,) + tuple(noise.shape)
            self._step_noise = step_noise
            self._check(self.lib.ddp_set_step_noise(self._h, step_noise.data_ptr()))
        B, c, h, w = x.shape
        R = noise.shape[1]
        assert c == EMBED and tuple(noise.shape) == (B, R, self.cin, h, w), (x.shape, noise.shape)
        assert x.is_cuda and noise.is_cuda and x.dtype == torch.float32 and noise.dtype == torch.float32
        self._on_device(x, noise, out)
        x = x.contiguous()
        noise = noise.contiguous()
        if B == 0:                      # empty batch: nothing to launch (the C ABI itself rejects B < 1)
            out = torch.empty((0, self.num_classes, h, w), dtype=torch.float32, device=x.device)
            return (out, torch.empty((0, h, w), dtype=torch.int32, device=x.device)) if (return_cls and self.task == "seg") else out
        self.plan(B, R, h, w)
        if out is None:
            out = torch.empty((B, self.num_classes, h, w), dtype=torch.float32, device=x.device)
        elif tuple(out.shape) != (B, self.num_classes, h, w) or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous CUDA fp32 tensor of shape {(B, self.num_classes, h, w)}")
        cls = torch.empty((B, h, w), dtype=torch.int32, device=x.device) if (return_cls and self.task == "seg") else None
        unc = None
        if return_uncertainty:
            unc = {"spread": torch.empty((B, h, w), dtype=torch.float32, device=x.device)}
            if self.task == "seg":
                unc["changes"] = torch.empty((B, h, w), dtype=torch.int32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_set_uncertainty_outputs(
                self._h, unc["changes"].data_ptr() if unc and "changes" in unc else None, unc["spread"].data_ptr() if unc else None))
            self._check(self.lib.ddp_sample(self._h, x.data_ptr(), noise.data_ptr(), out.data_ptr(),
                                            cls.data_ptr() if cls is not None else None,
                                            self._ws_ptr(), self._ws_bytes, stream))
        res = (out, cls) if return_cls else out
        if return_uncertainty:
            return (res + (unc,)) if return_cls else (res, unc)
        return res

    def head_forward(self, feat: torch.Tensor, time_embedding: torch.Tensor):
        """One denoiser call: feat (rows,256,h,w) CUDA fp32, time_embedding (1024,) -> (rows,C,h,w) logits / depth.
        rows must be B*R of the current plan (plan(rows, 1, h, w) is made if there is none that fits)."""
        rows, c, h, w = feat.shape
        assert c == EMBED and feat.is_cuda and feat.dtype == torch.float32
        self._on_device(feat)
        if self._plan is None or self._plan[0] * self._plan[1] != rows or self._plan[2:] != (h, w):
            self.plan(rows, 1, h, w)
        feat = feat.contiguous()
        temb = time_embedding.detach().to(feat.device, torch.float32).reshape(-1).contiguous()
        assert temb.numel() == 4 * EMBED
        out = torch.empty((rows, self.num_classes, h, w), dtype=torch.float32, device=feat.device)
        stream = torch.cuda.current_stream(feat.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_head_forward(self._h, feat.data_ptr(), temb.data_ptr(), out.data_ptr(),
                                                  self._ws_ptr(), self._ws_bytes, stream))
        return out

    def resize_argmax(self, logits: torch.Tensor, size):
        """(B,C,h,w) logits -> uint8 class map (B,H,W): bilinear resize + softmax + argmax in one kernel."""
        B, C, h, w = logits.shape
        H, W = int(size[0]), int(size[1])
        self._on_device(logits)
        logits = logits.contiguous()
        cls = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device)
        stream = torch.cuda.current_stream(logits.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_resize_argmax(self._h, logits.data_ptr(), B, C, h, w, H, W, cls.data_ptr(), stream))
        return cls

    def tail_probs(self, logits: torch.Tensor, img_size, crop=None, out_size=None, flip=None, accum: Optional[torch.Tensor] = None):
        """inference() tail of one view in one kernel: (B,C,h,w) logits -> resize to `img_size` -> [crop to `crop`, resize to
        `out_size`] -> softmax -> flip back ('horizontal' / 'vertical') -> (B,C,H,W) probabilities; added into `accum` when
        given (aug_test), else returned in a new tensor."""
        B, C, h, w = logits.shape
        self._on_device(logits, accum)
        logits = logits.contiguous()
        ih, iw = int(img_size[0]), int(img_size[1])
        rescale = out_size is not None
        ch, cw = (int(crop[0]), int(crop[1])) if crop is not None else (ih, iw)
        H, W = (int(out_size[0]), int(out_size[1])) if rescale else (ih, iw)
        fl = {None: 0, False: 0, "horizontal": 1, "vertical": 2}[flip]
        probs = accum if accum is not None else torch.empty((B, C, H, W), dtype=torch.float32, device=logits.device)
        assert tuple(probs.shape) == (B, C, H, W) and probs.dtype == torch.float32 and probs.is_contiguous()
        stream = torch.cuda.current_stream(logits.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_tail_probs(self._h, logits.data_ptr(), B, C, h, w, ih, iw, ch, cw, H, W, int(rescale), fl,
                                                int(accum is not None), probs.data_ptr(), stream))
        return probs

    def probs_argmax(self, probs: torch.Tensor):
        B, C, H, W = probs.shape
        self._on_device(probs)
        cls = torch.empty((B, H, W), dtype=torch.uint8, device=probs.device)
        stream = torch.cuda.current_stream(probs.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_probs_argmax(self._h, probs.contiguous().data_ptr(), B, C, H, W, cls.data_ptr(), stream))
        return cls

    def sample_host(self, x: torch.Tensor, noise: torch.Tensor, out: Optional[torch.Tensor] = None,
                    cls: Optional[torch.Tensor] = None, out_device: Optional[torch.Tensor] = None, chunks: int = 0):
        """Host tensors in (ideally pinned), host tensors out; the copies are part of the call and are pipelined against
        the loop in `chunks` groups of images (0 = automatic).  `out_device`: optional CUDA (B,C,h,w) tensor that also
        receives the result (the send buffer of the NCCL gather)."""
        B, c, h, w = x.shape
        R = noise.shape[1]
        assert not x.is_cuda and not noise.is_cuda
        x = x.contiguous()
        noise = noise.contiguous()
        self.plan(B, R, h, w)
        if out is None:
            out = torch.empty((B, self.num_classes, h, w), dtype=torch.float32).pin_memory()
        if out_device is not None:
            self._on_device(out_device)
            assert tuple(out_device.shape) == (B, self.num_classes, h, w) and out_device.dtype == torch.float32 and out_device.is_contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_sample_host_ex(self._h, x.data_ptr(), noise.data_ptr(), out.data_ptr(),
                                                    cls.data_ptr() if cls is not None else None,
                                                    out_device.data_ptr() if out_device is not None else None, int(chunks),
                                                    self._ws_ptr(), self._ws_bytes, stream))
        return out

    def submit_host(self, x: torch.Tensor, noise: torch.Tensor, out: torch.Tensor, cls: Optional[torch.Tensor] = None,
                    out_device: Optional[torch.Tensor] = None) -> int:
        """Streaming form of sample_host: returns a ticket at once, at most two calls in flight; wait_host(ticket) blocks
        until `out` (pinned host tensor) is complete.  The upload of the next call overlaps this call's loop."""
        B, c, h, w = x.shape
        R = noise.shape[1]
        assert not x.is_cuda and not noise.is_cuda and x.is_contiguous() and noise.is_contiguous()
        assert not out.is_cuda and tuple(out.shape) == (B, self.num_classes, h, w) and out.is_contiguous()
        self.plan(B, R, h, w)
        if out_device is not None:
            self._on_device(out_device)
        ticket = ctypes.c_int64()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_sample_host_submit(self._h, x.data_ptr(), noise.data_ptr(), out.data_ptr(),
                                                        cls.data_ptr() if cls is not None else None,
                                                        out_device.data_ptr() if out_device is not None else None,
                                                        self._ws_ptr(), self._ws_bytes, stream, ctypes.byref(ticket)))
        return ticket.value

    def wait_host(self, ticket: int):
        self._check(self.lib.ddp_sample_host_wait(self._h, ticket))

    @property
    def last_launch_count(self):
        return int(self.lib.ddp_last_launch_count(self._h))

    @property
    def graph_replays(self):
        """DDP_B200_GRAPH=1 latency mode: sample() calls served by replaying the captured CUDA graph so far."""
        return int(self.lib.ddp_graph_replays(self._h))

    @property
    def graph_captures(self):
        return int(self.lib.ddp_graph_captures(self._h))

    @property
    def graph_last_fallback(self):
        """Why the last sample() used ordinary launches ('' when it replayed the graph)."""
        return self.lib.ddp_graph_last_fallback(self._h).decode()

    def _on_device(self, *tensors):
        """The handle is bound to self.device: reject tensors of another GPU instead of launching into the wrong context."""
        for t in tensors:
            if t is not None and t.device != self.device:
                raise ValueError(f"tensor on {t.device}, this engine is bound to {self.device}")

    def profile(self, on=True):
        """Bracket every kernel launch of sample() with CUDA events (per kernel class)."""
        self._check(self.lib.ddp_profile_enable(self._h, int(on)))

    def profile_collect(self):
        """-> {class name: (total ms, launches)} since the last collect (synchronises on the events)."""
        ms = (ctypes.c_float * L.K_COUNT)()
        cnt = (ctypes.c_int64 * L.K_COUNT)()
        self._check(self.lib.ddp_profile_collect(self._h, ms, cnt, L.K_COUNT))
        return {self.lib.ddp_kernel_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(L.K_COUNT)}

    # ------------------------------------------------------------------ test hooks
    def add_tap(self, kind, step, layer, width_or_numel, per_token=True):
        B, R, h, w = self._plan
        n = B * R * h * w * width_or_numel if per_token else width_or_numel
        buf = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._check(self.lib.ddp_add_tap(self._h, kind, step, layer, buf.data_ptr()))
        self._keep.append(buf)
        return buf

    def set_state_override(self, step, state: torch.Tensor):
        state = state.contiguous()
        self._keep.append(state)
        self._check(self.lib.ddp_set_state_override(self._h, step, state.data_ptr()))

    def clear_debug(self):
        self._check(self.lib.ddp_clear_debug(self._h))
        self._keep = []
