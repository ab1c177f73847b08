"""GroundingEngine — Python owner of one `vgqa_ctx` (include/vgqa_b200.h).

PyTorch is used only for device memory and streams; all math runs in libvgqa_b200.so.  There is no CPU or
eager-PyTorch fallback: if the library or a B200 is missing the constructor raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib
from ._lib import c_float, c_int, c_void_p


class VgqaConfig(ctypes.Structure):
    _fields_ = [(n, c_int) for n in (
        "enc_layers", "dec_layers", "hidden", "heads", "ffn_dim", "app_num", "mot_num", "max_video_len",
        "max_clips", "max_frames", "max_hw", "max_text", "use_cuda_graph")]


class VgqaInputs(ctypes.Structure):
    _fields_ = [("clips", c_int), ("T", c_int), ("H", c_int), ("W", c_int), ("L", c_int),
                ("vis", c_void_p), ("vid", c_void_p), ("text", c_void_p), ("pos", c_void_p), ("pos_frames", c_int),
                ("vis_mask", c_void_p), ("text_mask", c_void_p), ("ori_sizes_hw", c_void_p),
                ("force_choose1", c_void_p), ("force_choose2", c_void_p), ("iteration_rate", c_int),
                ("stop_after_encoder", c_int),
                ("vis_raw", c_void_p), ("vid_raw", c_void_p), ("text_raw", c_void_p),
                ("vis_raw_ch", c_int), ("vid_raw_ch", c_int), ("text_raw_ch", c_int), ("text_ids", c_void_p),
                ("raw_layout", c_int), ("feat_layout", c_int)]


OUTPUT_FIELDS = ("pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m",
                 "att_sequences", "aux_boxes", "aux_sted", "aux_actioness", "choose1", "choose2", "actioness_pass1",
                 "boxes_px", "sted_idx", "encoded_feature", "frames_cls")


class VgqaOutputs(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in OUTPUT_FIELDS]


def _declare(L):
    if getattr(L, "_vgqa_engine_declared", False):
        return
    L.vgqa_create.restype = c_int
    L.vgqa_create.argtypes = [ctypes.POINTER(VgqaConfig), ctypes.POINTER(c_void_p)]
    L.vgqa_destroy.restype = None
    L.vgqa_destroy.argtypes = [c_void_p]
    L.vgqa_set_weight.restype = c_int
    L.vgqa_set_weight.argtypes = [c_void_p, ctypes.c_char_p, c_void_p, ctypes.POINTER(ctypes.c_int64), c_int]
    L.vgqa_finalize_weights.restype = c_int
    L.vgqa_finalize_weights.argtypes = [c_void_p]
    L.vgqa_forward.restype = c_int
    L.vgqa_forward.argtypes = [c_void_p, ctypes.POINTER(VgqaInputs), ctypes.POINTER(VgqaOutputs), c_void_p]
    L.vgqa_forward_async.restype = c_int
    L.vgqa_forward_async.argtypes = [c_void_p, ctypes.POINTER(VgqaInputs), ctypes.POINTER(VgqaOutputs), c_int, c_void_p]
    L.vgqa_forward_wait.restype = c_int
    L.vgqa_forward_wait.argtypes = [c_void_p, c_int, c_void_p, c_int]
    L.vgqa_forward_host.restype = c_int
    L.vgqa_forward_host.argtypes = [c_void_p, ctypes.POINTER(VgqaInputs), ctypes.POINTER(VgqaOutputs)]
    L.vgqa_forward_host_async.restype = c_int
    L.vgqa_forward_host_async.argtypes = [c_void_p, ctypes.POINTER(VgqaInputs), ctypes.POINTER(VgqaOutputs), c_int]
    L.vgqa_forward_host_wait.restype = c_int
    L.vgqa_forward_host_wait.argtypes = [c_void_p, c_int]
    L.vgqa_last_launch_count.restype = c_int
    L.vgqa_last_launch_count.argtypes = [c_void_p]
    L.vgqa_graph_capture_count.restype = c_int
    L.vgqa_graph_capture_count.argtypes = [c_void_p]
    L.vgqa_postprocess.restype = c_int
    L.vgqa_postprocess.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    L.vgqa_set_sharding.restype = c_int
    L.vgqa_set_sharding.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.vgqa_reference_flops.restype = ctypes.c_double
    L.vgqa_reference_flops.argtypes = [c_int] * 8
    L._vgqa_engine_declared = True


def reference_flops(T, H, W, L, enc_layers=6, dec_layers=6, ffn_dim=2048, passes=2) -> float:
    L_ = _lib.lib()
    _declare(L_)
    return float(L_.vgqa_reference_flops(T, H, W, L, enc_layers, dec_layers, ffn_dim, passes))


EXCHANGE_FN = ctypes.CFUNCTYPE(None, c_void_p, c_int, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p)


class GroundingEngine:
    """Owns the packed weights + workspace for up to (max_clips, max_frames, max_hw, max_text)."""

    def __init__(self, state_dict: Mapping[str, object], *, max_clips=1, max_frames=64, max_hw=49, max_text=32,
                 enc_layers=6, dec_layers=6, ffn_dim=2048, app_num=20, mot_num=34, max_video_len=200,
                 use_cuda_graph=False, device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("vgqa_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self._L = _lib.lib()
        _declare(self._L)
        if device is not None:
            torch.cuda.set_device(device)
        self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.init()
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        self.cfg = VgqaConfig(enc_layers, dec_layers, 256, 8, ffn_dim, app_num, mot_num, max_video_len, max_clips,
                              max_frames, max_hw, max_text, 1 if use_cuda_graph else 0)
        self._ctx = c_void_p()
        _lib.check(self._L.vgqa_create(ctypes.byref(self.cfg), ctypes.byref(self._ctx)))
        self.load_state_dict(state_dict)

    # -- weights ------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Mapping[str, object]):
        """Feeds every tensor of a reference state_dict (reference key names; extra keys are ignored by the
        library) and packs them."""
        for name, t in state_dict.items():
            if isinstance(t, torch.Tensor):
                a = t.detach().to("cpu", torch.float32).contiguous().numpy()
            else:
                a = np.ascontiguousarray(np.asarray(t), dtype=np.float32)
            if a.dtype != np.float32:
                continue
            shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
            _lib.check(self._L.vgqa_set_weight(self._ctx, name.encode(), a.ctypes.data_as(c_void_p), shape, a.ndim))
        _lib.check(self._L.vgqa_finalize_weights(self._ctx))

    def set_sharding(self, rank: int, world: int, exchange_cb=None):
        """Frame-shard one long clip over `world` ranks.  `exchange_cb` is a ctypes callback of type EXCHANGE_FN (see
        vgqa_b200/parallel.py: nccl_exchange); keep it alive as long as the engine."""
        self._exchange_cb = exchange_cb
        fn = ctypes.cast(exchange_cb, c_void_p) if exchange_cb is not None else None
        _lib.check(self._L.vgqa_set_sharding(self._ctx, rank, world, fn, None))

    def enable_p2p_sharding(self, rank: int, world: int, group=None):
        """Frame-shard one long clip over `world` ranks with the exchanges done on the device over NVLink peer memory
        (csrc/p2p_exchange.cu) — no NCCL call inside the forward, which is therefore captured into a CUDA graph.  The
        64-byte IPC handles of the ranks' exchange buffers are all-gathered once through torch.distributed."""
        import torch.distributed as dist
        L = self._L
        L.vgqa_shard_p2p_export.restype = c_int
        L.vgqa_shard_p2p_export.argtypes = [c_void_p, c_int, c_int, c_void_p]
        L.vgqa_shard_p2p_import.restype = c_int
        L.vgqa_shard_p2p_import.argtypes = [c_void_p, c_int, c_int, c_void_p]
        L.vgqa_shard_p2p_error.restype = c_int
        L.vgqa_shard_p2p_error.argtypes = [c_void_p]
        mine = (ctypes.c_ubyte * 64)()
        _lib.check(L.vgqa_shard_p2p_export(self._ctx, rank, world, mine))
        backend = dist.get_backend(group)
        dev = self.device if backend == "nccl" else torch.device("cpu")
        local = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, local, group=group)
        buf = (ctypes.c_ubyte * (world * 64))(*allh.cpu().tolist())
        _lib.check(L.vgqa_shard_p2p_import(self._ctx, rank, world, buf))
        dist.barrier(group=group)   # every rank has mapped every buffer before anyone pushes
        self._shard_cfg, self._shard_errors, self._p2p = (rank, world), [], True

    def p2p_error(self) -> int:
        return int(self._L.vgqa_shard_p2p_error(self._ctx)) if getattr(self, "_p2p", False) else 0

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.vgqa_destroy(self._ctx)
            self._ctx = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- forward ------------------------------------------------------------------------------------
    def output_shapes(self, B, T, H, W, L):
        D = self.cfg.dec_layers
        return {
            "pred_boxes": (B, T, 4), "pred_sted": (B, T, 2), "pred_actioness": (B, T), "logits_f_m": (B, T),
            "logits_f_a": (B, T), "logits_r_a": (B, self.cfg.app_num), "logits_r_m": (B, self.cfg.mot_num),
            "att_sequences": (B, T), "aux_boxes": (D, B, T, 4), "aux_sted": (D, B, T, 2), "aux_actioness": (D, B, T),
            "choose1": (B, T), "choose2": (B, T), "actioness_pass1": (B, T), "boxes_px": (B, T, 4), "sted_idx": (B, 2),
            "encoded_feature": (B * T, 2 * H * W + L, 256), "frames_cls": (B * T, 256),
        }

    def alloc_outputs(self, B, T, H, W, L, want=None, host=False) -> Dict[str, torch.Tensor]:
        want = set(want or [f for f in OUTPUT_FIELDS if f != "encoded_feature"])
        outs = {}
        for k, shp in self.output_shapes(B, T, H, W, L).items():
            if k not in want:
                continue
            dt = torch.int32 if k == "sted_idx" else torch.float32
            outs[k] = (torch.zeros(shp, dtype=dt).pin_memory() if host else torch.zeros(shp, dtype=dt, device=self.device))
        return outs

    @staticmethod
    def _dims(vis, raw):
        """(clips, T, H, W) of a feature tensor: [clips,T,C,H,W] fp32, or channels-last bf16 [clips,T,H,W,C] (raw only)."""
        if vis.dtype == torch.bfloat16:
            return vis.shape[0], vis.shape[1], vis.shape[2], vis.shape[3]
        return vis.shape[0], vis.shape[1], vis.shape[3], vis.shape[4]

    @staticmethod
    def _p(t):
        return None if t is None else c_void_p(t.data_ptr())

    def _pack_io(self, vis, vid, text, pos, vis_mask, text_mask, ori_sizes_hw, force1, force2, iteration_rate, outs,
                 stop_after_encoder=0, raw=False, text_ids=None):
        """raw=True: vis / vid / text are the extractor outputs ([clips,T,Cv,H,W], [clips,T,Cd,H,W], [clips,L,Ct]) and the
        library applies input_proj / input_proj2 / text_encoder.resizer itself (their weights must be in the state_dict)."""
        nhwc = bool(raw and vis.dtype == torch.bfloat16)   # channels-last bf16 maps [clips, T, H, W, C] (raw_layout = 1)
        rows = bool(not raw and vis.dtype == torch.bfloat16)   # projected maps as channels-last bf16 [clips, T, H, W, 256] (feat_layout = 1)
        if nhwc or rows:
            B, T, H, W, d = vis.shape
        else:
            B, T, d, H, W = vis.shape
        if text_ids is not None:   # RoBERTa token ids [clips, L] int32: the library runs the text tower + resizer (raw=True only)
            assert raw and text_ids.dtype == torch.int32 and text_ids.dim() == 2 and text_ids.shape[0] == B and text_ids.is_contiguous()
            text = torch.empty(B, text_ids.shape[1], 0, dtype=torch.float32, device=text_ids.device)   # placeholder (shape only)
        if rows:
            assert d == 256 and vid.dtype == torch.bfloat16 and tuple(vid.shape) == (B, T, H, W, 256), "vis / vid must both be bf16 [clips, T, H, W, 256]"
            assert tuple(text.shape) == (B, text.shape[1], 256), "text must be [clips, L, 256]"
        elif nhwc:
            assert vid.dtype == torch.bfloat16 and tuple(vid.shape[:4]) == (B, T, H, W), "vid_raw must be bf16 [clips, T, H, W, C] too"
        elif raw:
            assert tuple(vid.shape[:2]) == (B, T) and tuple(vid.shape[3:]) == (H, W), "vid_raw must be [clips, T, C, H, W]"
            assert text.dim() == 3 and text.shape[0] == B, "text_raw must be [clips, L, C]"
        else:
            assert d == 256 and tuple(vid.shape) == tuple(vis.shape), "vis/vid must be [clips, T, 256, H, W]"
            assert tuple(text.shape) == (B, text.shape[1], 256), "text must be [clips, L, 256]"
        Lt = text.shape[1]
        # pos = None: the library generates PositionEmbeddingSine itself (from vis_mask; one shared table without a mask)
        assert pos is None or (pos.shape[0] in (1, B * T) and tuple(pos.shape[1:]) == (256, H, W)), \
            "pos must be [1 or clips*T, 256, H, W] (or None)"
        for t in (vis, vid, text) + (() if pos is None else (pos,)):
            assert (t.dtype == torch.float32 or ((nhwc or rows) and (t is vis or t is vid))) and t.is_contiguous()
        n = None
        inp = VgqaInputs(B, T, H, W, Lt, n if raw else self._p(vis), n if raw else self._p(vid), n if raw else self._p(text),
                         self._p(pos), 0 if pos is None else pos.shape[0],
                         self._p(vis_mask), self._p(text_mask), self._p(ori_sizes_hw), self._p(force1), self._p(force2),
                         iteration_rate, stop_after_encoder,
                         self._p(vis) if raw else n, self._p(vid) if raw else n,
                         self._p(text) if (raw and text_ids is None) else n,
                         (vis.shape[4] if nhwc else vis.shape[2]) if raw else 0, (vid.shape[4] if nhwc else vid.shape[2]) if raw else 0,
                         text.shape[2] if raw else 0, self._p(text_ids), 1 if nhwc else 0, 1 if rows else 0)
        out = VgqaOutputs(**{k: self._p(outs.get(k)) for k in OUTPUT_FIELDS})
        return inp, out

    def encode(self, vis, vid, text, pos, *, vis_mask=None, text_mask=None, raw=False, text_ids=None):
        """CrossModalEncoder only: returns encoded_feature [clips*T, S, 256] (frame-major) and frames_cls [clips*T, 256]."""
        B, T, H, W = self._dims(vis, raw)
        outs = self.alloc_outputs(B, T, H, W, (text_ids if text_ids is not None else text).shape[1], ["encoded_feature", "frames_cls"])
        inp, out = self._pack_io(vis, vid, text, pos, vis_mask, text_mask, None, None, None, -1, outs, 1, raw=raw, text_ids=text_ids)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_forward(self._ctx, ctypes.byref(inp), ctypes.byref(out), c_void_p(st)))
        return outs

    def forward(self, vis, vid, text, pos=None, *, vis_mask=None, text_mask=None, ori_sizes_hw=None, force_choose1=None,
                force_choose2=None, iteration_rate=-1, outs=None, want=None, raw=False, text_ids=None):
        """Device-resident inputs (fp32 CUDA tensors, reference layouts); enqueues on the current stream."""
        B, T, H, W = self._dims(vis, raw)
        if outs is None:
            outs = self.alloc_outputs(B, T, H, W, (text_ids if text_ids is not None else text).shape[1], want)
            if ori_sizes_hw is None:
                outs.pop("boxes_px", None), outs.pop("sted_idx", None)
        inp, out = self._pack_io(vis, vid, text, pos, vis_mask, text_mask, ori_sizes_hw, force_choose1, force_choose2,
                                 iteration_rate, outs, raw=raw, text_ids=text_ids)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_forward(self._ctx, ctypes.byref(inp), ctypes.byref(out), c_void_p(st)))
        return outs

    def forward_async(self, vis, vid, text, pos, *, outs, slot, vis_mask=None, text_mask=None, ori_sizes_hw=None,
                      force_choose1=None, force_choose2=None, iteration_rate=-1, raw=False, text_ids=None):
        """Pipelined device path: alternate slot 0/1 on consecutive calls; `outs` are complete after `wait(slot)`."""
        inp, out = self._pack_io(vis, vid, text, pos, vis_mask, text_mask, ori_sizes_hw, force_choose1, force_choose2,
                                 iteration_rate, outs, raw=raw, text_ids=text_ids)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_forward_async(self._ctx, ctypes.byref(inp), ctypes.byref(out), slot, c_void_p(st)))
        return outs

    def wait(self, slot: int, host_sync: bool = False):
        """Make the current stream (or the host) wait for the call last issued on `slot`."""
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_forward_wait(self._ctx, slot, c_void_p(st), 1 if host_sync else 0))

    def forward_host(self, vis, vid, text, pos, *, vis_mask=None, text_mask=None, ori_sizes_hw=None,
                     force_choose1=None, force_choose2=None, iteration_rate=-1, outs=None, want=None, raw=False, text_ids=None):
        """Host buffers (CPU tensors, ideally pinned): H2D + forward + D2H inside the call (synchronous)."""
        B, T, H, W = self._dims(vis, raw)
        if outs is None:
            outs = self.alloc_outputs(B, T, H, W, (text_ids if text_ids is not None else text).shape[1], want, host=True)
            if ori_sizes_hw is None:
                outs.pop("boxes_px", None), outs.pop("sted_idx", None)
        inp, out = self._pack_io(vis, vid, text, pos, vis_mask, text_mask, ori_sizes_hw, force_choose1, force_choose2,
                                 iteration_rate, outs, raw=raw, text_ids=text_ids)
        _lib.check(self._L.vgqa_forward_host(self._ctx, ctypes.byref(inp), ctypes.byref(out)))
        return outs

    def forward_host_async(self, vis, vid, text, pos, *, outs, slot, vis_mask=None, text_mask=None, ori_sizes_hw=None,
                           force_choose1=None, force_choose2=None, iteration_rate=-1, raw=False, text_ids=None):
        """Pipelined host path: returns immediately; `outs` (pinned host tensors) are valid after `wait_host(slot)`."""
        inp, out = self._pack_io(vis, vid, text, pos, vis_mask, text_mask, ori_sizes_hw, force_choose1, force_choose2,
                                 iteration_rate, outs, raw=raw, text_ids=text_ids)
        _lib.check(self._L.vgqa_forward_host_async(self._ctx, ctypes.byref(inp), ctypes.byref(out), slot))
        return outs

    def wait_host(self, slot: int):
        _lib.check(self._L.vgqa_forward_host_wait(self._ctx, slot))

    def text_tower(self, text_ids, text_mask=None):
        """RoBERTa tower + resizer alone (parity tests): ids [clips, L] int32 (device) → (last_hidden_state [clips, L, Hd] fp32 of
        the bf16 rows the resizer reads, resized text [clips, L, 256] fp32)."""
        B, L = text_ids.shape
        self._L.vgqa_text_tower.restype = c_int
        self._L.vgqa_text_tower.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
        self._L.vgqa_text_tower_hidden.restype = c_int
        self._L.vgqa_text_tower_hidden.argtypes = [c_void_p]
        Hd = int(self._L.vgqa_text_tower_hidden(self._ctx))
        hidden = torch.zeros(B, L, Hd, device=self.device)
        text = torch.zeros(B, L, 256, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_text_tower(self._ctx, self._p(text_ids), self._p(text_mask), B, L, self._p(hidden), self._p(text),
                                           c_void_p(st)))
        return hidden, text

    def swin_stage(self, x, want_f32=False):
        """Last Video-Swin-T stage (`vid.layers[3]`, csrc/swin.cu): x fp32 channels-last [clips, T, H, W, 768] (device) →
        channels-last bf16 map of the same shape (the `vid` raw input of forward(raw=True)), and the fp32 map when asked."""
        assert x.dtype == torch.float32 and x.dim() == 5 and x.is_contiguous()
        B, T, H, W, C = x.shape
        self._L.vgqa_swin_stage.restype = c_int
        self._L.vgqa_swin_stage.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
        out = torch.empty(B, T, H, W, C, dtype=torch.bfloat16, device=self.device)
        out32 = torch.empty(B, T, H, W, C, dtype=torch.float32, device=self.device) if want_f32 else None
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_swin_stage(self._ctx, self._p(x), B, T, H, W, self._p(out), self._p(out32), c_void_p(st)))
        return (out, out32) if want_f32 else out

    def swin_backbone(self, frames, clips, want_stages=False):
        """Whole Video-Swin-T extractor (csrc/swin.cu): frames fp32 NCHW [clips*T, 3, R, R] (device) → the last stage's map as
        channels-last bf16 [clips, T, R/32, R/32, 768]; with want_stages also the four stage outputs (channels-last fp32).
        T >= 8, R a multiple of 32 from 224 on (sides that are not multiples of the (8,7,7) window are padded as the reference does)."""
        assert frames.dtype == torch.float32 and frames.dim() == 4 and frames.shape[1] == 3 and frames.is_contiguous()
        n, _, R, _ = frames.shape
        T = n // clips
        self._L.vgqa_swin_backbone.restype = c_int
        self._L.vgqa_swin_backbone.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        out = torch.empty(clips, T, R // 32, R // 32, 768, dtype=torch.bfloat16, device=self.device)
        stages, ptrs = None, None
        if want_stages:
            stages = [torch.empty(clips, T, (R // 4) >> s, (R // 4) >> s, 96 << s, device=self.device) for s in range(4)]
            ptrs = (c_void_p * 4)(*[c_void_p(t.data_ptr()) for t in stages])
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_swin_backbone(self._ctx, self._p(frames), clips, T, R, self._p(out), None, ptrs, c_void_p(st)))
        return (out, stages) if want_stages else out

    def resnet_backbone(self, frames, want_layers=False):
        """ResNet101 extractor (csrc/resnet.cu): frames fp32 NCHW [n, 3, R, R] (device) → the layer4 map as channels-last bf16
        [n, R/32, R/32, 2048]; with want_layers also the outputs of layer1..4 (channels-last fp32)."""
        assert frames.dtype == torch.float32 and frames.dim() == 4 and frames.shape[1] == 3 and frames.is_contiguous()
        n, _, R, _ = frames.shape
        self._L.vgqa_resnet_backbone.restype = c_int
        self._L.vgqa_resnet_backbone.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        out = torch.empty(n, R // 32, R // 32, 2048, dtype=torch.bfloat16, device=self.device)
        layers, ptrs = None, None
        if want_layers:
            layers = [torch.empty(n, (R // 4) >> l, (R // 4) >> l, 256 << l, device=self.device) for l in range(4)]
            ptrs = (c_void_p * 4)(*[c_void_p(t.data_ptr()) for t in layers])
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.vgqa_resnet_backbone(self._ctx, self._p(frames), n, R, self._p(out), None, ptrs, c_void_p(st)))
        return (out, layers) if want_layers else out

    def extract_features(self, frames, clips):
        """Both extractors on `frames` fp32 NCHW [clips*T, 3, R, R] → (ResNet101 layer4 map [clips, T, R/32, R/32, 2048],
        Video-Swin-T map [clips, T, R/32, R/32, 768]), channels-last bf16 = the raw_layout 1 inputs of forward().  Short clips do
        not fill the GPU with either network (layer3 of 32 frames is 64 tiles for 148 SMs), so up to 64 frames the two run
        concurrently on two streams."""
        n = frames.shape[0]
        assert clips >= 1 and n % clips == 0, "frames must hold clips * T pictures"
        if n > 64:
            vis = self.resnet_backbone(frames)
            vid = self.swin_backbone(frames, clips)
        else:
            cur = torch.cuda.current_stream()
            if getattr(self, "_aux_stream", None) is None:
                self._aux_stream = torch.cuda.Stream()
            self._aux_stream.wait_stream(cur)
            with torch.cuda.stream(self._aux_stream):
                vis = self.resnet_backbone(frames)
            vid = self.swin_backbone(frames, clips)
            cur.wait_stream(self._aux_stream)
            vis.record_stream(cur)
        return vis.view(clips, n // clips, vis.shape[1], vis.shape[2], 2048), vid

    @property
    def last_launch_count(self) -> int:
        return int(self._L.vgqa_last_launch_count(self._ctx))

    @property
    def graph_capture_count(self) -> int:
        """CUDA graphs captured so far: one per (phase, slot, shape) — fresh input / output tensors do not re-capture."""
        return int(self._L.vgqa_graph_capture_count(self._ctx))
