#!/usr/bin/env python
"""bench.py — grounding clips/sec on the VGQA hot path (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips B] [--impl reference]

A step = one pass of the hot path (encoder → classifiers → 2 decoder passes → heads → PostProcess) over one batch
of B synthetic clips of BASELINE.json configs[1]: 64 frames @224 (7x7 feature map), 20-token query, 6+6 layers,
random-init (seeded synthetic) weights.  Prints ONE JSON line (rank 0).

  value        clips/s with the inputs already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e          clips/s through the C-ABI host path (`vgqa_forward_host_async/_wait`, two slots) with pinned HOST buffers:
               every step uploads its inputs, computes, downloads and reads its results inside the timed region
  roofline     the dominant kernel (fused FFN block, tcgen05 cta_group::2) timed alone with CUDA events: algorithmic FLOPs / duration
  cpu_baseline the reference's OWN PyTorch modules (oracle/_ref, byte-compiled from /root/reference by tools/make_oracle_ref.py)
               timed on the host cores on a bounded sample (rank 0, N=1 only); the numpy port only if oracle/_ref is absent
  other_configs  cfg-5 (T=128, 12x12, L=64) at 1 / 8 / 64 clips per step and the real yaml shape (T=64, 14x14, L=20), timed in-run
  sharded_cfg4   (WORLD_SIZE > 1) ONE 256-frame clip frame-sharded over all ranks, peer-memory exchange: ms per clip + error vs golden
  eager_pytorch_b200  the reference modules in eager PyTorch bf16-autocast on the same B200 (the only existing implementation)
  --impl reference : the reference arm = the reference's own modules on the host cores, all threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))    # ref_loader / make_golden: the reference-module harness (CPU arm only)

T, H, W, L = 64, 7, 7, 20
METRIC = "grounding clips/sec (64f@224, bf16)"   # BASELINE.json's metric; both arms print the same string
WORKLOAD = "cfg2 grounding_vidstg.yaml@224: T=64 frames, 7x7 feature map, L=20 text tokens, 6 enc + 6 dec layers, 2 decoder passes"


TEXT_TOWER = (12, 50265)          # RoBERTa-base: encoder layers, vocabulary (bert.py:49)
FRONT_END_CH = (2048, 768, 768)   # ResNet101 layer-4, Video-Swin stage-3, RoBERTa-base channels (grounding_net.py:62,71; bert.py:77)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def reference_arm_available():
    try:
        from ref_loader import reference_modules_available
        return reference_modules_available()
    except Exception:
        return False


class ReferenceHotPath:
    """The reference's own PyTorch modules (imported from /root/reference here, from the byte-compiled oracle/_ref on the GPU
    box), wired and driven exactly as VSTGNet.forward lines 114-181 drive them (tests/golden/make_golden.py: RefHotPath), on the
    synthetic weights of `seed`.  fp32 on the host cores, or eager bf16-autocast on a CUDA device."""

    def __init__(self, seed=0, device="cpu", shape=(T, H, W, L)):
        import torch
        from vgqa_b200 import synth
        from make_golden import RefHotPath, load_synth, make_cfg
        from ref_loader import load_reference
        self.torch, self.device, self.shape = torch, device, shape
        self.R = load_reference()
        self.model = RefHotPath(self.R, make_cfg(max_video_len=max(200, shape[0]))).eval()
        load_synth(self.model, synth.synth_state_dict(seed, max_video_len=max(200, shape[0])))
        self.model.to(device)
        self.synth = synth

    def inputs(self, i):
        torch = self.torch
        Tn, Hn, Wn, Ln = self.shape
        vis, vid, pos, text = self.synth.synth_inputs(i, Tn, Hn, Wn, Ln)
        to = lambda a: torch.from_numpy(a).to(self.device)
        return (to(vis), to(vid), to(pos), to(text), torch.zeros(Tn, Hn, Wn, dtype=torch.bool, device=self.device),
                torch.zeros(1, Ln, dtype=torch.bool, device=self.device))

    def run(self, inp, autocast=False):
        torch = self.torch
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out, _ = self.model(*inp)
        return out


def cpu_clips_per_sec(n_clips, warmup=1):
    """Seconds per clip of the same workload on the host cores: the reference's own modules when they are importable (kind
    "reference"), else the numpy oracle port (kind "port").  All host threads."""
    cores = os.cpu_count()
    times = []
    if reference_arm_available():
        import torch
        torch.set_num_threads(cores)     # torch.distributed.run exports OMP_NUM_THREADS=1: undo it for the CPU arm
        ref = ReferenceHotPath(0, "cpu")
        for i in range(warmup + n_clips):
            inp = ref.inputs(i)
            t0 = time.perf_counter()
            ref.run(inp)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "reference"
        from ref_loader import REF_ROOT
        what = (f"the reference's own PyTorch modules (imported from {os.path.relpath(REF_ROOT, ROOT) if REF_ROOT.startswith(ROOT) else REF_ROOT}), "
                f"fp32, torch {torch.__version__}, {torch.get_num_threads()} threads")
    else:
        from oracle import vgqa_oracle as O
        from vgqa_b200 import synth
        sd = synth.synth_state_dict(0)
        for i in range(warmup + n_clips):
            vis, vid, pos, text = synth.synth_inputs(i, T, H, W, L)
            t0 = time.perf_counter()
            O.hot_path_forward(sd, vis, vid, pos, text)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind, what = "port", "numpy oracle port of the reference modules (oracle/_ref absent), fp32, all BLAS threads"
    return n_clips / sum(times), times, kind, what, cores


def run_reference_arm(args, rank):
    if rank != 0:
        return
    v, t_steps, kind, what, cores = cpu_clips_per_sec(args.steps, warmup=max(1, args.warmup))
    # BASELINE.json configs[0]: the reference's FULL VSTGNet.forward (backbones + RoBERTa + hot path) on the host cores, mini yaml
    # (224 px, 32 frames), synthetic clip, random-init weights — one warm-up + one timed call (≈15 s on 16 cores)
    full = None
    if kind == "reference" and not args.no_full_forward:
        try:   # own process: the hot-path harness above imported the reference packages with their __init__ files bypassed
            env = dict(os.environ)
            env.pop("OMP_NUM_THREADS", None)          # torch.distributed.run exports OMP_NUM_THREADS=1
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "full_forward_cpu.py")], capture_output=True, text=True,
                               timeout=900, cwd=ROOT, env=env)
            full = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-500:]}
        except Exception as ex:   # noqa: BLE001 — reported in the line
            full = {"error": f"{type(ex).__name__}: {ex}"}
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_steps) / len(t_steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step": 1},
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": kind,
                             "sample": f"{len(t_steps)} clips of the same workload, one per step ({what})"},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "full_forward_cfg1": full}
    print(json.dumps(line), flush=True)


def static_metrics():
    """ncu-derived figures that cannot be measured inside a timed run (DRAM bytes of one launch, tensor-pipe activity): read from
    the tracked profiles/static_metrics.json, which names the commit and the ncu command they were taken with.  Absent → null."""
    p = os.path.join(ROOT, "profiles", "static_metrics.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def time_dominant_kernel(B, pk):
    """The fused FFN block of one encoder layer over the whole batch — ffn_fused_kernel (CTA pairs, tcgen05 cta_group::2):
    y = LN(x32 + W2 relu(W1 x + b1) + b2), M = B*T*S rows, 256 -> 2048 -> 256.  Algorithmic FLOPs = 2 * (2*M*256*2048)."""
    import torch
    from vgqa_b200 import _lib
    Lb = _lib.lib()
    S = 2 * H * W + L
    M, F = B * T * S, 2048
    x32 = torch.randn(M, 256, device="cuda")
    x = x32.bfloat16()
    W1 = (torch.randn(F, 256, device="cuda") / 16).bfloat16()
    W2 = (torch.randn(256, F, device="cuda") / F ** 0.5).bfloat16()
    b1 = torch.zeros(F, device="cuda")
    b2 = torch.zeros(256, device="cuda")
    lw = torch.ones(256, device="cuda")
    pos = torch.randn(S, 256, device="cuda").bfloat16()
    C = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
    C2 = torch.empty_like(C)
    C32 = torch.empty(M, 256, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        _lib.check(Lb.vgqa_ffn_fused(_lib.ptr(x), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), M, F, _lib.ptr(x32),
                                     _lib.ptr(lw), _lib.ptr(b2), 1e-5, _lib.ptr(C), _lib.ptr(C32), _lib.ptr(C2), _lib.ptr(pos), S, 0, st))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / reps * 1e-3
    flops = 4.0 * M * F * 256
    alg_bytes = M * (512 + 1024 + 512 + 1024 + 512) + 2 * F * 256 * 2     # x, residual in; bf16, fp32, bf16(x+pos) out; weights once
    ach = flops / dt / 1e12
    sm = static_metrics().get("ffn_fused_dram_bytes", {})
    traffic = sm["bytes_per_launch_at_64_clips"] * (B / 64.0) if "bytes_per_launch_at_64_clips" in sm else None
    return {"kernel": "ffn_fused_kernel (encoder FFN block linear1+ReLU+linear2+residual+LayerNorm, M=%d, 256->2048->256)" % M,
            "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
            "traffic": traffic, "traffic_source": sm.get("source") if traffic is not None else None, "peak_source": pk["source"] + ", burst (kernel timed alone)",
            "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dt * 1e3,
            "frac_of_sustained_peak": ach / pk["bf16_tflops_sustained"],
            "hbm_gbs_at_algorithmic_bytes": alg_bytes / dt / 1e9}


def measure_with_backbone(torch, rank=0, clips=8, steps=5):
    """SURVEY §8d secondary number: clips/s with the extractors in front of the library, 8 clips x 64 frames of 3x224x224 per step.
      * ResNet101: this library's kernels (csrc/resnet.cu) with seeded weights; the same network in torchvision (FrozenBatchNorm2d,
        bf16 autocast, channels_last — cuDNN) is timed beside it and cross-checked against it; the layer-4 map goes into the forward
        zero-copy (channels-last bf16 = raw_layout 1);
      * Video-Swin-T: the whole extractor on this library's kernels (csrc/swin.cu) with the weights of the reference's own module
        (oracle/_ref, random init), which is timed beside it in eager PyTorch bf16.  Without oracle/_ref the Video-Swin map is synthetic;
      * RoBERTa states synthetic (the text tower has its own `front_end.from_token_ids` line)."""
    try:
        import torchvision
    except Exception:
        return None
    from vgqa_b200 import synth
    from vgqa_b200.engine import GroundingEngine
    from torchvision.ops.misc import FrozenBatchNorm2d            # the arithmetic of the reference's class (backbone.py:47-57, eps 1e-5)
    rsd = synth.synth_resnet101(0)
    net = torchvision.models.resnet101(weights=None, norm_layer=FrozenBatchNorm2d)
    net.load_state_dict({k[len("vis_encoder.0.body."):]: torch.from_numpy(v) for k, v in rsd.items()}, strict=False)
    body = torch.nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool, net.layer1, net.layer2, net.layer3, net.layer4)
    body = body.eval().cuda().to(memory_format=torch.channels_last)
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    frames = torch.randn(clips * T, 3, 224, 224, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    text = torch.randn(clips, L, FRONT_END_CH[2], device="cuda", generator=g)
    sizes = torch.tensor([[360.0, 640.0]] * clips, device="cuda")
    sd = synth.synth_state_dict(0, front_end_ch=FRONT_END_CH)
    sd.update(rsd)
    swin = None
    swin_note = "Video-Swin map synthetic (oracle/_ref absent)"
    if reference_arm_available():
        try:
            from make_golden_swin import load_swin_module
            M = load_swin_module()
            torch.manual_seed(5)
            swin = M.vidswin_model("video_swin_t_p4w7", None).eval().cuda()
            for k, v in swin.state_dict().items():
                if "relative_position_index" not in k:
                    sd["vid." + k] = v.detach().float().cpu().numpy()
            swin_note = ("Video-Swin-T: the WHOLE extractor on this library's kernels (csrc/swin.cu) with the weights of the reference "
                         "module (which is timed beside it in eager PyTorch bf16)")
        except Exception as ex:   # noqa: BLE001
            swin, swin_note = None, f"Video-Swin map synthetic ({type(ex).__name__}: {ex})"
    eng = GroundingEngine(sd, max_clips=clips, max_frames=T, max_hw=H * W, max_text=L, use_cuda_graph=True)
    vid_syn = torch.randn(clips, T, H, W, FRONT_END_CH[1], device="cuda", generator=g).to(torch.bfloat16)
    want = ["pred_boxes", "pred_sted", "boxes_px", "sted_idx"]
    outs = eng.alloc_outputs(clips, T, H, W, L, want)
    from einops import rearrange

    def swin_front():
        """patch embedding + stages 1-3 + PatchMerging of the reference module → the channels-last input of the last stage"""
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            x = rearrange(frames, "(b t) c h w -> b c t h w", b=clips, t=T)
            x = swin.pos_drop(swin.patch_embed(x))
            for idx in range(3):
                x = swin.layers[idx](x.contiguous())
                x = rearrange(swin.downsamples[idx](rearrange(x, "b c t h w -> b t h w c")), "b t h w c -> b c t h w")
        return rearrange(x, "b c t h w -> b t h w c").float().contiguous()          # [clips, T, 7, 7, 768]

    frames32 = frames.float().contiguous()       # `videos.tensors` as the reference feeds them (NCHW fp32)

    def resnet_pytorch():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return body(frames)

    def swin_pytorch():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return swin(frames, T)["3"]

    def step(mode="all"):
        fmap = eng.resnet_backbone(frames32)                               # [clips*T, 7, 7, 2048] bf16, channels-last
        if mode == "resnet":
            return
        vid = eng.swin_backbone(frames32, clips) if swin is not None else vid_syn
        if mode == "extractors":
            return
        vis = fmap.view(clips, T, H, W, FRONT_END_CH[0])                   # handed over zero-copy (raw_layout = 1)
        eng.forward(vis, vid, text, None, ori_sizes_hw=sizes, outs=outs, raw=True)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / steps

    res = {}
    if swin is not None:   # the repo's last stage against the PyTorch module's on the same activations
        x4 = swin_front()
        mine = eng.swin_stage(x4, want_f32=True)[1]
        with torch.no_grad():
            ref4 = swin.layers[3](rearrange(x4, "b t h w c -> b c t h w").contiguous())
        ref4 = rearrange(ref4, "b c t h w -> b t h w c")
        res["swin_stage4_max_abs_err_vs_pytorch_fp32"] = float((mine - ref4).abs().max())
        res["swin_stage4_mean_abs"] = float(ref4.abs().mean())
        sec_sw4 = timed(lambda: eng.swin_stage(x4))
        res["swin_stage4_repo_ms_per_step"] = 1e3 * sec_sw4
        # the whole extractor: this library vs the reference module in eager PyTorch (bf16 autocast) on the same frames / weights
        mine_full = eng.swin_backbone(frames32, clips).float()
        with torch.no_grad():
            ref_full = swin(frames32, T)["3"]                                          # fp32, (clips*T, 768, 7, 7)
        ref_full = ref_full.reshape(clips, T, 768, 7, 7).permute(0, 1, 3, 4, 2)
        res["swin_backbone_max_abs_err_vs_pytorch_fp32"] = float((mine_full - ref_full).abs().max())
        res["swin_backbone_mean_abs"] = float(ref_full.abs().mean())
        res["swin_backbone_repo_ms_per_step"] = 1e3 * timed(lambda: eng.swin_backbone(frames32, clips))
        res["swin_backbone_launches"] = eng.last_launch_count
        res["swin_backbone_pytorch_bf16_ms_per_step"] = 1e3 * timed(swin_pytorch)
    # the library's ResNet101 against the same network in PyTorch (fp32) on the same frames / weights
    with torch.no_grad():
        ref_r = body(frames32.contiguous(memory_format=torch.channels_last)).permute(0, 2, 3, 1)
    mine_r = eng.resnet_backbone(frames32).float()
    res["resnet_mean_abs_err_vs_pytorch_fp32"] = float((mine_r - ref_r).abs().mean())
    res["resnet_mean_abs"] = float(ref_r.abs().mean())
    res["resnet_pytorch_autocast_mean_abs_err"] = float((resnet_pytorch().permute(0, 2, 3, 1).float() - ref_r).abs().mean())
    res["resnet_launches"] = eng.last_launch_count
    res["resnet_pytorch_bf16_ms_per_step"] = 1e3 * timed(resnet_pytorch)
    del ref_r, mine_r
    sec, sec_bb = timed(step), timed(lambda: step("resnet"))
    sec_ex = timed(lambda: step("extractors"))
    eng.close()
    del body, frames, swin, frames32
    torch.cuda.empty_cache()
    res.update({"value": clips / sec, "unit": "clips/s", "ms_per_step": 1e3 * sec, "clips_per_step": clips,
                "backbone_only_ms_per_step": 1e3 * sec_bb, "extractors_ms_per_step": 1e3 * sec_ex,
                "what": "ResNet101 on this library's kernels (csrc/resnet.cu; `backbone_only_ms_per_step`; the same network in torchvision / cuDNN "
                        "bf16 channels_last = `resnet_pytorch_bf16_ms_per_step`) on 64 x 3x224x224 frames per clip → layer-4 map handed over "
                        "zero-copy (channels-last bf16, raw_layout = 1); " + swin_note +
                        " → this library's raw-input forward; RoBERTa states synthetic"})
    return res


def measure_full_forward(torch, frames=32, reps=10):
    """BASELINE.json configs[0] (the reference's CPU-runnable case: ONE full `VSTGNet.forward`, 32 frames @224, batch 1) on the
    library alone: ResNet101 + Video-Swin-T + RoBERTa-base tower + front end + hot path from NCHW fp32 pixels and token ids.  Its
    CPU counterpart is `full_forward_cfg1` of `--impl reference` (the unmodified reference model on the host cores)."""
    from vgqa_b200 import synth
    from vgqa_b200.engine import GroundingEngine
    sd = synth.synth_state_dict(0, front_end_ch=FRONT_END_CH, text_tower=(12, 50265))
    sd.update(synth.synth_resnet101(0))
    sd.update(synth.synth_swin_backbone(0))
    ids = torch.tensor([[0, 5910, 36777, 43933, 5772, 33081, 34763, 2]], dtype=torch.int32, device="cuda")   # "a person jumping over the fence"
    eng = GroundingEngine(sd, max_clips=1, max_frames=frames, max_hw=H * W, max_text=ids.shape[1], use_cuda_graph=True)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(frames, 3, 224, 224, device="cuda", generator=g)
    sizes = torch.tensor([[360.0, 640.0]], device="cuda")
    outs = eng.alloc_outputs(1, frames, H, W, ids.shape[1], ["pred_boxes", "pred_sted", "boxes_px", "sted_idx"])
    host = torch.empty(frames, 4).pin_memory()

    def step():
        vis, vid = eng.extract_features(x, 1)
        eng.forward(vis, vid, None, None, ori_sizes_hw=sizes, outs=outs, raw=True, text_ids=ids)
        host.copy_(outs["boxes_px"][0], non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    eng.close()
    torch.cuda.empty_cache()
    return {"ms_per_forward": ms, "forwards_per_s": 1e3 / ms, "frames": frames, "resolution": 224,
            "what": "one full forward of the model from pixels (NCHW fp32, device) + token ids to PostProcess boxes read back to the host, "
                    "every layer on this library's kernels (ResNet101, Video-Swin-T, RoBERTa-base, input_proj*, encoder, decoders, heads); "
                    "seeded random weights; the reference's own number is `full_forward_cfg1.seconds_per_forward` of `--impl reference`"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=64, help="clips per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-full-forward", action="store_true", help="reference arm: skip the full VSTGNet.forward (cfg-1) measurement")
    ap.add_argument("--quick", action="store_true", help="headline + e2e + roofline only (skip the other configs / eager-PyTorch lines)")
    ap.add_argument("--cpu-clips", type=int, default=60, help="clips of the bounded CPU-baseline sample (≈10-15 s of host work on 16 cores)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from vgqa_b200 import synth as O         # seeded synthetic weights / inputs (pure numpy; nothing from oracle/ on the GPU arm)
    from vgqa_b200.engine import GroundingEngine, reference_flops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: vgqa_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.clips
    pk = peaks()
    # hot-path weights of seed 0 + the front-end weights (input_proj / input_proj2 / text resizer) for the secondary
    # `front_end` measurement; the hot-path tensors of a seed do not depend on front_end_ch
    eng = GroundingEngine(O.synth_state_dict(0, front_end_ch=FRONT_END_CH, text_tower=TEXT_TOWER), max_clips=B, max_frames=T, max_hw=H * W, max_text=L,
                          use_cuda_graph=not args.no_graph)
    # synthetic inputs: 8 distinct seeded clips tiled to B (per-rank offset), fp32 reference layouts
    base = [O.synth_inputs(rank * 8 + i, T, H, W, L) for i in range(8)]
    pos = torch.from_numpy(base[0][2][:1].copy())
    h_vis = torch.from_numpy(np.stack([base[i % 8][0] for i in range(B)])).pin_memory()
    h_vid = torch.from_numpy(np.stack([base[i % 8][1] for i in range(B)])).pin_memory()
    h_text = torch.from_numpy(np.stack([base[i % 8][3][:, 0, :] for i in range(B)])).pin_memory()
    h_sizes = torch.tensor([[360.0, 640.0]] * B).pin_memory()
    h_pos = pos.pin_memory()
    d_vis, d_vid, d_text, d_pos, d_sizes = (x.cuda() for x in (h_vis, h_vid, h_text, h_pos, h_sizes))
    want = ["pred_boxes", "pred_sted", "pred_actioness", "att_sequences", "boxes_px", "sted_idx", "logits_r_a", "logits_r_m"]
    d_outs = eng.alloc_outputs(B, T, H, W, L, want)
    h_outs = eng.alloc_outputs(B, T, H, W, L, want, host=True)

    d_outs2 = [d_outs, eng.alloc_outputs(B, T, H, W, L, want)]
    dev_i = [0]

    def step_dev():
        # pipelined public API (two boundary slots): the decoder phase of step i overlaps the encoder phase of step i+1
        slot = dev_i[0] & 1
        eng.forward_async(d_vis, d_vid, d_text, d_pos, ori_sizes_hw=d_sizes, outs=d_outs2[slot], slot=slot)
        dev_i[0] += 1

    def drain_dev():
        eng.wait(0), eng.wait(1)     # the current stream waits for both slots → the closing CUDA event covers all steps

    h_outs2 = [h_outs, eng.alloc_outputs(B, T, H, W, L, want, host=True)]
    host_i = [0]
    # the host boundary's B200-native feature format (vgqa_inputs.feat_layout = 1): channels-last bf16 [clips, T, H, W, 256] —
    # half the PCIe bytes of the reference's fp32 NCHW maps (the path rounds its GEMM operands to bf16 anyway)
    h_vis_cl = h_vis.permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16).pin_memory()
    h_vid_cl = h_vid.permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16).pin_memory()
    host_fmt = [h_vis_cl, h_vid_cl]

    def step_host():
        # pipelined public API: the upload of this step overlaps the compute of the previous one; the results of step
        # i-2 (same slot) are complete — and read — before the slot is reused.
        slot = host_i[0] & 1
        eng.wait_host(slot)
        _ = float(h_outs2[slot]["pred_boxes"][0, 0, 0])   # host read of the step's result
        eng.forward_host_async(host_fmt[0], host_fmt[1], h_text, h_pos, ori_sizes_hw=h_sizes, outs=h_outs2[slot], slot=slot)
        host_i[0] += 1

    def drain_host():
        eng.wait_host(0), eng.wait_host(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, device_events=True, drain=None):
        for _ in range(args.warmup):
            fn()
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sec = e0.elapsed_time(e1) * 1e-3 if device_events else wall
        barrier()
        if world > 1:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sec = timed(step_dev, args.steps, drain=drain_dev)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.last_launch_count
    # e2e: pinned host buffers through the C-ABI (vgqa_forward_host_async/_wait, two slots); timed by wall clock
    # around enqueue + final drain (every step's H2D, compute and D2H are inside)
    sec_e2e = timed(step_host, args.steps, device_events=False, drain=drain_host)
    host_fmt[0], host_fmt[1] = h_vis, h_vid          # the same through the reference's fp32 NCHW layout (twice the upload)
    sec_e2e_f32 = timed(step_host, args.steps, device_events=False, drain=drain_host)
    total_clips = B * world * args.steps

    def secondary():
        """Secondary lines (front end, batch 1): same collectives on every rank; a failure here must not lose the headline."""
        # secondary: the same step fed with the RAW extractor maps resident in HBM (ResNet101 2048-ch + Video-Swin 768-ch fp32
        # NCHW, RoBERTa 768-d states): input_proj / input_proj2 / resizer run inside the forward (csrc/input_proj.cu)
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        r_vis = torch.randn(B, T, FRONT_END_CH[0], H, W, device="cuda", generator=g).relu_()
        r_vid = torch.randn(B, T, FRONT_END_CH[1], H, W, device="cuda", generator=g)
        r_text = torch.randn(B, L, FRONT_END_CH[2], device="cuda", generator=g)
        raw_i = [0]

        def step_raw():
            slot = raw_i[0] & 1
            eng.forward_async(r_vis, r_vid, r_text, d_pos, ori_sizes_hw=d_sizes, outs=d_outs2[slot], slot=slot, raw=True)
            raw_i[0] += 1

        sec_raw = timed(step_raw, args.steps, drain=drain_dev)
        # ... and from RoBERTa token ids: the 12-layer text tower (csrc/text_tower.cu + tcgen05 GEMMs) runs inside the forward too
        r_ids = torch.from_numpy(O.synth_text_ids(rank, B, L, TEXT_TOWER[1])[0]).cuda()

        def step_ids():
            slot = raw_i[0] & 1
            eng.forward_async(r_vis, r_vid, None, d_pos, ori_sizes_hw=d_sizes, outs=d_outs2[slot], slot=slot, raw=True, text_ids=r_ids)
            raw_i[0] += 1

        sec_ids = timed(step_ids, args.steps, drain=drain_dev)
        # ... and with the maps in the B200-native layout (raw_layout = 1): channels-last bf16, read by TMA as the GEMM operand itself
        c_vis = r_vis.permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)
        c_vid = r_vid.permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)

        def step_cl():
            slot = raw_i[0] & 1
            eng.forward_async(c_vis, c_vid, r_text, d_pos, ori_sizes_hw=d_sizes, outs=d_outs2[slot], slot=slot, raw=True)
            raw_i[0] += 1

        sec_cl = timed(step_cl, args.steps, drain=drain_dev)
        del c_vis, c_vid
        raw_bytes = int((r_vis.numel() + r_vid.numel() + r_text.numel()) * 4)
        del r_vis, r_vid, r_text
        # the literal configs[1] setting — ONE clip per forward call (the reference's batch 1): latency-bound, reported beside the
        # throughput headline.  Same engine, same two-slot pipelined API, clip 0 of the batch.
        one = [eng.alloc_outputs(1, T, H, W, L, want) for _ in range(2)]
        v1, w1, t1, s1 = d_vis[:1].contiguous(), d_vid[:1].contiguous(), d_text[:1].contiguous(), d_sizes[:1].contiguous()
        one_i = [0]

        def step_one():
            slot = one_i[0] & 1
            eng.forward_async(v1, w1, t1, d_pos, ori_sizes_hw=s1, outs=one[slot], slot=slot)
            one_i[0] += 1

        n_one = 50
        sec_one = timed(step_one, n_one, drain=drain_dev)

        def step_one_sync():
            eng.forward(v1, w1, t1, d_pos, ori_sizes_hw=s1, outs=one[0])
            torch.cuda.synchronize()

        sec_one_sync = timed(step_one_sync, n_one, device_events=False)
        front_end = {"value": total_clips / sec_raw, "unit": "clips/s", "ms_per_step": 1e3 * sec_raw / args.steps,
                     "what": "same step from RAW extractor maps resident in HBM: input_proj (2048->256) + input_proj2 (768->256) "
                             "+ text resizer fused into the forward (SURVEY 8f rank 2); adds 4.5 GFLOP and "
                             f"{raw_bytes / B / 1e6:.1f} MB of fp32 reads per clip",
                     "raw_input_bytes_per_step": raw_bytes,
                     "channels_last_bf16": {"value": total_clips / sec_cl, "unit": "clips/s", "ms_per_step": 1e3 * sec_cl / args.steps,
                                            "what": "same, maps given as channels-last bf16 [clips,T,H,W,C] (raw_layout = 1): half the bytes, "
                                                    "TMA-fed with no conversion pass"},
                     "from_token_ids": {"value": total_clips / sec_ids, "unit": "clips/s", "ms_per_step": 1e3 * sec_ids / args.steps,
                                        "what": f"as above, text from RoBERTa token ids: the {TEXT_TOWER[0]}-layer RoBERTa-base tower "
                                                f"({B} queries x {L} tokens per step) also runs inside the forward"}}
        batch1 = {"value": n_one * world / sec_one, "unit": "clips/s", "ms_per_clip_pipelined": 1e3 * sec_one / n_one,
                  "ms_per_clip_synchronous": 1e3 * sec_one_sync / n_one,
                  "what": "ONE clip per forward call (the reference's batch 1, BASELINE configs[1] read literally): CUDA-graph replay "
                          "of the same ≈455 launches, latency-bound; pipelined = two calls in flight, synchronous = host waits per clip"}
        return front_end, batch1

    def other_configs():
        """The other BASELINE.json shapes, timed in-run on their own engines (pipelined device path, CUDA graph, 5 steps):
        cfg-5 (384 px → 12x12, T = 128, L = 64) at 1 / 8 / 64 clips per step and the real yaml resolution (14x14, T = 64, L = 20)."""
        res = {}
        sd = O.synth_state_dict(0)

        def run(eng2, Bc, Tc, Hc, Wc, Lc, steps):
            g = torch.Generator(device="cuda").manual_seed(77 + rank)
            v = torch.randn(Bc, Tc, 256, Hc, Wc, device="cuda", generator=g)
            w = torch.randn(Bc, Tc, 256, Hc, Wc, device="cuda", generator=g)
            t = torch.randn(Bc, Lc, 256, device="cuda", generator=g)
            sz = torch.tensor([[360.0, 640.0]] * Bc, device="cuda")
            outs2 = [eng2.alloc_outputs(Bc, Tc, Hc, Wc, Lc, want) for _ in range(2)]
            k = [0]

            def step():
                slot = k[0] & 1
                eng2.forward_async(v, w, t, None, ori_sizes_hw=sz, outs=outs2[slot], slot=slot)   # pos generated in the library
                k[0] += 1

            sec2 = timed(step, steps, drain=lambda: (eng2.wait(0), eng2.wait(1)))
            fl = reference_flops(Tc, Hc, Wc, Lc)
            cps = Bc * steps / sec2
            return {"value": cps * world, "unit": "clips/s", "clips_per_step_per_gpu": Bc, "ms_per_step": 1e3 * sec2 / steps,
                    "algorithmic_gflop_per_clip": fl / 1e9, "algorithmic_tflops_per_gpu": cps * fl / 1e12,
                    "frac_of_sustained_bf16_peak": cps * fl / 1e12 / pk["bf16_tflops_sustained"]}

        e5 = GroundingEngine(sd, max_clips=64, max_frames=128, max_hw=144, max_text=64, use_cuda_graph=not args.no_graph)
        for Bc in (1, 8, 64):
            res[f"cfg5_T128_12x12_L64_B{Bc}"] = run(e5, Bc, 128, 12, 12, 64, 30 if Bc == 1 else 5)
        e5.close()
        torch.cuda.empty_cache()
        ey = GroundingEngine(sd, max_clips=16, max_frames=64, max_hw=196, max_text=20, use_cuda_graph=not args.no_graph)
        res["yaml_T64_14x14_L20_B16"] = run(ey, 16, 64, 14, 14, 20, 5)
        ey.close()
        torch.cuda.empty_cache()
        return res

    def sharded_clip(gname):
        """BASELINE configs[3]: ONE 256-frame clip, frames sharded over all ranks, exchanges on the device over
        NVLink peer memory (csrc/p2p_exchange.cu) inside the CUDA graph.  Parity against the reference golden of the same clip
        (tests/golden/ev_cfg4_* at 7x7, ev_long_* at 12x12: partial frame selections in both passes) is checked in the same run."""
        from vgqa_b200.parallel import forward_sharded_clip, shard_frames
        gp = os.path.join(ROOT, "tests", "golden", gname + ".npz")
        g = np.load(gp)
        Tc, Hc, Wc, Lc, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
        if Tc % world != 0:
            return {"skipped": f"T={Tc} is not divisible by {world} ranks"}
        sd = O.apply_calibration(O.synth_state_dict(seed, max_video_len=int(g["max_video_len"])), g)
        vis, vid, pos, text = O.synth_event_inputs(seed, Tc, Hc, Wc, Lc, amp=float(g["event_amp"]))
        s0, e0 = shard_frames(Tc, world, rank)
        es = GroundingEngine(sd, max_clips=1, max_frames=e0 - s0, max_hw=Hc * Wc, max_text=Lc, max_video_len=int(g["max_video_len"]),
                             use_cuda_graph=True)
        es.enable_p2p_sharding(rank, world)
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        a4 = (tt(vis[None, s0:e0]), tt(vid[None, s0:e0]), tt(text[None, :, 0]), tt(pos[:1]))
        out = forward_sharded_clip(es, *a4, ori_size_hw=(360, 640))
        err = max(float(np.abs(out["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max()),
                  float(np.abs(out["pred_sted"].cpu().numpy() - g["pred_sted"][0]).max()),
                  float(np.abs(out["pred_actioness"].cpu().numpy() - g["pred_actioness"][0, :, 0]).max()))
        r1 = np.zeros(Tc); r1[g["choose_pass1"]] = 1
        r2 = np.zeros(Tc); r2[g["choose_pass2"]] = 1
        sel = bool((out["choose1"].cpu().numpy() == r1).all() and (out["choose2"].cpu().numpy() == r2).all())
        si, ei = (int(x) for x in out["sted_idx"].cpu().numpy())
        fid = g["frame_ids"]
        sted_ok = [int(fid[si]), int(fid[ei]) + 1] == g["post_sted"][0].tolist()
        n = 20
        sec4 = timed(lambda: forward_sharded_clip(es, *a4, ori_size_hw=(360, 640), check_errors=False), n)
        perr = es.p2p_error()
        es.close()
        return {"ms_per_clip": 1e3 * sec4 / n, "ranks": world, "frames_per_rank": e0 - s0, "exchange": "peer memory (NVLink), CUDA graph",
                "max_abs_err_vs_golden": err, "selections_identical": sel, "sted_argmax_identical": sted_ok, "p2p_error": perr,
                "fixture": gname,
                "what": f"ONE {Tc}-frame {Hc}x{Wc} clip (BASELINE configs[3]) frame-sharded over all ranks; includes the gather of the "
                        f"per-rank outputs and PostProcess; the unsharded reference point is other_configs.long_T{Tc}_{Hc}x{Wc}_L{Lc}_B1 "
                        "of the 1-GPU run"}

    def long_clip_single_gpu(Hc):
        sd = O.synth_state_dict(0, max_video_len=256)
        e4 = GroundingEngine(sd, max_clips=1, max_frames=256, max_hw=Hc * Hc, max_text=20, max_video_len=256, use_cuda_graph=not args.no_graph)
        g = torch.Generator(device="cuda").manual_seed(5)
        v = torch.randn(1, 256, 256, Hc, Hc, device="cuda", generator=g)
        w = torch.randn(1, 256, 256, Hc, Hc, device="cuda", generator=g)
        t = torch.randn(1, 20, 256, device="cuda", generator=g)
        sz = torch.tensor([[360.0, 640.0]], device="cuda")
        o4 = e4.alloc_outputs(1, 256, Hc, Hc, 20, want)

        def step():
            e4.forward(v, w, t, None, ori_sizes_hw=sz, outs=o4)
            torch.cuda.synchronize()

        n = 20
        sec4 = timed(step, n, device_events=False)
        e4.close()
        return {"ms_per_step": 1e3 * sec4 / n, "clips_per_step_per_gpu": 1, "what": f"ONE 256-frame {Hc}x{Hc} clip on one GPU, synchronous "
                "(the unsharded reference point of the sharded_* lines of the N > 1 runs)"}

    def eager_pytorch():
        """The reference's own modules (oracle/_ref) in eager PyTorch under bf16 autocast on this B200, one clip per forward as the
        reference runs them (grounding_net.py:108-110): the only pre-existing implementation on the same box."""
        if not reference_arm_available():
            return {"unavailable": "oracle/_ref (byte-compiled reference modules) is not present"}
        ref = ReferenceHotPath(0, "cuda")
        inp = ref.inputs(0)
        res = {}
        for name, ac in (("bf16_autocast", True), ("fp32", False)):
            for _ in range(3):
                ref.run(inp, autocast=ac)
            torch.cuda.synchronize()
            n, t0 = 10, time.perf_counter()
            for _ in range(n):
                ref.run(inp, autocast=ac)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            res[name] = {"value": 1.0 / dt, "unit": "clips/s", "ms_per_clip": 1e3 * dt}
        res["what"] = ("reference PyTorch modules (ground_encoder, *_clas, ground_decoder x2, heads), eager, cuBLAS/ATen kernels, one "
                       f"clip per forward with its host syncs; torch {torch.__version__}")
        del ref
        torch.cuda.empty_cache()
        return res

    def swin_stage_line():
        """SURVEY §8f rank 3, first piece: the LAST stage of the Video-Swin-T extractor (`vid.layers[3]`: 2 SwinTransformerBlock3D,
        dim 768, 24 heads, window (8,7,7)) on this library's kernels (csrc/swin.cu), 8 clips x 64 frames at 7x7, next to the
        reference's own `BasicLayer` in eager PyTorch bf16-autocast on the same B200 (oracle/_ref)."""
        clips = 8
        sd = O.synth_state_dict(0)
        sd.update(O.synth_swin_stage(0))
        es = GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=49, max_text=8)
        g = torch.Generator(device="cuda").manual_seed(21)
        x = torch.randn(clips, T, 7, 7, 768, device="cuda", generator=g)
        n = 10
        sec_s = timed(lambda: es.swin_stage(x), n)
        launches_s = es.last_launch_count
        tokens = clips * T * 49
        flops = 2 * tokens * (2.0 * 768 * 2304 + 2.0 * 768 * 768 + 4.0 * 768 * 3072 + 4.0 * 392 * 768)
        res = {"ms_per_step": 1e3 * sec_s / n, "clips_per_step": clips, "value": clips * n / sec_s, "unit": "clips/s",
               "tflops": flops / (sec_s / n) / 1e12, "launches": launches_s,
               "what": "vid.layers[3] (2 Swin blocks: LN, qkv, 392-token window attention with relative position bias / shift mask on "
                       "tcgen05, proj, LN, fc1 + GELU, fc2) on 8 x 64 frames of 7x7x768; output = the channels-last bf16 vid_raw map"}
        es.close()
        try:
            if reference_arm_available():
                from make_golden_swin import load_swin_module
                M = load_swin_module()
                layer = M.BasicLayer(dim=768, depth=2, num_heads=24, window_size=(8, 7, 7), mlp_ratio=4.0, qkv_bias=True).eval()
                layer.load_state_dict({k: torch.from_numpy(v) for k, v in O.synth_swin_stage(0, prefix="").items()}, strict=False)
                layer.cuda()
                xt = x.permute(0, 4, 1, 2, 3).contiguous()

                def ref_step():
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                        layer(xt)

                sec_r = timed(ref_step, n)
                res["eager_pytorch_b200_ms_per_step"] = 1e3 * sec_r / n
                del layer
        except Exception as ex:   # noqa: BLE001
            res["eager_pytorch_b200_ms_per_step"] = f"unavailable: {type(ex).__name__}: {ex}"
        torch.cuda.empty_cache()
        return res

    def guarded(fn):
        try:
            return fn()
        except Exception as ex:   # noqa: BLE001 — reported in the line, must not lose the headline
            return {"error": f"{type(ex).__name__}: {ex}"}

    try:
        front_end, batch1 = secondary()
    except Exception as ex:   # noqa: BLE001 — reported in the line
        front_end = batch1 = {"error": f"{type(ex).__name__}: {ex}"}
    eng.close()
    torch.cuda.empty_cache()
    # (own engine: it carries the weights of the Video-Swin module's last stage)
    with_bb = guarded(lambda: measure_with_backbone(torch, rank=rank)) if (rank == 0 and not args.quick) else None
    others = guarded(other_configs) if not args.quick else None
    if world == 1 and not args.quick:
        others = dict(others or {}, long_T256_7x7_L20_B1=guarded(lambda: long_clip_single_gpu(7)),
                      long_T256_12x12_L20_B1=guarded(lambda: long_clip_single_gpu(12)))
    sharded = guarded(lambda: sharded_clip("ev_cfg4_T256_7x7_L20_s0")) if world > 1 else None
    sharded_long = guarded(lambda: sharded_clip("ev_long_T256_12x12_L20_s0")) if world > 1 else None
    eager_ref = guarded(eager_pytorch) if (rank == 0 and world == 1 and not args.quick) else None
    swin_line = guarded(swin_stage_line) if (world == 1 and not args.quick) else None
    full_fwd = guarded(lambda: measure_full_forward(torch)) if (rank == 0 and world == 1 and not args.quick) else None
    value = total_clips / sec
    e2e = total_clips / sec_e2e
    h2d_f32 = int(h_vis.numel() * 4 * 2 + h_text.numel() * 4 + h_pos.numel() * 4 + h_sizes.numel() * 4)
    h2d = int(h_vis_cl.numel() * 2 * 2 + h_text.numel() * 4 + h_pos.numel() * 4 + h_sizes.numel() * 4)
    d2h = int(sum(v.numel() * v.element_size() for v in h_outs.values()))

    if rank == 0:
        flops_clip = reference_flops(T, H, W, L)
        roof = time_dominant_kernel(B, pk)
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step_per_gpu": B, "T": T, "H": H, "W": W, "L": L,
                       "l2_policy": f"inputs larger than L2: {h2d_f32 / 1e6:.0f} MB of fp32 features per step",
                       "cuda_graph": not args.no_graph, "parallelism": f"clips partitioned over {world} GPU(s), no data-path collective"},
            "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * sec_e2e / args.steps,
                    "input_format": "pinned host buffers; projected maps as channels-last bf16 [clips,T,H,W,256] (vgqa_inputs.feat_layout = 1), "
                                    "text / pos / sizes fp32",
                    "h2d_gbs_per_rank": h2d / (sec_e2e / args.steps) / 1e9,
                    "fp32_nchw_inputs": {"value": total_clips / sec_e2e_f32, "unit": "clips/s", "h2d_bytes_per_step": h2d_f32,
                                         "ms_per_step": 1e3 * sec_e2e_f32 / args.steps,
                                         "what": "same call with the maps in the reference's fp32 NCHW layout (twice the upload)"}},
            "front_end": front_end,
            "batch1": batch1,
            "with_backbone": with_bb,
            "full_forward_cfg1": full_fwd,
            "other_configs": others,
            "sharded_cfg4": sharded,
            "sharded_long": sharded_long,
            "eager_pytorch_b200": eager_ref,
            "swin_stage4": swin_line,
            "gpu_launches": int(launches) * args.steps,
            "gpu_launches_per_step": int(launches),
            "clocks": clocks,
            "roofline": roof,
            # second half of BASELINE.json's metric: an ncu figure (it cannot be measured inside a timed run): read from the tracked
            # profiles/static_metrics.json, which names the commit / command it was taken with; null when absent
            "decoder_tensor_pipe_util_pct": static_metrics().get("decoder_tensor_pipe_util_pct"),
            "algorithmic_gflop_per_clip": flops_clip / 1e9,
            "algorithmic_tflops_whole_path": value / world * flops_clip / 1e12,
            "frac_of_sustained_bf16_peak_whole_path": value / world * flops_clip / 1e12 / pk["bf16_tflops_sustained"],
        }
        if world == 1:
            v, times, kind, what, cores = cpu_clips_per_sec(args.cpu_clips)
            line["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": cores, "kind": kind,
                                    "sample": f"{args.cpu_clips} clips of the same workload, one per forward ({what}), {sum(times):.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
