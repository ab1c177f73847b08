"""Golden fixture of the WHOLE Video-Swin-T extractor from the REAL reference class (build container only):

    python tests/golden/make_golden_swin_full.py        # writes tests/golden/swin_full_*.npz

`vidswin_model('video_swin_t_p4w7')` (VideoSwinTransformerBackbone, vgqa/core/vision/video_swin_transformer.py:626-685; the `self.vid`
of VSTGNet) is constructed, loaded with `synth_swin_backbone(seed)` and run on seeded frames; the stage-4 map ('3') is stored in full
(fp16) and every 97th token row of the four stage outputs in fp32, together with torch's own bf16-autocast deviation (the yardstick of
the CUDA path's tolerance)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_golden_swin import load_swin_module  # noqa: E402
from oracle import vgqa_oracle as O  # noqa: E402

CASES = [("swin_full_T16_224_s0", 1, 16, 224, 0),
         ("swin_full_T12_256_s1", 1, 12, 256, 1)]      # 12 frames → 16, maps 64 / 32 / 16 / 8 → 70 / 35 / 21 / 14: the padding path


def swin_frames(seed, clips, T, R):
    rng = np.random.Generator(np.random.PCG64(14000 + seed))
    return rng.standard_normal((clips * T, 3, R, R), dtype=np.float32)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    M = load_swin_module()
    for name, clips, T, R, seed in CASES:
        model = M.vidswin_model("video_swin_t_p4w7", None).eval()
        sd = O.synth_swin_backbone(seed, prefix="")
        missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not unexpected and all("relative_position_index" in m for m in missing), (missing, unexpected)
        x = swin_frames(seed, clips, T, R)
        with torch.no_grad():
            out = model(torch.from_numpy(x), T)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                outb = model(torch.from_numpy(x), T)
        rec = dict(clips=clips, T=T, R=R, seed=seed, torch_version=torch.__version__)
        mine = O.video_swin_backbone({("vid." + k): v for k, v in sd.items()}, x, clips)
        for s in range(4):
            y = out[str(s)]                                                   # (clips*T, C, H, W)
            ycl = y.reshape(clips, T, *y.shape[1:]).permute(0, 1, 3, 4, 2).contiguous().numpy()   # channels-last
            ac = (outb[str(s)].float() - y).abs()
            rows = ycl.reshape(-1, ycl.shape[-1])[::97].copy()
            rec[f"rows{s}"] = rows
            rec[f"autocast_err_mean{s}"] = np.float32(ac.mean())
            rec[f"autocast_err_max{s}"] = np.float32(ac.max())
            rec[f"abs_mean{s}"] = np.float32(np.abs(ycl).mean())
            print(name, "stage", s, ycl.shape, "|y| mean", float(np.abs(ycl).mean()), "oracle max-abs diff", float(np.abs(mine[s] - ycl).max()),
                  "autocast err mean/max", float(ac.mean()), float(ac.max()))
            if s == 3:
                rec["y3"] = ycl.astype(np.float16)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
