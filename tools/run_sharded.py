"""Frame-sharded long clip on N GPUs (torchrun): parity against the cfg-4 golden (T = 256) and timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vgqa_oracle as O  # synthetic weights / inputs + golden comparison (test tool, not product)
from vgqa_b200.engine import GroundingEngine
from vgqa_b200.parallel import forward_sharded_clip, shard_frames

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# the fixture: by default the "decisive" 256-frame one (partial frame selections in both passes — the masked means and the
# selection counts then really cross the ranks); VGQA_SHARD_GOLDEN picks another
gname = os.environ.get("VGQA_SHARD_GOLDEN", "ev_cfg4_T256_7x7_L20_s0")
g = np.load(os.path.join(ROOT, "tests", "golden", gname + ".npz"))
T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
sd = O.apply_calibration(O.synth_state_dict(seed, max_video_len=int(g["max_video_len"])), g)
amp = float(g["event_amp"]) if "event_amp" in g.files else 0.0
vis, vid, pos, text = O.synth_event_inputs(seed, T, H, W, L, amp=amp) if amp > 0 else O.synth_inputs(seed, T, H, W, L)
s, e = shard_frames(T, world, rank)
# VGQA_SHARD_P2P=1: exchanges on the device over NVLink peer memory (CUDA-graph forward) instead of the NCCL callback (eager)
p2p = os.environ.get("VGQA_SHARD_P2P") == "1"
eng = GroundingEngine(sd, max_clips=1, max_frames=e - s, max_hw=H * W, max_text=L, max_video_len=int(g["max_video_len"]),
                      use_cuda_graph=p2p)
if p2p and world > 1:
    eng.enable_p2p_sharding(rank, world)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
args = (t(vis[None, s:e]), t(vid[None, s:e]), t(text[None, :, 0]), t(pos[:1]))
if world == 1:
    # reference point: the same clip unsharded on one GPU (CUDA graph when VGQA_SHARD_P2P=1, else eager)
    sizes = torch.tensor([[360.0, 640.0]], device="cuda")
    run = lambda: eng.forward(*args, ori_sizes_hw=sizes, want=["pred_boxes", "pred_sted", "boxes_px", "sted_idx"])
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    print(f"UNSHARDED T={T} on one GPU: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per clip ({'CUDA graph' if p2p else 'eager'})")
    dist.destroy_process_group()
    sys.exit(0)
out = forward_sharded_clip(eng, *args, ori_size_hw=(int(g["ori_size"][0]), int(g["ori_size"][1])))
torch.cuda.synchronize()
errs = {"pred_boxes": float(np.abs(out["pred_boxes"].cpu().numpy() - g["pred_boxes"]).max()),
        "pred_sted": float(np.abs(out["pred_sted"].cpu().numpy() - g["pred_sted"][0]).max()),
        "pred_actioness": float(np.abs(out["pred_actioness"].cpu().numpy() - g["pred_actioness"][0, :, 0]).max()),
        "att_sequences": float(np.abs(out["att_sequences"].cpu().numpy() - g["att_sequences"][0]).max()),
        "logits_r_a": float(np.abs(out["logits_r_a"].cpu().numpy() - g["logits_r_a"][0]).max())}
ref1 = np.zeros(T); ref1[g["choose_pass1"]] = 1
ref2 = np.zeros(T); ref2[g["choose_pass2"]] = 1
sel_ok = bool((out["choose1"].cpu().numpy() == ref1).all()) and bool((out["choose2"].cpu().numpy() == ref2).all())
fid = g["frame_ids"]
si, ei = (int(x) for x in out["sted_idx"].cpu().numpy())
sted_ok = [int(fid[si]), int(fid[ei]) + 1] == g["post_sted"][0].tolist()
for _ in range(3):
    forward_sharded_clip(eng, *args, ori_size_hw=(360, 640), check_errors=False)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
n = 10
for _ in range(n):
    forward_sharded_clip(eng, *args, ori_size_hw=(360, 640), check_errors=False)
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / n
if rank == 0:
    ok = all(v <= 2e-2 for v in errs.values()) and sel_ok and sted_ok
    print(f"SHARDED world={world} {gname} T={T} ({e - s} frames/rank, K1={len(g['choose_pass1'])}, K2={len(g['choose_pass2'])}): max-abs errors {errs} selection_identical={sel_ok} "
          f"sted_argmax_identical={sted_ok} -> {'PASS' if ok else 'FAIL'}; {dt * 1e3:.2f} ms per clip ({'peer-memory exchange, CUDA graph' if p2p else 'NCCL callback, eager'}, {world} GPUs), p2p_error={eng.p2p_error()}")
dist.destroy_process_group()
