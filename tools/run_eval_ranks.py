"""Multi-GPU `do_eval`: torchrun --nproc-per-node N tools/run_eval_ranks.py [items] — every rank evaluates its contiguous share of a
synthetic dataset on its own GPU (no data-path collective), the prediction dicts are merged with all_gather_object and rank 0
prints the summary; the merged result must equal a single-rank evaluation of the same items."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import vgqa_oracle as O  # synthetic weights / inputs only
from vgqa_b200 import evaluate as E
from vgqa_b200.engine import GroundingEngine
from test_evaluate_gpu import make_items

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T2, H, W, L = 64, 7, 7, 20
items, gt = make_items(n, T2, H, W, L)
sd = O.synth_state_dict(0)
eng = GroundingEngine(sd, max_clips=16, max_frames=T2 // 2, max_hw=H * W, max_text=L, use_cuda_graph=True)
ev = E.VidSTGEvaluator(gt, [0.3, 0.5])
torch.cuda.synchronize()
t0 = time.perf_counter()
out = E.do_eval(eng, items, ev, clips_per_call=8, rank=rank, world=world)
dt = time.perf_counter() - t0
if rank == 0:
    ref_ev = E.VidSTGEvaluator(gt, [0.3, 0.5], distributed=False)
    ref = E.do_eval(eng, items, ref_ev, clips_per_call=8) if world > 1 else out
    same = all(abs(out[k] - ref[k]) < 1e-9 for k in ref) and ev.video_predictions == ref_ev.video_predictions if world > 1 else True
    print(f"do_eval world={world}: {n} items ({2 * n} clips of {T2 // 2} frames) in {dt * 1e3:.1f} ms; merged == single-rank: {same}; "
          f"declar_tiou={out['declar_tiou']:.4f} inter_viou={out['inter_viou']:.4f}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
