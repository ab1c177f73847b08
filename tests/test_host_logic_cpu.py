"""CPU: host-side logic of the product (no CUDA): tube interpolation / predict() merge against the reference goldens
and the oracle, clip partitioning, and the world_size-2 gloo gather."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_path
from oracle import vgqa_oracle as O
from vgqa_b200 import postprocess as PP
from vgqa_b200.parallel import gather_predictions, partition_clips


def test_linear_interp_matches_reference_golden():
    g = np.load(golden_path("interp"))
    fids = g["fids"].tolist()
    bi = PP.linear_interp({f: [g["boxes"][i].tolist()] for i, f in enumerate(fids)})
    ci = PP.linear_interp_conf({f: [float(g["conf"][i])] for i, f in enumerate(fids)})
    assert sorted(bi) == g["out_fids"].tolist()
    np.testing.assert_allclose(np.asarray([bi[f][0] for f in sorted(bi)]), g["out_boxes"], rtol=1e-12)
    np.testing.assert_allclose(np.asarray([ci[f][0] for f in sorted(ci)]), g["out_conf"], rtol=0)


def test_interp_edge_cases_match_oracle():
    for d in ({5: [[1.0, 2.0, 3.0, 4.0]]}, {}, {0: [[0.0, 0.0, 0.0, 0.0]], 7: [[7.0, 14.0, 21.0, 28.0]], 8: [[1.0, 1.0, 1.0, 1.0]]}):
        assert PP.linear_interp({k: [list(v[0])] for k, v in d.items()}) == O.linear_interp({k: [list(v[0])] for k, v in d.items()})
    c = {0: [0.1], 5: [0.9], 6: [0.3]}
    assert PP.linear_interp_conf(dict(c)) == O.linear_interp_conf(dict(c))
    with pytest.raises(TypeError):            # non-integer frame ids fail exactly like the reference's range()
        PP.linear_interp({0: [[0, 0, 0, 0]], 2.5: [[1, 1, 1, 1]]})


def test_merge_predictions_schema_and_values():
    p1 = ({0: {0: [[0.0, 0.0, 10.0, 10.0]], 4: [[4.0, 4.0, 14.0, 14.0]]}}, {0: {0: [0.5], 4: [0.7]}}, {0: {"sted": [0, 5], "qtype": "declar"}}, {})
    p2 = ({0: {2: [[2.0, 2.0, 12.0, 12.0]], 6: [[6.0, 6.0, 16.0, 16.0]]}}, {0: {2: [0.6], 6: [0.8]}}, {0: {"sted": [2, 7], "qtype": "declar"}}, {})
    r = PP.merge_predictions(p1, p2, fps=2.0)
    ref = O.merge_predict((p1[0][0], p1[1][0], [0, 5]), (p2[0][0], p2[1][0], [2, 7]), fps=2.0)
    assert r == ref
    assert set(r) == {"temporal", "tube"} and set(r["temporal"]) == {"start", "end", "score"}
    assert r["temporal"]["score"] == 1.0 and [t["frame"] for t in r["tube"]] == list(range(7))
    assert all(set(t) == {"frame", "bbox", "score"} and len(t["bbox"]) == 4 for t in r["tube"])


def test_partition_clips_covers_everything_once():
    for n in (0, 1, 7, 16, 129):
        for w in (1, 2, 3, 8):
            spans = [partition_clips(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = partition_clips(5, world, rank)
    local = {f"vid{i}": {"sted": [i, i + 3], "rank": rank} for i in range(s, e)}
    merged = gather_predictions(local)
    ret[rank] = merged
    dist.barrier()
    dist.destroy_process_group()


def test_gather_predictions_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert set(ret[0]) == {f"vid{i}" for i in range(5)} and dict(ret[0]) == dict(ret[1])
    assert ret[0]["vid0"]["rank"] == 0 and ret[0]["vid4"]["rank"] == 1


# ---------------------------------------------------------------------------------------------------------------
# Frame sharding: the exchange plan of vgqa_set_sharding, checked on CPU with gloo (world_size 2) using the
# oracle's building blocks: (1) text-token mean = all-reduce of local sums, (2) temporal self-attention of the local
# queries over ALL frames = all-gather of the in-projected K|V rows, (3) masked means = all-reduce of weighted sums.
# ---------------------------------------------------------------------------------------------------------------
def _shard_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vgqa_b200.parallel import shard_frames
    rng = np.random.Generator(np.random.PCG64(3))
    T, L, d = 12, 5, 256
    text_tok = rng.standard_normal((L, T, d)).astype(np.float32)          # encoded text tokens of every frame
    tgt = rng.standard_normal((T, 1, d)).astype(np.float32)
    te = O.seq_embedding_sine(T + 1)[:T]
    in_w = (rng.standard_normal((768, 256)) / 16).astype(np.float32)
    in_b = (rng.standard_normal(768) * 0.05).astype(np.float32)
    out_w = (rng.standard_normal((256, 256)) / 16).astype(np.float32)
    out_b = np.zeros(256, np.float32)
    w = (rng.uniform(size=T) > 0.4).astype(np.float32)
    x = rng.standard_normal((T, 7)).astype(np.float32)
    s, e = shard_frames(T, world, rank)
    # (1) f_text_cls
    loc = torch.from_numpy(text_tok[:, s:e].sum(1))
    dist.all_reduce(loc)
    ok1 = np.allclose(loc.numpy() / T, text_tok.mean(1), atol=1e-5)
    # (2) temporal self-attention: local queries, gathered keys/values (rows of the packed in-projection)
    qk = tgt + te
    q_full = O.linear(qk, in_w[:256], in_b[:256]); k_full = O.linear(qk, in_w[256:512], in_b[256:512])
    v_full = O.linear(tgt, in_w[512:], in_b[512:])
    qkv_loc = torch.from_numpy(np.concatenate([q_full, k_full, v_full], -1)[s:e, 0])     # [T_loc, 768]
    gathered = [torch.empty_like(qkv_loc) for _ in range(world)]
    dist.all_gather(gathered, qkv_loc)
    kv = torch.cat(gathered).numpy()                                                     # rank-major == frame order
    dh = 32
    ql = qkv_loc.numpy()[:, :256].reshape(-1, 8, dh).transpose(1, 0, 2) * dh ** -0.5
    kl = kv[:, 256:512].reshape(-1, 8, dh).transpose(1, 0, 2)
    vl = kv[:, 512:].reshape(-1, 8, dh).transpose(1, 0, 2)
    att = (O.softmax(ql @ kl.transpose(0, 2, 1)) @ vl).transpose(1, 0, 2).reshape(-1, 256)
    mine = O.linear(att, out_w, out_b)
    ref = O.torch_mha(qk, qk, tgt, in_w, in_b, out_w, out_b, 8)[s:e, 0]
    ok2 = np.allclose(mine, ref, atol=1e-4)
    # (3) masked mean over the chosen frames
    red = torch.from_numpy(np.concatenate([(w[s:e, None] * x[s:e]).sum(0), [w[s:e].sum()]]).astype(np.float32))
    dist.all_reduce(red)
    ok3 = np.allclose(red[:-1].numpy() / red[-1].item(), x[w > 0].mean(0), atol=1e-5)
    ret[rank] = (ok1, ok2, ok3)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_exchange_plan_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == (True, True, True) and ret[1] == (True, True, True)
