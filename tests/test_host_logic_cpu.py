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
