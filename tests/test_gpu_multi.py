"""Two-GPU, NCCL: the time-sharded path end to end (needs >= 2 CUDA devices; skipped otherwise), and
mixed use of the per-frame and batched entry points on one handle."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

from conftest import assert_nms_equivalent, load_det_case, ragged_get

pytestmark = pytest.mark.gpu


def _cfg(c):
    from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg
    return BinaryCfg(BinaryCoreCfg(c["adaptive"], c["init_value"], c["sensitivity"], c["area"], c["interval"]),
                     HoughLineCfg(*c["hough"]), DynamicCfg(c["dy_mask"], 5))


def _worker(rank, world, port, outdir, name):
    import torch
    import torch.distributed as dist
    from metdetpy_b200 import sharding as S
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    g = load_det_case(name)
    cfg, n, T = _cfg(g["cfg"]), g["n"], len(g["frames"])
    shard = S.plan_shards(T, world, n)[rank]
    eng = S.CudaEngine(g["mask"], n, g["fps"], cfg, device=rank, max_batch=16)
    res, dst, records, thr = S.detect_sharded(eng, g["frames"][shard.halo_start:shard.end], shard, T, n, cfg,
                                              device=torch.device("cuda", rank), want_dst=True)
    with open(os.path.join(outdir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(dict(shard=shard, dst=dst, thr=thr[0], records=records,
                         res=[(np.asarray(l).reshape(-1, 4), np.asarray(c)) for l, c in res]), f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_nccl_sharded_equals_reference_golden():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    name = "synth_384x216_n12_dyon_mask"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d, name), nprocs=2, join=True)
        out = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(2)]
    g = load_det_case(name)
    total = 0
    for o in out:
        sh = o["shard"]
        assert np.array_equal(o["thr"], g["bi_threshold"][sh.start:sh.end])
        assert np.array_equal(o["dst"], g["dst"][sh.start:sh.end])
        for i, (lines, cls) in enumerate(o["res"]):
            t = sh.start + i
            ref = ragged_get(g["nms_lines"], g["nms_offs"], t)
            refc = ragged_get(g["cls_pred"], g["nms_offs"], t)
            raw = ragged_get(g["raw_lines"], g["raw_offs"], t)
            assert_nms_equivalent(lines, cls.reshape(-1, 10)[:, -1], ref, refc[:, -1], raw, t)
            total += len(lines)
    assert out[1]["records"] is None and len(out[0]["records"]) == total > 0
    assert [r[0] for r in out[0]["records"]] == sorted(r[0] for r in out[0]["records"])


def test_mixed_per_frame_and_batched_calls_on_one_handle():
    """update()/detect() and detect_many() interleaved on the same detector == the reference sequence
    (the generic and the streaming kernels share the frame ring and the act ring)."""
    from metdetpy_b200.detector import M3Detector
    g = load_det_case("synth_384x216_n12_dyon_mask")
    det = M3Detector(g["n"] / g["fps"] + 1e-9, g["fps"], g["mask"], 10, _cfg(g["cfg"]), None, max_batch=9)
    T, t = len(g["frames"]), 0
    plan = [("one", 3), ("many", 9), ("one", 1), ("many", 4), ("many", 9), ("one", 5)]
    k = 0
    while t < T:
        kind, cnt = plan[k % len(plan)]
        k += 1
        cnt = min(cnt, T - t)
        if kind == "one":
            for _ in range(cnt):
                det.update(g["frames"][t])
                lines, cls = det.detect()
                assert det.bi_threshold == g["bi_threshold"][t], t
                assert np.array_equal(det.dst, g["dst"][t]), t
                assert np.array_equal(np.asarray(det.linesp_ext).reshape(-1, 4),
                                      ragged_get(g["raw_lines"], g["raw_offs"], t)), t
                t += 1
        else:
            res, dst = det.detect_many(g["frames"][t:t + cnt], return_dst=True)
            for i in range(cnt):
                assert det.last_infos[i]["bi_threshold"] == g["bi_threshold"][t + i], t + i
                assert np.array_equal(dst[i], g["dst"][t + i]), t + i
                assert np.array_equal(det.last_raw[i].reshape(-1, 4), ragged_get(g["raw_lines"], g["raw_offs"], t + i))
            t += cnt
