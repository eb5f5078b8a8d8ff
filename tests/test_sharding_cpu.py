"""CPU / gloo, world_size 2: host logic of the time-sharded path (metdetpy_b200/sharding.py) --
shard planning with the (2n-2)-frame halo, all-gather of integer noise sums, bit-identical threshold
replay on every rank, gather of line records to rank 0 -- with the CPU oracle standing in for the GPU
engine.  The result must equal one sequential pass over the whole stream."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth
from metdetpy_b200 import sharding as S

W, H, FPS, N, T = 160, 120, 30, 5, 64


def _cfg():
    return BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.2, 1), HoughLineCfg(8, 8, 6), DynamicCfg(True, 5))


def _stream():
    return synth.make_stream(T, W, H, FPS, speed_scale=4.0, thickness=2)


def _worker(rank, world, port, outdir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sharding_helpers import OracleEngine
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    frames = _stream()
    mask = np.ones((H, W), np.uint8)
    cfg = _cfg()
    shard = S.plan_shards(T, world, N)[rank]
    eng = OracleEngine(mask, N, FPS, cfg)
    res, dst, records, thr = S.detect_sharded(eng, frames[shard.halo_start:shard.end], shard, T, N, cfg,
                                              want_dst=True)
    with open(os.path.join(outdir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(dict(shard=shard, dst=dst, thr=thr, records=records,
                         lines=[np.asarray(l).reshape(-1, 4) for l, _ in res]), f)
    dist.barrier()
    dist.destroy_process_group()


def test_plan_and_schedule():
    sh = S.plan_shards(100, 3, 10)
    assert [(s.start, s.end, s.halo_start) for s in sh] == [(0, 34, 0), (34, 67, 16), (67, 100, 49)]
    assert [t for t in range(1, 45) if S.is_noise_sample(t, 5, 2)] == [2, 3, 4, 5, 10, 20, 30, 40]
    with pytest.raises(KeyError):
        S.replay_thresholds({}, 10, 5, adaptive=True, init_value=7, sensitivity="normal", interval=2)


def test_replay_matches_sequential_oracle():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sharding_helpers import OracleEngine, sequential_reference
    frames, mask, cfg = _stream(), np.ones((H, W), np.uint8), _cfg()
    thr_ref, snr_ref, _, _ = sequential_reference(frames, mask, N, FPS, cfg)
    eng = OracleEngine(mask, N, FPS, cfg)
    sums = eng.noise_sums(frames, 0)
    samples = {t: S.sigma_from_sums(int(sums[t - 1, 0]), int(sums[t - 1, 1]), min(N, t), eng.roi_pixels)
               for t in range(1, T + 1) if S.is_noise_sample(t, N, 1)}
    thr, thr_f, snr = S.replay_thresholds(samples, T, N, adaptive=True, init_value=7, sensitivity="normal", interval=1)
    assert np.array_equal(thr, thr_ref)
    assert np.allclose(snr, snr_ref, rtol=1e-12, atol=0)


def test_two_rank_gloo_equals_sequential():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from sharding_helpers import sequential_reference
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        out = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(2)]
    frames, mask, cfg = _stream(), np.ones((H, W), np.uint8), _cfg()
    thr_ref, snr_ref, dst_ref, lines_ref = sequential_reference(frames, mask, N, FPS, cfg)
    n_lines = 0
    for o in out:
        sh = o["shard"]
        assert np.array_equal(o["thr"][0], thr_ref[sh.start:sh.end])
        assert np.allclose(o["thr"][2], snr_ref[sh.start:sh.end], rtol=1e-12, atol=0)
        assert np.array_equal(o["dst"], dst_ref[sh.start:sh.end]), f"rank {sh.rank}: masks differ"
        for i, l in enumerate(o["lines"]):
            assert np.array_equal(l, lines_ref[sh.start + i][0]), (sh.rank, i)
            n_lines += len(l)
    assert out[1]["records"] is None
    rec = out[0]["records"]  # gathered to rank 0, in frame order
    flat = [[t, *row.tolist()] for t, (l, _) in enumerate(lines_ref) for row in l]
    assert [r[:5] for r in rec] == flat and n_lines == len(flat) and n_lines > 0


@pytest.mark.parametrize("n,interval,sens,adaptive", [(5, 1, "normal", True), (12, 2, "high", True), (30, 2, "low", True),
                                                      (7, 3, "normal", False)])
def test_native_replay_equals_python_replay(n, interval, sens, adaptive):
    """mdb_replay_thresholds (host C++ in the library, what the multi-GPU path uses) against the Python statement of
    the recurrence: identical thresholds, bit-identical doubles; sub-ranges; a missing sample is an error."""
    rng = np.random.default_rng(n)
    T, roi_px = 700, 37 * 53
    samples = []
    for tau in range(1, T + 1):
        if S.is_noise_sample(tau, n, interval):
            L = min(n, tau)
            N_ = L * roi_px
            mean = rng.uniform(0, 3)
            s1 = int(mean * N_)
            s2 = int((mean * mean + rng.uniform(0.5, 9)) * N_)
            samples.append((tau, s1, s2))
    sig = {tau: S.sigma_from_sums(s1, s2, min(n, tau), roi_px) for tau, s1, s2 in samples}
    thr, thr_f, snr = S.replay_thresholds(sig, T, n, adaptive=adaptive, init_value=6, sensitivity=sens, interval=interval)
    a, b, c = S.replay_thresholds_native(samples, roi_px, n, 0, T, adaptive=adaptive, init_value=6, sensitivity=sens,
                                         interval=interval)
    assert np.array_equal(a, thr) and np.array_equal(b, thr_f) and np.array_equal(c, snr)
    a, b, c = S.replay_thresholds_native(samples, roi_px, n, 123, 611, adaptive=adaptive, init_value=6, sensitivity=sens,
                                         interval=interval)
    assert np.array_equal(a, thr[123:611]) and np.array_equal(b, thr_f[123:611]) and np.array_equal(c, snr[123:611])
    if len(samples) > 3:
        with pytest.raises(ValueError):
            S.replay_thresholds_native(samples[:2] + samples[3:], roi_px, n, 0, T, adaptive=adaptive, init_value=6,
                                       sensitivity=sens, interval=interval)
