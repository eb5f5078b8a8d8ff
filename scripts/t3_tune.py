"""Times the temporal pass (stack -> diff -> threshold) of one 4K batch for every kernel generation / temporal3 shape
variant, alone on the GPU (one batch in flight), with CUDA events inside the library (mdb_get_info "temporal_ms").
usage: python scripts/t3_tune.py [W H n B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metdetpy_b200 import BinaryCfg, BinaryCoreCfg, DynamicCfg, HoughLineCfg, synth
from metdetpy_b200.detector import M3Detector

W, H, n, B = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (3840, 2160, 30, 512)))
fps = float(n)
dev = torch.device("cuda", 0)
cfg = BinaryCfg(BinaryCoreCfg(True, 7, "normal", 0.1, 2), HoughLineCfg(10, 10, 10), DynamicCfg(True, 5))
det = M3Detector(n / fps + 1e-9, fps, np.ones((H, W), np.uint8), 10, cfg, None, max_batch=B)
batches = [synth.make_stream_device(B, W, H, 30.0, dev, t0=s * B, loop=2 * B, quiet=n) for s in range(2)]
torch.cuda.synchronize()
peak = 6533.2


def run(label, opts, reps=6):
    for k, v in opts.items():
        det._eng.set_option(k, v)
    t, sp = [], []
    for r in range(reps):
        det.submit(batches[r % 2].data_ptr(), B, True)
        det.collect(want_lines=False)
        t.append(det._eng.info("temporal_ms")); sp.append(det._eng.info("spatial_ms"))
    gen = int(det._eng.info("temporal_generation"))
    tm, sm = float(np.median(t[2:])), float(np.median(sp[2:]))
    hist = (n - 1) if B > n else 0
    gbs = (B + hist) * W * H * (1 + 1 / 8) / (tm * 1e-3) / 1e9
    print(f"{label:34s} gen {gen}  temporal {tm:7.3f} ms  spatial {sm:6.3f} ms  chain {tm + sm:7.3f} ms  "
          f"chain frac {2.0 * W * H * B / ((tm + sm) * 1e-3) / 1e9 / peak:5.3f}  temporal DRAM ~{gbs:6.0f} GB/s", flush=True)


only = [int(v) for v in os.environ.get("T3_ONLY", "").split(",") if v]
reps = int(os.environ.get("T3_REPS", "6"))
if only:
    for v in only:
        run(f"temporal3 variant {v}", {"temporal_version": 3, "t3_variant": v}, reps)
    sys.exit(0)
run("temporal2 (round 1)", {"temporal_version": 2})
run("temporal3 default", {"temporal_version": 3, "t3_variant": 0})
if n == 60:
    for v, name in [(1, "U30 P2 BL10 K10 minb3"), (2, "U30 P2 BL10 K15 minb3"), (3, "U30 P2 BL15 K15 minb3"),
                    (4, "U30 P2 BL15 K10 minb3"), (5, "U20 P3 BL10 K10 minb3"), (6, "U15 P4 BL15 K15 minb2"),
                    (7, "U20 P3 BL20 K10 minb2"), (8, "U30 P2 BL10 K6 minb3")]:
        run(f"temporal3 variant {v}: {name}", {"temporal_version": 3, "t3_variant": v})
if n == 30:
    for v, name in [(1, "BL15 K5 minb4"), (2, "BL10 K6 minb3"), (3, "BL10 K5 minb4"), (4, "BL15 K6 minb3"),
                    (5, "BL10 K10 minb3"), (6, "BL6 K6 minb4"), (7, "bulk BL10 K5 minb4"), (8, "bulk BL10 K5 minb3"),
                    (9, "bulk BL15 K5 minb3"), (10, "bulk BL10 K10 minb4"), (11, "bulk BL6 K6 minb4"),
                    (12, "bulk BL10 K3 minb4"), (13, "L2pf BL10 K5 minb4"), (14, "L2pf BL10 K3 minb4"),
                    (15, "L2pf BL10 K6 minb4"), (16, "L2pf BL10 K10 minb3"), (17, "L2pf BL10 K2 minb4"),
                    (18, "U15 P2 BL15 K5 minb4"), (19, "U30 P1 BL10 K6 minb4 (spills)"), (20, "U15 P2 BL5 K5 minb5"),
                    (21, "U15 P2 BL5 K15 minb5"), (22, "U15 P2 BL15 K15 minb5"), (23, "BL10 K15 minb3")]:
        run(f"temporal3 variant {v}: {name}", {"temporal_version": 3, "t3_variant": v})
