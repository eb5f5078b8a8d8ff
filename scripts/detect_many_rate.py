import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from metdetpy_b200 import BinaryCfg, synth
from metdetpy_b200.detector import M3Detector
B=512; W,H,n=3840,2160,30
det=M3Detector(n/30+1e-9,30,np.ones((H,W),np.uint8),10,BinaryCfg(),None,max_batch=B)
dev=torch.device("cuda",0)
x=synth.make_stream_device(B,W,H,30,dev,t0=0)
torch.cuda.synchronize()
for _ in range(2): det.detect_many((x.data_ptr(),B),on_device=True)
t0=time.perf_counter()
for _ in range(10): r=det.detect_many((x.data_ptr(),B),on_device=True)
dt=time.perf_counter()-t0
print("detect_many (synchronous API) frames/s:", round(10*B/dt))
