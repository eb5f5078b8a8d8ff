"""BASELINE configs 2-5 at their full frame sizes through the CPU-emulated streaming path against the CPU checker
(about six minutes and 5 GB of host memory; output of the last run: profiles/r02_emu_fullsize_configs.txt).  Test tooling."""
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np, cv2
import emu_fuzz as F
from metdetpy_b200 import synth
lib=F.build_lib(tempfile.mkdtemp())
print("# BASELINE configs at full frame size through the CPU-emulated streaming path (tests/emu_fuzz.py run_case) vs the CPU checker (cv2 backend); None = identical",flush=True)
for name,(W,H,fps,n,T,B,dy,masked) in {"config2":(1920,1080,30,5,12,6,False,False),"config3":(3840,2160,30,30,34,4,True,False),
                            "config4 (n=60, polygon mask applied inside the kernels)":(3840,2160,60,60,64,8,True,True),"config5":(7680,4320,30,30,33,3,True,False)}.items():
    t=time.time()
    fr=synth.make_stream(T,W,H,fps,speed_scale=3.0,thickness=2)
    mask=np.ones((H,W),np.uint8)
    if masked:
        cv2.fillPoly(mask,[np.array([[0,H],[0,int(H*0.7)],[int(W*0.6),H]],np.int32)],0)
    case=dict(W=W,H=H,n=n,T=T,batch=B,cfg=dict(adaptive=True,init_value=7,sensitivity="normal",area=0.1,interval=2,hough=(10,10,10),dy_mask=dy),
              mask=mask,frames=fr*mask,raw_frames=fr,apply_mask=masked)
    r=F.run_case(lib,case); print(name,dict(W=W,H=H,n=n,frames=T,batch=B,dy_mask=dy),"->",r,"(%.0f s)"%(time.time()-t),flush=True)
    if name=="config3":  # the per-frame resident-state path (update(); detect() frame by frame) at the bench's frame size
        t=time.time(); r=F.run_case(lib,case,per_frame=True)
        print(name,"per-frame API path",dict(W=W,H=H,n=n,frames=T),"->",r,"(%.0f s)"%(time.time()-t),flush=True)
