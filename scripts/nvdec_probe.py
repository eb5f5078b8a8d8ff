"""Is the hardware video decoder (NVDEC through libnvcuvid) usable on this box?  Prints what it finds; the on-device
decode feed of SURVEY 8(f) row 4 (videowrapper.py:90-167) stands or falls with this library, which ships with the
driver, not with the CUDA toolkit."""
import ctypes as C, ctypes.util, glob, os, subprocess
print("ldconfig:", [l.strip() for l in subprocess.run(["ldconfig", "-p"], capture_output=True, text=True).stdout.splitlines()
                    if "nvcuvid" in l or "nvidia-encode" in l or "libcuda.so" in l])
hits = []
for root in ("/usr/lib", "/usr/lib64", "/usr/local", "/opt", "/lib"):
    hits += glob.glob(root + "/**/libnvcuvid*", recursive=True)
print("files:", hits)
lib = None
for name in ["libnvcuvid.so.1", "libnvcuvid.so"] + hits:
    try:
        lib = C.CDLL(name); print("loaded", name); break
    except OSError as e:
        print("cannot load", name, "-", e)
if lib is None:
    print("RESULT: libnvcuvid is not present -> NVDEC is not usable from this image")
    raise SystemExit(0)
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")


class CAPS(C.Structure):  # CUVIDDECODECAPS (nvcuvid.h)
    _fields_ = [("eCodecType", C.c_int), ("eChromaFormat", C.c_int), ("nBitDepthMinus8", C.c_uint), ("reserved1", C.c_uint * 3),
                ("bIsSupported", C.c_ubyte), ("nNumNVDECs", C.c_ubyte), ("nOutputFormatMask", C.c_ushort),
                ("nMaxWidth", C.c_uint), ("nMaxHeight", C.c_uint), ("nMaxMBCount", C.c_uint),
                ("nMinWidth", C.c_ushort), ("nMinHeight", C.c_ushort), ("bIsHistogramSupported", C.c_ubyte),
                ("nCounterBitDepth", C.c_ubyte), ("nMaxHistogramBins", C.c_ushort), ("reserved3", C.c_uint * 10)]


for codec, name in [(4, "H264"), (8, "HEVC"), (10, "VP9"), (11, "AV1")]:
    c = CAPS(); c.eCodecType = codec; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0
    rc = lib.cuvidGetDecoderCaps(C.byref(c))
    print(f"{name}: rc={rc} supported={c.bIsSupported} nvdecs={c.nNumNVDECs} max={c.nMaxWidth}x{c.nMaxHeight}")
