"""NVDEC probe, second pass: driver API context made current explicitly, then cuvidGetDecoderCaps and cuvidCreateDecoder."""
import ctypes as C, os
print("NVIDIA_DRIVER_CAPABILITIES =", os.environ.get("NVIDIA_DRIVER_CAPABILITIES"))
cu = C.CDLL("libcuda.so.1")
print("cuInit", cu.cuInit(0))
ver = C.c_int(); cu.cuDriverGetVersion(C.byref(ver)); print("driver API version", ver.value)
dev = C.c_int(); print("cuDeviceGet", cu.cuDeviceGet(C.byref(dev), 0))
ctx = C.c_void_p(); print("cuDevicePrimaryCtxRetain", cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev))
print("cuCtxSetCurrent", cu.cuCtxSetCurrent(ctx))
nv = C.CDLL("libnvcuvid.so.1")


class CAPS(C.Structure):
    _fields_ = [("eCodecType", C.c_int), ("eChromaFormat", C.c_int), ("nBitDepthMinus8", C.c_uint), ("reserved1", C.c_uint * 3),
                ("bIsSupported", C.c_ubyte), ("nNumNVDECs", C.c_ubyte), ("nOutputFormatMask", C.c_ushort),
                ("nMaxWidth", C.c_uint), ("nMaxHeight", C.c_uint), ("nMaxMBCount", C.c_uint),
                ("nMinWidth", C.c_ushort), ("nMinHeight", C.c_ushort), ("bIsHistogramSupported", C.c_ubyte),
                ("nCounterBitDepth", C.c_ubyte), ("nMaxHistogramBins", C.c_ushort), ("reserved3", C.c_uint * 10)]


for codec, name in [(4, "H264"), (8, "HEVC"), (11, "AV1")]:
    c = CAPS(); c.eCodecType = codec; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0
    rc = nv.cuvidGetDecoderCaps(C.byref(c))
    print(f"{name}: rc={rc} supported={c.bIsSupported} nvdecs={c.nNumNVDECs} max={c.nMaxWidth}x{c.nMaxHeight}")
print("devices:", [d for d in os.listdir("/dev") if d.startswith("nvidia")])
