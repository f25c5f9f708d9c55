"""Probe: does this box support NVLink multicast objects (multimem.* instructions)?  Single process, all visible devices."""
import ctypes as C, sys
cu = C.CDLL("libcuda.so.1")
def ck(r, what):
    if r != 0:
        s = C.c_char_p()
        cu.cuGetErrorString(r, C.byref(s))
        print("FAIL %s: %d %s" % (what, r, (s.value or b"").decode()))
        sys.exit(0)
ck(cu.cuInit(0), "cuInit")
n = C.c_int()
ck(cu.cuDeviceGetCount(C.byref(n)), "count")
print("devices", n.value)
CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED = 132
CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FD = 103
CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC = 128
for d in range(n.value):
    dev = C.c_int()
    ck(cu.cuDeviceGet(C.byref(dev), d), "get")
    v = C.c_int()
    for name, a in (("multicast", CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED), ("posix_fd", CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FD), ("fabric", CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC)):
        ck(cu.cuDeviceGetAttribute(C.byref(v), a, dev), name)
        print(" dev", d, name, v.value)
if n.value < 2:
    sys.exit(0)
# try to create a multicast object over all devices
class McProp(C.Structure):
    _fields_ = [("numDevices", C.c_uint), ("size", C.c_size_t), ("handleTypes", C.c_ulonglong), ("flags", C.c_ulonglong)]
prop = McProp(n.value, 0, 1, 0)  # CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR = 1
gran = C.c_size_t()
cu.cuMulticastGetGranularity.argtypes = [C.POINTER(C.c_size_t), C.POINTER(McProp), C.c_int]
ck(cu.cuMulticastGetGranularity(C.byref(gran), C.byref(prop), 1), "granularity")   # RECOMMENDED = 1
print("granularity", gran.value)
prop.size = max(gran.value, 2 << 20)
h = C.c_ulonglong()
cu.cuMulticastCreate.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(McProp)]
ck(cu.cuMulticastCreate(C.byref(h), C.byref(prop)), "cuMulticastCreate")
print("multicast object created")
ctxs = []
for d in range(n.value):
    dev = C.c_int(); cu.cuDeviceGet(C.byref(dev), d)
    ctx = C.c_void_p()
    ck(cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev), "ctx")
    ctxs.append(ctx)
    ck(cu.cuMulticastAddDevice(C.c_ulonglong(h.value), dev), "addDevice %d" % d)
print("all devices added")
fd = C.c_int()
r = cu.cuMemExportToShareableHandle(C.byref(fd), C.c_ulonglong(h.value), 1, C.c_ulonglong(0))
print("export to fd:", r, fd.value)
