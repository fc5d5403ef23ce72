"""Per-kernel times of the standalone NMS at one size (profile events, direct launches)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from dspnet_b200 import _lib, synth
from dspnet_b200.nms import nms_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = _lib.lib()
for kv in sys.argv[2:]:
    k, v = kv.split('=')
    L.dspmb_set_tuning(int(k), int(v))
d = torch.from_numpy(synth.nms_boxes(5, n)).cuda()
for _ in range(5):
    nms_device(d, 0.45)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    keep, num = nms_device(d, 0.45)
e1.record()
torch.cuda.synchronize()
print('N=%d  %.1f us per call, kept %d' % (n, e0.elapsed_time(e1) / 20 * 1e3, int(num.item())))
L.dspmb_profile_enable(1)
for _ in range(10):
    nms_device(d, 0.45)
torch.cuda.synchronize()
ms = (ctypes.c_float * 32)(); ln = (ctypes.c_int * 32)()
k = L.dspmb_profile_read(ms, ln, 32)
L.dspmb_profile_enable(0)
L.dspmb_profile_kernel_name.restype = ctypes.c_char_p
print({L.dspmb_profile_kernel_name(i).decode(): (round(ms[i] / 10 * 1e3, 1), ln[i] // 10) for i in range(k) if ln[i]})
