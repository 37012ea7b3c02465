"""Times the two head-tail implementations (egr_head_tail_stage impl 0 / 1) at the headline geometry: 4 views x 64 frames."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from egorear_b200 import _lib  # noqa: E402

lib = _lib.load()
B, G, J = 64, 4, 15
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn((G, B, 1024, 128), generator=g, device="cuda").half()
w = torch.randn((2, J, 128), generator=g, device="cuda") * 128 ** -0.5
bias = torch.randn((2, J), generator=g, device="cuda")
hm = torch.empty((B, G, J, 64, 64), device="cuda")
hm_t = torch.empty((G, B, J, 4096), device="cuda", dtype=torch.float16)
sel = (ctypes.c_int32 * 4)(0, 0, 1, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for precise in (1, 0):
    for impl in (0, 1):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.egr_head_tail_stage(ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(bias.data_ptr()),
                                               sel, B, G, J, ctypes.c_void_p(hm.data_ptr()), G * J * 4096, J * 4096,
                                               ctypes.c_void_p(hm_t.data_ptr()), precise, impl, st))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        mb = (z.numel() * 2 + hm.numel() * 4 + hm_t.numel() * 2) / 1e6
        print("precise %d impl %d: %.1f us (median of %d), %.0f MB algorithmic -> %.0f GB/s" % (precise, impl, ts[len(ts) // 2] * 1e3, len(ts), mb, mb / ts[len(ts) // 2]))
