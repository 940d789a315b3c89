#!/usr/bin/env python
"""Timeline of one CTA of the tcgen05 sparse-conv kernel (debug hook crb3d_debug_set_tc_trace).
The clock stamps are compiled into csrc/spconv_tc.cu only with -DCRB3D_TC_TRACE (add it to the nvcc flags in __graft_entry__.py)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crb-active-3ddet_b200"))
import numpy as np, torch
from crb3d import _lib, ops, second, synth
dev = torch.device("cuda:0")
model = second.SECONDNet().eval().to_device(dev)
frames = [synth.make_frame(i) for i in range(4)]
offs = torch.from_numpy(np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)).to(dev)
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
books = model.geometry(pts, offs, 4)["rulebooks"]
lib = _lib.load()
lib.crb3d_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
for key, cin, cout in (("subm3", 64, 64), ("subm1", 16, 16)):
    d = books[key]
    feat = torch.randn(d.indices.shape[0], cin, device=dev)
    w = torch.randn(cout, d.nbr.shape[0], cin, device=dev) * 0.05
    for _ in range(3):
        ops.spconv_forward(feat, d.nbr, w, tf32=True)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    lib.crb3d_debug_set_tc_trace(ctypes.c_void_p(buf.data_ptr()))
    ops.spconv_forward(feat, d.nbr, w, tf32=True)
    torch.cuda.synchronize()
    lib.crb3d_debug_set_tc_trace(None)
    t = buf.cpu().numpy()
    a, b = t[:128].reshape(32, 4), t[128:].reshape(32, 4)
    t0 = b[0, 0]
    print(key, cin, cout, "rows", d.nbr.shape[1])
    print(" it | producer: stage free  lists+barrier  copies issued  arrive returned | mma: stage full   issued+committed")
    for it in range(27):
        if b[it, 0] == 0:
            break
        print("%3d | %10d %10d %10d %12d | %10d %12d" % (it, b[it, 0] - t0, b[it, 2] - t0, b[it, 3] - t0, b[it, 1] - t0, a[it, 2] - t0, a[it, 3] - t0))
