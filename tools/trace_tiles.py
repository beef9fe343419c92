import sys, os
os.environ["XVEC_TRACE"] = "1"
import importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
importlib.import_module("speaker-recognition-x-vectors_b200.build").build(debug=True)  # -DXVEC_DEBUG variant with the trace stamps
os.environ.setdefault("XVEC_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speaker-recognition-x-vectors_b200", "libxvec_b200_debug.so"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch, xvec_b200
from xvec_b200 import ops
lib = xvec_b200._lib.load()
rows, n = 74 * 256 * 8, 512
for cin in [int(v) for v in os.environ.get('KS','128,512').split(',')]:
    x = torch.randn(rows, cin, device="cuda").bfloat16()
    w = ops.pack_weight(torch.randn(n, cin, device="cuda") / cin ** 0.5, 1, cin, torch.bfloat16)
    b = torch.zeros(n, device="cuda")
    out = torch.empty(rows, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.tdnn_layer_flat(x, w, n, [0], b, None, None, relu=True, out=out)
    torch.cuda.synchronize()
    buf = np.zeros(32 * 32, dtype=np.int64)
    m = lib.xvec_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    tr = buf.reshape(32, 32)[:16]
    t0 = tr[0, 0]
    print(f"K={cin}: per tile stamps (cycles since first load issue): prod_first_load, mma_enter, mma_tempty_ok, mma_full0_ok, mma_commit_issued, epi_enter, epi_tfull_ok, epi_done")
    for it in range(4, 14):
        print(it, " ".join(f"{int(v - t0):7d}" for v in tr[it][:8]), "| epi rel tfull_ok:", " ".join(f"{int(v - tr[it][6]):6d}" for v in list(tr[it][8:11]) + list(tr[it][14:16]) + list(tr[it][11:14])), f"done {int(tr[it][7]-tr[it][6])}")
        print("      mma steps 4-7 [top, A ready, B ready, issued+committed] rel. to step-4 top:", " | ".join(" ".join(f"{int(tr[it][16+4*k+q]-tr[it][16]):5d}" for q in range(4)) for k in range(4)))
