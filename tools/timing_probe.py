"""Phase timing of the SpMV kernel (needs a library built with -DSMLE_TIMING)."""
import ctypes as C, sys
sys.path.insert(0, "sparse-matrix-linear-equations_b200/python")
import torch, smle_b200 as S
L = S.lib()
S.init(0); st = torch.cuda.Stream(); S.set_stream(st.cuda_stream)
w = int(sys.argv[1]) if len(sys.argv) > 1 else 150
ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0); n = len(ro) - 1
a = S.CsrMatrix(ro, ci, va)
out = (C.c_longlong * 8)()
with torch.cuda.stream(st):
    x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    for _ in range(3): a.spmv(x, out=y)
    L.smle_debug_timing(out, 1)
    for _ in range(10): a.spmv(x, out=y)
    L.smle_debug_timing(out, 1)
    t = [int(v) for v in out]
    print("standalone: per tile cycles  mbar_wait %.0f  compute %.0f  barrier %.0f  issue %.0f (tiles %d); first->last tile span per CTA %.0f cycles" % (t[0]/t[3], t[1]/t[3], t[2]/t[3], t[4]/t[3], t[3]//10, (t[6]-t[5])/(292*10)))
    b = torch.rand(n, dtype=torch.float64, device="cuda"); xs = torch.empty_like(b)
    a.cg_run_fixed(b.view(n,1), xs.view(n,1), 16)
    L.smle_debug_timing(out, 1)
    a.cg_run_fixed(b.view(n,1), xs.view(n,1), 64)
    L.smle_debug_timing(out, 1)
    t = [int(v) for v in out]
    print("in CG     : per tile cycles  mbar_wait %.0f  compute %.0f  barrier %.0f" % (t[0]/t[3], t[1]/t[3], t[2]/t[3]))
