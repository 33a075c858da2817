import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import indelminer_b200
from indelminer_b200 import lib as _lib, synth
L = _lib.load()
for g in (0, 4):
    R = indelminer_b200.Realigner(device=0, numgaps=g)
    ref = synth.make_reference(64_000_000, seed=1)
    R.set_reference([ref.tobytes()])
    n = 262144
    w = synth.make_candidates(ref, n, seed=20261018)
    dev = torch.device('cuda', 0)
    d = {k: torch.from_numpy(w[k]).to(dev) for k in ('read_bases', 'read_off', 'tid', 'position', 'range1')}
    cap = int(L.indelgpu_seg_bound(n, n * 150))
    o = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)] + [torch.empty(n, dtype=torch.int64, device=dev), torch.empty(cap, dtype=torch.int32, device=dev)]
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    b = _lib.Batch(n, d['read_bases'].data_ptr(), d['read_off'].data_ptr(), d['tid'].data_ptr(), d['position'].data_ptr(), d['range1'].data_ptr())
    r = _lib.Result(o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), o[4].data_ptr(), cap, 0, None, None, None, 0)
    st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
    for it in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        rc = L.indelgpu_realign_batch_device(R._ctx, C.byref(b), 150, 700, C.byref(r), cnt.data_ptr(), C.c_void_p(st.cuda_stream))
        e1.record(st)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"g={g} it={it} rc={rc} host_call={1e3*(t1-t0):.2f} ms  events={e0.elapsed_time(e1):.2f} ms  wall={1e3*(t2-t0):.2f} ms")
    R.close()
