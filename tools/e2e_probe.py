"""Where does the host-buffer step go?  The e2e loop of bench.py with parts of it left out (1080p, one GPU)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
W, H, NREF, LAG = 1920, 1080, 4, 2
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=NREF, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
n = len(recs)
eng = fme.Fme(W, H, num_ref_slots=NREF, max_pus=n)
eng.set_nn_weights(fme.nn_weights.load_blob(22))
lib, hnd = eng.lib, eng.h
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_org = pin(org.astype(np.int16)); h_ref = [pin(r.astype(np.int16)) for r in refs]
h_heads = pin(fme.pu_list.heads_of(recs).view(np.uint8).reshape(n, -1))
h_pus = pin(recs.view(np.uint8).reshape(n, -1))
gr = fme.pu_list.grids_of(recs, 128)
h_gr = pin(gr.view(np.uint8).reshape(-1, 40))
h_out = [torch.zeros((n, 16), dtype=torch.uint8).pin_memory() for _ in range(LAG + 1)]
eng.set_slice(fme.pu_list.slice_lambda(22))
eng.upload_org(org)
for s in range(NREF): eng.upload_ref(s, refs[s])

def loop(count, pics=True, mode=fme.MODE_BOTH | fme.MODE_RESULT8, kind="grids", slice_=True):
    for i in range(count):
        if slice_: eng.set_slice(fme.pu_list.slice_lambda(22) * (1 + 0.01 * (i % 4)))
        if pics:
            eng._check(lib.fme_upload_ref(hnd, i % NREF, ctypes.c_void_p(h_ref[i % NREF].data_ptr()), W))
            eng._check(lib.fme_upload_org(hnd, ctypes.c_void_p(h_org.data_ptr()), W))
        o = h_out[i % (LAG + 1)].data_ptr()
        if kind == "grids": eng.submit_heads_grids_async(h_heads.data_ptr(), n, h_gr.data_ptr(), len(gr), o, mode)
        elif kind == "heads": eng.submit_heads_async(h_heads.data_ptr(), n, o, mode)
        else: eng.submit_async(h_pus.data_ptr(), n, o, mode)
        if i >= LAG: eng.wait_oldest()
    for _ in range(min(LAG, count)): eng.wait_oldest()

def timed(name, **kw):
    loop(10, **kw); eng.synchronize()
    t0 = time.perf_counter(); loop(100, **kw); eng.synchronize()
    print("%-58s %.4f ms per frame" % (name, (time.perf_counter() - t0) * 10))

timed("grids, pictures, BOTH|RESULT8")
timed("grids, no picture uploads", pics=False)
timed("grids, pictures, STD only", mode=fme.MODE_STD | fme.MODE_RESULT8)
timed("grids, pictures, NN only", mode=fme.MODE_NN | fme.MODE_RESULT8)
timed("heads, pictures")
timed("heads, no picture uploads", kind="heads", pics=False)
timed("full records, no picture uploads, RESULT8", kind="full", pics=False)
timed("full records, pictures, RESULT8", kind="full")
# host-side cost of the calls alone: a batch of 16 PUs
nn = n
n = 16
timed("16-PU batches (host call overhead + fixed launch costs)")
