"""Time the bi-predictive refinement pass (k2_refine<true>) on the 1080p PU list with every record flagged FME_PU_BI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
W, H = 1920, 1080
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=4, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
rng = np.random.default_rng(5)
bi = recs.copy()
bi["flags"] = fme.PU_BI
bi["err"] = 0
bi["err"][:, 0] = rng.integers(0, 4, len(bi))
omx, omy = rng.integers(-40, 41, len(bi)), rng.integers(-40, 41, len(bi))
bi["err"][:, 1] = (omx & 0xffff) | ((omy & 0xffff) << 16)
eng = fme.Fme(W, H, num_ref_slots=4, max_pus=len(recs), bi_pred=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
eng.set_slice(fme.pu_list.slice_lambda(22)); eng.upload_org(org)
for s in range(4): eng.upload_ref(s, refs[s])
d_res = torch.zeros((len(recs), 16), dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, r in (("uni", recs), ("bi", bi)):
    d = torch.from_numpy(np.ascontiguousarray(r).view(np.uint8).reshape(len(r), -1)).cuda()
    for _ in range(2): eng.submit_device(d.data_ptr(), len(r), d_res.data_ptr(), fme.MODE_STD)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(5): eng.submit_device(d.data_ptr(), len(r), d_res.data_ptr(), fme.MODE_STD)
    e1.record(st); torch.cuda.synchronize()
    print("%s records: %.3f ms per %d PUs (both K2 passes launched)" % (name, e0.elapsed_time(e1) / 5, len(r)))
