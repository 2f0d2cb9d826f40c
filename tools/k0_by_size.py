"""K0 time when only the PUs below an area threshold are left to the device (the larger ones arrive with their error
grid from the host): how many records / bytes the host sends extra, what K0 still costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
W, H = 1920, 1080
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=4, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2, err_on_gpu=True)
eng = fme.Fme(W, H, num_ref_slots=4, max_pus=len(recs))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
eng.set_slice(fme.pu_list.slice_lambda(22)); eng.upload_org(org)
for s in range(4): eng.upload_ref(s, refs[s])
area = recs["w"].astype(np.int32) * recs["h"].astype(np.int32)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for thr in (1 << 30, 4096, 2048, 1024, 512, 256, 128, 64):
    sub = np.ascontiguousarray(recs[area < thr])
    if len(sub) == 0:
        continue
    d = torch.from_numpy(sub.view(np.uint8).reshape(len(sub), -1)).cuda()
    for _ in range(3): eng.int_surface_device(d.data_ptr(), len(sub))
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(10): eng.int_surface_device(d.data_ptr(), len(sub))
    e1.record(st); torch.cuda.synchronize()
    big = len(recs) - len(sub)
    print("area < %-10d: %7d PUs on the device, K0 %.3f ms; %6d PUs (%.1f %% of the area) with host error grids = +%.2f MB"
          % (thr, len(sub), e0.elapsed_time(e1) / 10, big, 100.0 * area[area >= thr].sum() / area.sum(), big * 36 / 1e6))
