"""Time K2 per PU shape class on the 1080p workload (device-resident records), to see where K2's time goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
W, H = 1920, 1080
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=4, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
eng = fme.Fme(W, H, num_ref_slots=4, max_pus=len(recs))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
eng.set_slice(fme.pu_list.slice_lambda(22)); eng.upload_org(org)
for s in range(4): eng.upload_ref(s, refs[s])
d_res = torch.zeros((len(recs), 16), dtype=torch.uint8, device="cuda")
shapes = sorted(set(zip(recs["w"].tolist(), recs["h"].tolist())), key=lambda s: -s[0] * s[1])
tot_px = float((recs["w"].astype(np.int64) * recs["h"]).sum())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def time_recs(r):
    d = torch.from_numpy(np.ascontiguousarray(r).view(np.uint8).reshape(len(r), -1)).cuda()
    for _ in range(2): eng.submit_device(d.data_ptr(), len(r), d_res.data_ptr(), fme.MODE_STD)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(5): eng.submit_device(d.data_ptr(), len(r), d_res.data_ptr(), fme.MODE_STD)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
print("all: %.3f ms" % time_recs(recs))
acc = 0.0
for (w, h) in shapes:
    r = recs[(recs["w"] == w) & (recs["h"] == h)]
    ms = time_recs(r); acc += ms
    px = float(len(r)) * w * h
    print("%2dx%-2d n=%7d px-share %5.1f%%  %.3f ms  %.2f ns/PU  %.1f ps/px-cand" % (w, h, len(r), 100 * px / tot_px, ms, 1e6 * ms / len(r), 1e9 * ms / (px * 17)))
print("sum of classes: %.3f ms" % acc)
