"""K2 + K3 over one 1080p PU list: mode BOTH (two launches, or K3 inside K2 with FME_K3_FUSE=1) against STD and NN alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
W, H = 1920, 1080
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=4, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
eng = fme.Fme(W, H, num_ref_slots=4, max_pus=len(recs))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
eng.set_nn_weights(fme.nn_weights.load_blob(22))
eng.set_slice(fme.pu_list.slice_lambda(22)); eng.upload_org(org)
for s in range(4): eng.upload_ref(s, refs[s])
d = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).reshape(len(recs), -1)).cuda()
d_res = torch.zeros((len(recs), 16), dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode, name in ((fme.MODE_BOTH, "BOTH"), (fme.MODE_STD, "STD"), (fme.MODE_NN, "NN")):
    for _ in range(2): eng.submit_device(d.data_ptr(), len(recs), d_res.data_ptr(), mode)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(5): eng.submit_device(d.data_ptr(), len(recs), d_res.data_ptr(), mode)
    e1.record(st); torch.cuda.synchronize()
    print("%s: %.4f ms" % (name, e0.elapsed_time(e1) / 5))
