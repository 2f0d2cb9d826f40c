"""One rank's share of the banded 2160p step, emulated on ONE GPU: band b of N (pu_list.band_mask_balanced), K1 on the rows
its records reference, binning + K2, K3 -- CUDA events between the calls, to see which part does not shrink with N.
  python tools/band_one.py [N=8] [bands...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bands = [int(a) for a in sys.argv[2:]] or [0, N // 2, N - 1]
W, H, NREF = 3840, 2160, 4
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=NREF, seed=4)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
eng = fme.Fme(W, H, num_ref_slots=NREF, max_pus=len(recs))
eng.set_stream(st.cuda_stream)
eng.set_nn_weights(fme.nn_weights.load_blob(22)); eng.set_slice(fme.pu_list.slice_lambda(22))
d_org = torch.from_numpy(org).to(dev); d_refs = [torch.from_numpy(r).to(dev) for r in refs]
eng.upload_org_device_u8(d_org.data_ptr(), W)
for s in range(NREF): eng.upload_ref_device_u8(s, d_refs[s].data_ptr(), W)
d_full = torch.from_numpy(recs.view(np.uint8).reshape(len(recs), -1).copy()).to(dev)
eng.int_surface_device(d_full.data_ptr(), len(recs))
full = d_full.cpu().numpy().view(fme.PU_DTYPE).reshape(-1).copy(); full["flags"] = 0
d_res = torch.zeros((len(recs), 16), dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

def run(name, d, n, rows, orows=None):
    acc = np.zeros(4)
    for it in range(3 + 10):
        slot = it % NREF
        ev[0].record(st)
        if rows is None: eng.upload_ref_device_u8(slot, d_refs[slot].data_ptr(), W)
        else: eng.upload_ref_device_u8_rows(slot, d_refs[slot].data_ptr(), W, rows[0], rows[1])
        ev[1].record(st)
        if orows is None: eng.upload_org_device_u8(d_org.data_ptr(), W)
        else: eng.upload_org_device_u8_rows(d_org.data_ptr(), W, orows[0], orows[1])
        ev[2].record(st)
        eng.submit_device(d.data_ptr(), n, d_res.data_ptr(), fme.MODE_STD)
        ev[3].record(st)
        eng.submit_device(d.data_ptr(), n, d_res.data_ptr(), fme.MODE_NN)
        ev[4].record(st)
        torch.cuda.synchronize()
        if it >= 3: acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    acc /= 10
    # the whole step as bench issues it (one submit, mode BOTH)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for it in range(12):
        slot = it % NREF
        if rows is None: eng.upload_ref_device_u8(slot, d_refs[slot].data_ptr(), W)
        else: eng.upload_ref_device_u8_rows(slot, d_refs[slot].data_ptr(), W, rows[0], rows[1])
        if orows is None: eng.upload_org_device_u8(d_org.data_ptr(), W)
        else: eng.upload_org_device_u8_rows(d_org.data_ptr(), W, orows[0], orows[1])
        eng.submit_device(d.data_ptr(), n, d_res.data_ptr(), fme.MODE_BOTH)
    e1.record(st); torch.cuda.synchronize()
    print("%-14s %8d PUs  K1 %.4f  org %.4f  K2 pass %.4f  K3 pass %.4f  | step %.4f ms" % (name, n, acc[0], acc[1], acc[2], acc[3], e0.elapsed_time(e1) / 12), flush=True)

run("full frame", d_full, len(recs), None)
for b in bands:
    mine = np.ascontiguousarray(full[fme.pu_list.band_mask_balanced(full, b, N, W)])
    d = torch.from_numpy(mine.view(np.uint8).reshape(len(mine), -1).copy()).to(dev)
    run("band %d of %d" % (b, N), d, len(mine), fme.pu_list.referenced_rows(mine), fme.pu_list.source_rows(mine))
