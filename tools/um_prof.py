"""Phase timers of k2_refine_umma (debug build: tools/build_variant.sh umprof "-DFME_UM_PROF").
  FME_B200_LIB=variants/libfme_umprof.so FME_K2_PATH=4 python tools/um_prof.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
W, H = 1920, 1080
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=4, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
eng = fme.Fme(W, H, num_ref_slots=4, max_pus=len(recs))
eng.set_slice(fme.pu_list.slice_lambda(22)); eng.upload_org(org)
for s in range(4): eng.upload_ref(s, refs[s])
lib = fme.load_library()
d = torch.from_numpy(np.ascontiguousarray(recs).view(np.uint8).reshape(len(recs), -1)).cuda()
d_res = torch.zeros((len(recs), 16), dtype=torch.uint8, device="cuda")
out = (ctypes.c_ulonglong * 16)()
for _ in range(2): eng.submit_device(d.data_ptr(), len(recs), d_res.data_ptr(), fme.MODE_STD)
lib.fme_debug_um_prof(out, 16)
eng.submit_device(d.data_ptr(), len(recs), d_res.data_ptr(), fme.MODE_STD)
n = lib.fme_debug_um_prof(out, 16)
names = ["setup", "stage_issue", "stage_wait", "done_wait", "collect", "book", "repack", "submit", "sched", "total"]
tot = float(out[9])
for k, nm in enumerate(names): print("%-12s %14d clk  %5.1f%%" % (nm, out[k], 100.0 * out[k] / tot))
