"""Time K1 alone (16 back-to-back launches over 4 slots) -- used for what-if builds of k1_interp.cu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
arg = sys.argv[1] if len(sys.argv) > 1 else "1920x1080"
W, H = (3840, 2160) if arg == "4k" else tuple(int(v) for v in arg.split("x"))
eng = fme.Fme(W, H, num_ref_slots=4, max_pus=16)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); eng.set_stream(st.cuda_stream)
pic = np.random.default_rng(0).integers(0, 256, (H, W)).astype(np.uint8)
d = torch.from_numpy(pic).cuda()
for s in range(4): eng.upload_ref_device_u8(s, d.data_ptr(), W)   # aligned device picture: K1 reads it in place
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e0.record(st)
    for r in range(32): eng.upload_ref_device_u8(r % 4, d.data_ptr(), W)
    e1.record(st); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 32 * 1e3
print("K1 %dx%d: %.2f us per launch, %.0f GB/s algorithmic" % (W, H, us, (W + 160) * (H + 160) * 16 / us / 1e3))
