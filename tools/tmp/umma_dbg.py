import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, fme_loader
fme = fme_loader.load()
W, H, M = 128, 64, 16
def run(pic, path):
    eng = fme.Fme(W, H, num_ref_slots=1, max_pus=16, margin=M, k1_path=path)
    eng.upload_ref(0, pic)
    out = [eng.download_plane(0, k // 4, k % 4) for k in range(16)]
    eng.close()
    return out
tests = {"const100": np.full((H, W), 100, np.uint8),
         "xramp": np.tile(np.arange(W, dtype=np.uint8), (H, 1)),
         "yramp": np.tile((np.arange(H, dtype=np.uint8) * 2)[:, None], (1, W))}
for name, pic in tests.items():
    a, b = run(pic, 2), run(pic, 3)
    for k in (0, 1, 4, 10):
        print(name, "plane", k, "ref row40:", a[k][40, 30:42].tolist(), "umma:", b[k][40, 30:42].tolist())
    print(name, "plane 0 col 60 ref:", a[0][20:52:4, 60].tolist(), "umma:", b[0][20:52:4, 60].tolist())
