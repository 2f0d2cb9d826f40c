import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, fme_loader
fme = fme_loader.load()
for (W, H, M) in ((416, 240, 80), (1920, 1080, 80), (136, 72, 16)):
    rng = np.random.default_rng(W)
    pic = rng.integers(0, 256, (H, W)).astype(np.uint8)
    pic[: H // 3] = np.where(rng.integers(0, 2, (H // 3, W)) > 0, 255, 0)
    planes = {}
    for path in (2, 3):
        eng = fme.Fme(W, H, num_ref_slots=1, max_pus=16, margin=M, k1_path=path)
        eng.upload_ref(0, pic)
        planes[path] = [eng.download_plane(0, k // 4, k % 4) for k in range(16)]
        eng.close()
    for k in range(16):
        bad = np.argwhere(planes[2][k] != planes[3][k])
        if len(bad):
            print(W, H, "plane", k, "mismatches", len(bad), "first", bad[:4].tolist(), planes[2][k][tuple(bad[0])], planes[3][k][tuple(bad[0])],
                  "rows", sorted(set(bad[:, 0]))[:8], "cols", sorted(set(bad[:, 1]))[:8])
            break
    else:
        print(W, H, "all 16 planes identical")
