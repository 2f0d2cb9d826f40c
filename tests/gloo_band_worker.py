"""2-rank gloo worker for tests/test_host.py::test_two_rank_gloo_band_merge (CPU only)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fme_loader  # noqa: E402
import oracle_bindings as ob  # noqa: E402

fme = fme_loader.load()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
W, H = 192, 192
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=2, seed=11)
# rank 0 "owns" the newly reconstructed reference and broadcasts it (banded mode)
t = torch.from_numpy(np.stack(refs)) if rank == 0 else torch.zeros((2, H, W), dtype=torch.uint8)
dist.broadcast(t, src=0)
refs = [t[i].numpy() for i in range(2)]
recs = fme.pu_list.make_records(W, H, motions, seed=1)
frame = ob.CpuFrame(org, refs)
frame.oracle_fill_surface(recs)
lam = fme.pu_list.slice_lambda(22)
# bench.py's banded leg: pixel-balanced contiguous CTU runs (argv[1] == "balanced"), else whole CTU rows
balanced = len(sys.argv) > 1 and sys.argv[1] == "balanced"
mine = fme.pu_list.band_of_pus_balanced(recs, rank, world, W) if balanced else fme.pu_list.band_of_pus(recs, rank, world, H)
res = frame.oracle_run(np.ascontiguousarray(mine), 3, lam, 1, fme.nn_weights.load_blob(22))
gathered = [None] * world
dist.all_gather_object(gathered, res.tobytes())
if rank == 0:
    full = frame.oracle_run(recs, 3, lam, 1, fme.nn_weights.load_blob(22))
    merged = np.concatenate([np.frombuffer(b, fme.RESULT_DTYPE) for b in gathered])
    bands = [fme.pu_list.band_rows(b, world, H) for b in range(world)]
    if balanced:
        order = np.concatenate([np.nonzero(fme.pu_list.band_mask_balanced(recs, b, world, W))[0] for b in range(world)])
    else:
        order = np.concatenate([np.nonzero((recs["y"] // 64 >= lo) & (recs["y"] // 64 < hi))[0] for lo, hi in bands])
    assert np.array_equal(merged.view(np.uint8), full[order].view(np.uint8))
    print("BAND_MERGE_OK", len(merged))
dist.barrier()
dist.destroy_process_group()
