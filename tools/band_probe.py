"""Banded 2160p leg of bench.py alone, for several values of the per-PU term of the band cost model
(torchrun --nproc-per-node N tools/band_probe.py 22 40 60): per-rank device times show which band is the slow one."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench, fme_loader
fme = fme_loader.load()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
for v in [float(a) for a in sys.argv[1:]] or [fme.pu_list.PER_PU_WORK]:
    fme.pu_list.PER_PU_WORK = v
    r = bench.banded_leg(fme, torch, dist, dev, stream, rank, world, local, 12, 3)
    if rank == 0:
        print("per-PU term %5.1f: %.4f ms per frame, efficiency %.3f, per rank %s, compute alone %s, band PUs %s"
              % (v, r["ms_per_step"], r.get("efficiency_vs_1gpu_same_run", 1.0), r.get("ms_per_rank"),
                 r.get("compute_ms_per_rank_without_broadcast"), r.get("band_pus")), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
