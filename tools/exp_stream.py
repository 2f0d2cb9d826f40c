import sys, numpy as np, os
sys.path.insert(0,'/root/repo')
import torch
variant = sys.argv[1]
dev=torch.device("cuda",0); torch.cuda.set_device(0)
import fme_loader; fme=fme_loader.load()
W,H=(1920,1080) if 'big' in variant else (128,96)
eng=fme.Fme(W,H,num_ref_slots=4,max_pus=16, device=0)
pic=np.random.default_rng(0).integers(0,256,(H,W)).astype(np.uint8)
if 'ownfirst' in variant:
    eng.upload_ref(0,pic); eng.synchronize()
st=torch.cuda.Stream(device=dev)
if 'setstream' in variant: torch.cuda.set_stream(st)
eng.set_stream(st.cuda_stream)
if 'nn' in variant:
    eng.set_nn_weights(fme.nn_weights.load_blob(22)); eng.set_slice(9.3)
d=torch.from_numpy(pic).to(dev)
try:
    if 'org' in variant: eng.upload_org_device_u8(d.data_ptr(), W)
    eng.upload_ref_device_u8(0, d.data_ptr(), W); eng.synchronize(); print(variant, "OK")
except Exception as e: print(variant, "FAIL", e)
