"""Kernel-stream timeline of the host-buffer (e2e) loop: where the stream is busy and where it waits (debug build:
tools/build_variant.sh stamps "-DFME_STAMPS";  FME_B200_LIB=variants/libfme_stamps.so python tools/e2e_timeline.py [kind])."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, fme_loader
fme = fme_loader.load()
kind = sys.argv[1] if len(sys.argv) > 1 else "compact"
W, H, NREF, LAG = 1920, 1080, 4, 2
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=NREF, seed=2022)
recs = fme.pu_list.make_records(W, H, motions, seed=2)
n = len(recs)
eng = fme.Fme(W, H, num_ref_slots=NREF, max_pus=n)
eng.set_nn_weights(fme.nn_weights.load_blob(22))
lib, hnd = eng.lib, eng.h
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_org = pin(org.astype(np.int16)); h_ref = [pin(r.astype(np.int16)) for r in refs]
h_pus = pin(recs.view(np.uint8).reshape(n, -1))
comp, big = fme.pu_list.compact_of(recs)
h_cmp = pin(comp.view(np.uint8).reshape(n, -1))
h_heads = pin(fme.pu_list.heads_of(recs).view(np.uint8).reshape(n, -1))
h_out = [torch.zeros((n, 8), dtype=torch.uint8).pin_memory() for _ in range(LAG + 1)]
eng.set_slice(fme.pu_list.slice_lambda(22)); eng.upload_org(org)
for s in range(NREF): eng.upload_ref(s, refs[s])
mode = fme.MODE_BOTH | fme.MODE_RESULT8

def loop(count):
    for i in range(count):
        eng.set_slice(fme.pu_list.slice_lambda(22) * (1 + 0.01 * (i % 4)))
        eng._check(lib.fme_upload_ref(hnd, i % NREF, ctypes.c_void_p(h_ref[i % NREF].data_ptr()), W))
        eng._check(lib.fme_upload_org(hnd, ctypes.c_void_p(h_org.data_ptr()), W))
        o = h_out[i % (LAG + 1)].data_ptr()
        if kind == "compact": eng.submit_compact_async(h_cmp.data_ptr(), n, 0, 0, o, mode)
        elif kind == "heads": eng.submit_heads_async(h_heads.data_ptr(), n, o, mode)
        else: eng.submit_async(h_pus.data_ptr(), n, o, mode)
        if i >= LAG: eng.wait_oldest()
    for _ in range(min(LAG, count)): eng.wait_oldest()

buf = (ctypes.c_ulonglong * 65536)()
loop(10); eng.synchronize(); lib.fme_debug_stamps(buf, 65536)
t0 = time.perf_counter(); loop(60); eng.synchronize(); wall = (time.perf_counter() - t0) / 60 * 1e3
m = lib.fme_debug_stamps(buf, 65536)
st = [(buf[i] >> 4, buf[i] & 15) for i in range(m)]
names = {(1, 2): "K1", (2, 3): "gap K1 end -> submit begin (org convert wait, records wait)", (3, 4): "expand / grids / K0",
         (4, 5): "K2 (binning + refine)", (5, 6): "K3", (6, 7): "pack results", (7, 1): "gap frame end -> next K1 begin"}
acc = {}
for (ta, a), (tb, b) in zip(st[:-1], st[1:]):
    acc.setdefault((a, b), []).append((tb - ta) * 1e-6)
print("%s: %.4f ms per frame wall" % (kind, wall))
tot = 0.0
for k in [(1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 1)]:
    v = acc.get(k, [0.0]); tot += float(np.mean(v))
    print("  %-62s %.4f ms (median %.4f, n=%d)" % (names[k], float(np.mean(v)), float(np.median(v)), len(v)))
print("  sum %.4f ms" % tot)
