"""Device-path throughput on frames that are not the benchmark's: `noisy` (default) = the scene generator of tests/gpu_fuzz.py,
`kitti` = the KITTI extractor's conventions.  Checks a sample of the outputs against the oracle as well (a tool run, it loads the oracle)."""
import sys, importlib.util, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, torch
from _load_pkg import load_pkg, load_synth, load_oracle
spec = importlib.util.spec_from_file_location("gpu_fuzz", "tests/gpu_fuzz.py"); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
pkg, O = load_pkg(), load_oracle()
sp = O.sensor("HDL_64E")
kind = sys.argv[1] if len(sys.argv) > 1 else "noisy"
if kind == "kitti":      # KITTI-extractor convention: every valid point has intensity -1, every empty slot is an all-zero record at row = col = 0
    synth = load_synth()
    frames = [synth.make_frame("HDL_64E", 100 + s, kitti_quirk=True) for s in range(32)]
elif kind in ("slot_order", "column_order"):   # the benchmark's frames, not shuffled: in slot order (organised files) / column after column (MulRan order)
    synth = load_synth()
    frames = []
    for s_ in range(32):
        f = synth.make_frame("HDL_64E", 100 + s_)
        r, c = f["row"].astype(np.int64), f["col"].astype(np.int64)
        o = np.argsort(r * sp.horizon_scan + c if kind == "slot_order" else c * sp.n_scan + r, kind="stable")
        frames.append({k: v[o] for k, v in f.items()})
else:
    frames = [m.scene_frame(np.random.default_rng(100 + s), sp) for s in range(32)]
offs = np.zeros(33, np.int64); offs[1:] = np.cumsum([len(f["x"]) for f in frames])
distinct = {k: np.concatenate([f[k] for f in frames]) for k in bench.FIELDS}; distinct["offsets"] = offs
F = 2220
g = pkg.BevGen("HDL_64E", device=0, max_frames_per_batch=1110)
batch = bench.tile_batch(distinct, F); n_total = int(batch["offsets"][-1]); dev = torch.device("cuda", 0)
din = {k: torch.from_numpy(batch[k]).to(dev) for k in bench.FIELDS}
dout = dict(label=torch.empty((F, g.S), dtype=torch.int16, device=dev), winner=torch.zeros(pkg.winner_words(n_total, F), dtype=torch.int32, device=dev),
            single=torch.empty((F, 224 * 224), dtype=torch.uint8, device=dev), multi=torch.empty((F, 24 * 224 * 224), dtype=torch.uint8, device=dev))
pin, pout = {k: v.data_ptr() for k, v in din.items()}, {k: v.data_ptr() for k, v in dout.items()}
stream = torch.cuda.ExternalStream(g.compute_stream(), device=dev)
for _ in range(3): g.process_device(F, batch["offsets"], pin, pout)
g.sync(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(4): g.process_device(F, batch["offsets"], pin, pout)
e1.record(stream); g.sync(); torch.cuda.synchronize()
g.set_profiling(True); g.process_device(F, batch["offsets"], pin, pout); g.sync(); st = g.stage_ms(); g.set_profiling(False)
ref = O.frames(sp, offs, *[distinct[k] for k in bench.FIELDS], n_threads=16)
ok = all(np.array_equal(dout["label"][i].cpu().numpy(), ref["label"][i % 32]) and np.array_equal(dout["multi"][i].cpu().numpy().reshape(24, 224, 224), ref["multi"][i % 32]) for i in list(range(40)) + [F - 1])
print({"kitti": "KITTI-convention frames (all empties on slot 0)", "slot_order": "benchmark frames with their points in slot order", "column_order": "benchmark frames with their points column after column"}.get(kind, "noisy scene frames (1.2 k - 12 k segments, median 6.3 k)") + ": %.0f frames/s, %s, parity %s" % (F * 4 / (e0.elapsed_time(e1) * 1e-3), {k: round(v[0] / F * 1e3, 3) for k, v in st.items() if v[1] > 0}, ok))
