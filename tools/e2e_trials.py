"""Developer aid: repeated e2e trials (HostStreamer, all-host inputs) to see how stable the PCIe-bound step time is."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, einx
bench = importlib.import_module("bench")
synth = importlib.import_module("ei-nexus_official_b200.synth")
dev = torch.device("cuda", 0)
name, B = "c2_ec_superpoint", 64
c = synth.CONFIGS[name]
cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"], descriptor_scale=c["scale"], precision="fp16x3")
pipe = einx.ExtractMatchPipeline(cfg)
hs = [einx.HostBatch(*bench.make_batch(synth, name, B, s * B), chunks=2) for s in range(3)]
K = c["top_k"]
out_host = {"matches0": torch.empty((B, K), dtype=torch.int64).pin_memory(), "num_matches": torch.empty((B,), dtype=torch.int32).pin_memory()}
st = einx.HostStreamer(pipe, dev)
for trial in range(12):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(20): st.run(hs[i % 3], out_host)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f"trial {trial}: {e0.elapsed_time(e1) / 20:.2f} ms/step (host issue {1e3 * (t1 - t0) / 20:.2f} ms/step)", flush=True)
    if trial == 5: time.sleep(2.0)
