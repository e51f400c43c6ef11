import sys, torch
sys.path.insert(0, ".")
import msc_thesis_b200 as h3
from bench import make_workload
dev = torch.device("cuda", 0)
n, T = 1 << 20, 200
chunk = n // 8
def make(nc):
    lp = make_workload(h3, torch, nc, "f64", dev, seed=1, store=0, ticks=T)
    lp.switch_step = 600 + T // 2
    return lp
for compact in (False, True):
    pipe = h3.HostBatchPipeline(make, chunk, n_slots=2, compact=compact)
    for _ in range(3):
        pipe.slots[0].run(T)
    pipe.k0, pipe.t0 = pipe.slots[0].k, pipe.slots[0].env.t
    hc = pipe.host_chunks_from(pipe.slots[0], 8)
    hr = pipe.result_buffers(8)
    pipe.run(hc, hr, T); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pipe.run(hc, hr, T)
    e1.record(); torch.cuda.synchronize()
    print("compact", compact, "ms/step", e0.elapsed_time(e1) / 3, "bytes", sum(t.numel() * t.element_size() for ch in hc for t in ch) / 1e9)
    # H2D only
    e0.record()
    for _ in range(3):
        for c, src in enumerate(hc):
            with torch.cuda.stream(pipe.streams[c % 2]):
                for i, (h, d) in enumerate(zip(src, h3.host_batch.loop_state_tensors(pipe.slots[c % 2]))):
                    if h.shape == d.shape:
                        d.copy_(h, non_blocking=True)
    for s in pipe.streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    print("   H2D of the same-shape tensors only: ms/step", e0.elapsed_time(e1) / 3)
    del pipe, hc, hr
    torch.cuda.empty_cache()
