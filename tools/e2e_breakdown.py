"""Where does the end-to-end step time go?  (run under gpurun)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases
from emphases_b200 import engine, scheduler

lengths, times = bench.corpus_layout(3000, 1234)
host, offsets = bench.make_audio(lengths, 99, pin=True)
packed = scheduler.PackedAudio(host, offsets, lengths)
state = bench.random_state(); emphases.configure(PRECISION='bf16')
model = emphases.Model(); model.load_state_dict(state); model = model.cuda().eval()
dev = torch.device('cuda', 0)
print('cpus', os.cpu_count(), 'threads', torch.get_num_threads())
for _ in range(3):
    t = time.perf_counter(); d = host.to(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f'H2D 2.1GB: {dt*1e3:.1f} ms = {host.numel()*4/dt/1e9:.1f} GB/s')
u = [(t_, int(n)) for t_, n in zip(times, lengths)]
for _ in range(3):
    t = time.perf_counter(); plan = engine.make_plan(u, None, 'sum'); print(f'make_plan: {(time.perf_counter()-t)*1e3:.1f} ms')
for rows in (1 << 19, 1 << 22):
    emphases.configure(MAX_ROWS_PER_LAUNCH=rows)
    for _ in range(2):
        emphases.from_alignments_and_audio(times, packed, 16000, model=model, gpu=0)
    steps = []
    for _ in range(8):
        torch.cuda.synchronize(); t = time.perf_counter()
        emphases.from_alignments_and_audio(times, packed, 16000, model=model, gpu=0)
        torch.cuda.synchronize(); steps.append((time.perf_counter() - t) * 1e3)
    print(f'rows/launch {rows}: e2e ms', ' '.join(f'{s:.1f}' for s in steps))
import cProfile, pstats
emphases.configure(MAX_ROWS_PER_LAUNCH=1 << 19)
pr = cProfile.Profile(); pr.enable()
emphases.from_alignments_and_audio(times, packed, 16000, model=model, gpu=0)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
