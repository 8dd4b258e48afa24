"""Per-step e2e timings, fp32 / int16 / fp32 again, inside a bench-like process"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases
from emphases_b200 import engine, scheduler
dev = torch.device('cuda', 0)
scheduler.bind_to_gpu_numa_node(0)
lengths, times = bench.corpus_layout(3000, 1234)
host, offsets = bench.make_audio(lengths, 99, pin=True)
packed = scheduler.PackedAudio(host, offsets, lengths)
state = bench.random_state(); emphases.configure(PRECISION='bf16')
model = emphases.Model(); model.load_state_dict(state); model = model.cuda().eval()
if len(sys.argv) > 1:       # mimic the kernel-only leg first
    eng = emphases.get_engine(dev); weights = model.packed_weights()
    plan = engine.make_plan([(t, int(n)) for t, n in zip(times, lengths)], None, 'sum')
    device_audio = host.to(dev); views = eng.upload_plan(plan)
    for _ in range(5):
        eng.forward_packed(device_audio, plan, weights, precision=emphases.precision_code(), views=views)
    torch.cuda.synchronize()
    print('allocated GB', torch.cuda.memory_allocated() / 1e9, 'reserved', torch.cuda.memory_reserved() / 1e9)
pcm = scheduler.PackedAudio((host * 32768.).round_().clamp_(-32768, 32767).to(torch.int16).pin_memory(), offsets, lengths)
def leg(name, audio, n=6):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter()
        emphases.from_alignments_and_audio(times, audio, 16000, model=model, gpu=0)
        torch.cuda.synchronize(); out.append((time.perf_counter() - t) * 1e3)
    print(name, ' '.join(f'{x:.1f}' for x in out), flush=True)
leg('fp32 ', packed); leg('int16', pcm); leg('fp32 ', packed); leg('int16', pcm)
print('reserved GB', torch.cuda.memory_reserved() / 1e9)
