"""Time from_alignments_and_audio on a LIST of pageable per-utterance CPU tensors
(the reference's calling convention), generic fp32 and 16-bit-PCM-valued audio"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases
from emphases_b200 import scheduler
lengths, times = bench.corpus_layout(3000, 1234)
host, offsets = bench.make_audio(lengths, 99, pin=False)
audios = [host[o:o+n][None].clone() for o, n in zip(offsets, lengths)]
pcm_audios = [((a * 32768).round().clamp(-32768, 32767) / 32768.) for a in audios]
state = bench.random_state(); emphases.configure(PRECISION='bf16')
model = emphases.Model(); model.load_state_dict(state); model = model.cuda().eval()
for name, data in (('fp32 list', audios), ('pcm-exact fp32 list', pcm_audios)):
    for _ in range(2):
        emphases.from_alignments_and_audio(times, data, 16000, model=model, gpu=0)
    steps = []
    for _ in range(5):
        torch.cuda.synchronize(); t = time.perf_counter()
        emphases.from_alignments_and_audio(times, data, 16000, model=model, gpu=0)
        torch.cuda.synchronize(); steps.append((time.perf_counter() - t) * 1e3)
    print(name, ' '.join(f'{s:.1f}' for s in steps), 'ms')
t = time.perf_counter(); p = scheduler.pack_audio(audios); print('pack_audio', (time.perf_counter()-t)*1e3, 'ms')
