"""from_files_to_files on 3000 files with gpu=[0] vs gpu=[0, 1, ...] (in-process sharding)"""
import os, sys, time, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases
from pathlib import Path

count = 3000
lengths, times = bench.corpus_layout(count, 77)
root = Path(tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None))
generator = torch.Generator().manual_seed(0)
text_files, audio_files, prefixes = [], [], []
for i, (n, t) in enumerate(zip(lengths, times)):
    audio = (0.1 * torch.randn(1, int(n), generator=generator)).clamp(-1, 1)
    emphases.load.save_wav(root / f'u{i}.wav', audio)
    emphases.Alignment.from_times([tuple(x) for x in t.tolist()]).save(root / f'u{i}.TextGrid')
    text_files.append(root / f'u{i}.TextGrid'); audio_files.append(root / f'u{i}.wav')
    prefixes.append(root / 'out' / f'u{i}')
(root / 'out').mkdir()
state = bench.random_state(); emphases.configure(PRECISION='bf16')
ckpt = root / 'ckpt.pt'; torch.save({'model': state}, ckpt)
for gpus in ([0], list(range(torch.cuda.device_count()))):
    for rep in range(4):
        t0 = time.perf_counter()
        emphases.from_files_to_files(text_files, audio_files, prefixes, checkpoint=ckpt, gpu=gpus if len(gpus) > 1 else gpus[0])
        dt = time.perf_counter() - t0
    print(f'gpu={gpus}: {dt*1e3:.0f} ms = {lengths.sum()/16000/dt:.0f} audio-s/s')
