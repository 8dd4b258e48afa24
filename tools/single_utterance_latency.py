"""BASELINE config 1 as a latency number: emphases.from_alignment_and_audio on one
10 s utterance with a 25-word alignment (host audio in, scores on the host out),
median of 200 calls, plus a cProfile of one call."""
import os, sys, time, statistics, cProfile, pstats
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases

state = bench.random_state()
torch.manual_seed(0)
audio = (0.1 * torch.randn(1, 160000)).clamp(-1, 1)
generator = torch.Generator().manual_seed(1)
cuts = torch.sort(torch.rand(24, generator=generator) * 10).values.tolist()
edges = [0.] + cuts + [10.]
alignment = emphases.Alignment.from_times(list(zip(edges[:-1], edges[1:])))
path = '/tmp/single_utterance_checkpoint.pt'
torch.save({'model': state}, path)
for precision in ('bf16x6', 'bf16'):
    emphases.configure(PRECISION=precision)
    for _ in range(20):
        emphases.from_alignment_and_audio(alignment, audio, 16000, checkpoint=path, gpu=0).cpu()
    samples = []
    for _ in range(200):
        torch.cuda.synchronize()
        t = time.perf_counter()
        scores = emphases.from_alignment_and_audio(alignment, audio, 16000, checkpoint=path, gpu=0).cpu()
        samples.append((time.perf_counter() - t) * 1e3)
    print(f'{precision}: median {statistics.median(samples):.3f} ms, p10 {sorted(samples)[20]:.3f}, '
          f'p90 {sorted(samples)[180]:.3f}  ({10. / (statistics.median(samples) * 1e-3):.0f} x real time)')
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    emphases.from_alignment_and_audio(alignment, audio, 16000, checkpoint=path, gpu=0).cpu()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
