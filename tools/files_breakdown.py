"""Where the time of emphases_b200.from_files_to_files goes: wall time and a
cProfile of one call on bench.py's on-disk corpus (files x copies)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases

count = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 8
state = bench.random_state()
emphases.configure(PRECISION='bf16')
root = bench.corpus_root() + '_breakdown'
text, audio, prefixes, checkpoint, seconds, words, samples = bench.build_corpus(
    emphases, root, count, copies, state)
try:
    for rep in range(3):
        t0 = time.perf_counter()
        emphases.from_files_to_files(text, audio, prefixes, checkpoint=checkpoint, gpu=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'from_files_to_files: {dt * 1e3:.0f} ms = {seconds / dt:.0f} audio-s/s, '
              f'{dt / len(text) * 1e3:.3f} ms/file')
    profile = cProfile.Profile()
    profile.enable()
    emphases.from_files_to_files(text, audio, prefixes, checkpoint=checkpoint, gpu=0)
    torch.cuda.synchronize()
    profile.disable()
    pstats.Stats(profile).sort_stats('cumulative').print_stats(40)
finally:
    import shutil
    shutil.rmtree(root, ignore_errors=True)
