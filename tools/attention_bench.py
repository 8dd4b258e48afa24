"""Time the attention forms (fp32 CUDA cores; split-bf16 and fp16 tensor
cores) of csrc/attention.cu / attention_tc.cu over the frames of bench.py's corpus, and
the whole Transformer variant in each PRECISION.

    python tools/attention_bench.py [utterances]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import emphases_b200 as emphases  # noqa: E402
from emphases_b200 import _lib, engine, transformer  # noqa: E402


def main():
    utterances = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    device = torch.device('cuda', 0)
    torch.cuda.set_device(device)
    eng = emphases.get_engine(device)
    lengths, times = bench.corpus_layout(utterances, seed=1234)
    audio, _ = bench.make_audio(lengths, seed=99, pin=True)
    plan = engine.make_plan([(t, int(n)) for t, n in zip(times, lengths)], None, 'sum')
    frames = np.asarray(plan.n_rows, dtype=np.int64)
    starts, total = engine.packed_starts(frames.tolist())
    row_start = torch.from_numpy(starts.astype(np.int32)).to(device)
    n_rows = torch.from_numpy(frames.astype(np.int32)).to(device)
    row_seq = eng.row_index(row_start, n_rows, len(frames), total)
    block_seq, block_q0 = transformer.query_blocks(frames)
    d_seq, d_q0 = torch.from_numpy(block_seq).to(device), torch.from_numpy(block_q0).to(device)
    channels, heads = 80, transformer.HEADS
    q, k, v = (torch.randn(total, channels, device=device) for _ in range(3))
    out = torch.empty_like(q)
    scale = 1.0 / np.sqrt(channels // heads)
    pairs = float((frames ** 2).sum()) * heads
    flops = pairs * 4 * (channels // heads)
    report = {'rows': int(total), 'utterances': utterances, 'score_elements': pairs,
              'flops_per_call': flops}

    def run(mode):
        if mode is None:
            _lib.call(
                'emph_attention_rows', _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), channels, heads,
                _lib.ptr(row_start), _lib.ptr(n_rows), _lib.ptr(n_rows), _lib.ptr(row_seq),
                total, _lib.ptr(d_seq), _lib.ptr(d_q0), len(block_seq), scale, _lib.ptr(out),
                _lib.stream_ptr())
        else:
            _lib.call(
                'emph_attention_rows_tc', _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), channels, heads,
                _lib.ptr(row_start), _lib.ptr(n_rows), _lib.ptr(n_rows), _lib.ptr(row_seq),
                total, _lib.ptr(d_seq), _lib.ptr(d_q0), len(block_seq), scale, mode,
                _lib.ptr(workspace), workspace.numel(), _lib.ptr(out), _lib.stream_ptr())

    workspace = transformer.attention_workspace(total, channels, 1, device)
    reference = None
    for name, mode in (('fp32', None), ('bf16x3', 1), ('fp16', 0)):
        for _ in range(2):
            run(mode)
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(5):
            run(mode)
        end.record()
        torch.cuda.synchronize()
        ms = start.elapsed_time(end) / 5
        if reference is None:
            reference = out.clone()
        report[name] = {
            'ms': round(ms, 3), 'tflops': round(flops / ms * 1e-9, 1),
            'max_abs_vs_fp32': float((out - reference).abs().max())}
    print(json.dumps(report))

    # the whole variant (log-mel, input layer, 6 + 6 layers, pooling, head)
    views = eng.upload_plan(plan)
    device_audio = audio.to(device)
    passes = [('fp32', None), ('bf16x6', 'fp32'), ('bf16x6', None), ('bf16', None)]
    for precision, override in passes:
        os.environ.pop('EMPHASES_B200_ATTENTION', None)
        if override:
            os.environ['EMPHASES_B200_ATTENTION'] = override
        emphases.configure(ARCHITECTURE='transformer', PRECISION=precision)
        torch.manual_seed(0)
        model = emphases.Model().to(device).eval()
        weights = model.packed_weights()
        code = emphases.precision_code()
        timers = {}

        def step(timers=None):
            return eng.forward_packed(
                device_audio, plan, weights, method='sum', location='intermediate',
                precision=code, views=views, timers=timers)
        scores = step()['scores']
        if precision == 'fp32':
            exact = scores.clone()
        torch.cuda.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(3):
            step()
        end.record()
        torch.cuda.synchronize()
        step(timers)
        torch.cuda.synchronize()
        stages = {name: round(t[0].elapsed_time(t[1]), 3) for name, t in timers.items()}
        print(json.dumps({
            'precision': precision, 'attention': override or 'default', 'ms_per_pass': round(start.elapsed_time(end) / 3, 2),
            'max_abs_score_vs_fp32': float((scores - exact).abs().max()), 'stages': stages}))


if __name__ == '__main__':
    main()
