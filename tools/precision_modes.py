"""Score error of every conv precision mode against the reference's fp32
forward (golden C1 example, trained checkpoint, sum pooling) and conv-stack
error against the oracle; run under gpurun."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from golden_util import state_from_golden
from emphases_b200 import _lib, engine

data = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'c1.npz'))
state = state_from_golden(data)
eng = engine.Engine('cuda:0')
weights = engine.pack_weights(state, torch.device('cuda:0'), layers=6, activation='ReLU', dropout=None, has_decoder=True)
times = np.asarray(data['times'])
audio = torch.from_numpy(data['audio'])[0].cuda()
for name, code in (('fp32 (FFMA)', _lib.PREC_FP32), ('bf16', _lib.PREC_BF16_TC),
                   ('bf16x3', _lib.PREC_BF16X3_TC), ('bf16x6', _lib.PREC_BF16X6_TC)):
    worst = 0.
    for batch_size, tag in ((None, 'full'), (300, 'bs300'), (100, 'bs100')):
        plan = engine.make_plan([(times, 160000)], batch_size)
        result = eng.forward_packed(audio, plan, weights, precision=code)
        scores = torch.cat([result['scores'][s:s + n] for s, n in zip(plan.word_row_start, plan.n_words)]).cpu().numpy()
        worst = max(worst, float(np.abs(scores - data[f'{tag}.scores'][0]).max()))
    print(f'{name:12s} max-abs score error vs the reference fp32 forward: {worst:.3e}')
