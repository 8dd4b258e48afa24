"""Timeline of the transposed conv kernel's CTA 0 (-DEXP_TRACE build, EMPHASES_B200_TC=transposed)"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import emphases_b200 as emphases
from emphases_b200 import engine, _lib
lengths, times = bench.corpus_layout(3000, 1234)
state = bench.random_state(); emphases.configure(PRECISION='bf16')
model = emphases.Model(); model.load_state_dict(state); model = model.cuda().eval()
dev = torch.device('cuda', 0)
audio, offsets = bench.make_audio(lengths, 99, device=dev)
plan = engine.make_plan([(t, int(n)) for t, n in zip(times, lengths)], None, 'sum')
eng = emphases.get_engine(dev); views = eng.upload_plan(plan)
for _ in range(2):
    eng.forward_packed(audio, plan, model.packed_weights(), method='sum', location='intermediate',
                       precision=emphases.precision_code(), views=views)
torch.cuda.synchronize()
lib = _lib.load()
host = np.zeros((12, 256), dtype=np.int64)
lib.emph_conv_tct_trace_read.argtypes = [ctypes.c_void_p]
print('rc', lib.emph_conv_tct_trace_read(host.ctypes.data))
t0 = host[0, 14]
names = ['mma:wait_w', 'mma:w_ready', 'mma:act0', 'mma:act1', 'mma:issued', 'epi0:wait', 'epi0:done', 'epi0:end', 'wgt:free', 'wgt:ready', 'epi0:ld0', 'epi0:ld1']
for i in range(14, 30):
    print(i, ' '.join(f'{names[e]}={host[e, i] - t0}' for e in range(12)))
