"""Timeline of the tcgen05 conv kernel's CTA 0 (tuning build with -DEXP_TRACE):
    python -c "from emphases_b200 import build; build.build(output='emphases_b200/exp_TRACE.so', extra_flags=['-DEXP_TRACE'])"
    EMPHASES_B200_LIB=$PWD/emphases_b200/exp_TRACE.so python tools/conv_trace.py
Events (clock64 of SM 0): 0/1 epilogue slot 0 waits for / sees mma_done, 2 its
operand stores are issued, 3 act_ready arrive; 4/5 MMA warp waits for / sees
act_ready[0], 6 slot 0's MMAs issued, 7 last slot's MMAs issued."""
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
eng = emphases.get_engine(dev)
views = eng.upload_plan(plan)
for _ in range(3):
    eng.forward_packed(audio, plan, model.packed_weights(), method='sum',
                       location='intermediate', precision=emphases.precision_code(), views=views)
torch.cuda.synchronize()
lib = _lib.load()
host = np.zeros((8, 512), dtype=np.int64)
lib.emph_conv_trace_read.argtypes = [ctypes.c_void_p]
print('rc', lib.emph_conv_trace_read(host.ctypes.data))
t0 = host[4, 0]
rel = host - t0
names = ['epi wait', 'epi done-seen', 'epi stored', 'epi arrive', 'mma wait0', 'mma ready0', 'mma issued0', 'mma issued3']
for i in range(7, 35):
    print(i, ' '.join(f'{names[e]}={rel[e, i]}' for e in range(8)))
d = np.diff(host[5, 7:200])
print('period of slot-0 layer steps (clk): median', np.median(d), 'mean', d.mean())
print('epilogue latency mma_done-seen -> arrive: median', np.median((host[3] - host[1])[7:200]))
print('mma issue->done-seen: median', np.median((host[1, 8:200] - host[6, 8:200])))
print('arrive -> mma ready seen: median', np.median((host[5, 8:200] - host[3, 7:199])))
