"""Time the fused LogSV slice kernel alone (CUDA events) for the library named by $B200SV_LIB -- kernel tuning helper.
usage: B200SV_LIB=... python tools/time_slice.py [paths] [nsteps] [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import byref, c_void_p
from stochvolmodels_b200 import _capi as C, engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 136
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pc = engine.logsv_params_c(0.8376, 1.0413, 3.1844, 3.058, 0.1514, 1.8458)
dt = torch.float32 if flags & 1 else torch.float64
st = torch.empty((3, n), dtype=dt, device="cuda")
mom = torch.zeros(2, dtype=torch.float64, device="cuda")
stream = c_void_p(torch.cuda.current_stream().cuda_stream)
def run(seed):
    C.call("b200sv_dev_logsv_slice", c_void_p(st[0].data_ptr()), c_void_p(st[1].data_ptr()), c_void_p(st[2].data_ptr()), n, 0, 1,
           byref(pc), 1.0, 1, S, 0.25 / S, 0, 1.0, seed, flags, c_void_p(mom.data_ptr()), None, stream)
for w in range(3):
    run(w)
torch.cuda.synchronize()
ts = []
for k in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(100 + k); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
med = ts[len(ts) // 2]
print(f"{os.path.basename(os.environ.get('B200SV_LIB', 'default')):40s} flags={flags} n={n} S={S} median {med:8.3f} ms  {n * S / med / 1e6:8.2f} Gpath-steps/s  mean_exp_x={float(mom[0] / mom[1]):.6f}")
