"""HBM roofline of the strict fixed-random stepper (calibration inner loop): the kernel READS 16 B of normals per
path-step from HBM (plus 48 B of state per path per slice).  usage: python tools/bench_fixed_randoms.py [nb_path] [nb_steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import byref, c_void_p
from stochvolmodels_b200 import _capi as C, engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
pc = engine.logsv_params_c(0.8376, 1.0413, 3.1844, 3.058, 0.1514, 1.8458)
g = torch.Generator(device="cuda").manual_seed(1)
W0 = torch.randn((S, n), dtype=torch.float64, device="cuda", generator=g)
W1 = torch.randn((S, n), dtype=torch.float64, device="cuda", generator=g)
st = torch.zeros((3, n), dtype=torch.float64, device="cuda")
stream = c_void_p(torch.cuda.current_stream().cuda_stream)
FAST = int(sys.argv[3]) if len(sys.argv) > 3 else 1
def run():
    st[0].zero_(); st[1].fill_(0.8376); st[2].zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    C.call("b200sv_dev_logsv_step_fixed", c_void_p(st[0].data_ptr()), c_void_p(st[1].data_ptr()), c_void_p(st[2].data_ptr()),
           c_void_p(W0.data_ptr()), c_void_p(W1.data_ptr()), S, n, 0.25 / S, byref(pc), 1.0, 1, FAST, stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(3): run()
ts = sorted(run() for _ in range(7))
ms = ts[len(ts) // 2]
bytes_alg = 16.0 * n * S + 48.0 * n
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
print(f"logsv_step_fixed{'_fast' if FAST else ''}_kernel: {n} paths x {S} steps (W0+W1 = {2 * 8 * n * S / 1e9:.2f} GB > L2): median {ms:.3f} ms, "
      f"{n * S / ms / 1e6:.1f} Gpath-steps/s, algorithmic {bytes_alg / ms / 1e6:.0f} GB/s = {bytes_alg / ms / 1e6 / peak:.3f} of measured HBM peak {peak} GB/s")
