import numpy as np, torch, ctypes as C
from deltakd_b200 import _lib
from deltakd_b200.functional import _ptr, _stream
rng = np.random.default_rng(0)
x = rng.standard_normal((20000, 384))
g = torch.tensor(np.stack([x.T @ x] * 3), device="cuda")
w = g.clone()
ws = torch.zeros(8192, dtype=torch.uint8, device="cuda")
sw = torch.zeros(3, dtype=torch.int32, device="cuda")
_lib.call("dkd_lrkd_eigensolve", _ptr(w), 3, 64, _ptr(sw), 1, _ptr(ws), ws.numel(), _stream())
torch.cuda.synchronize()
prof = ws[1024:].view(torch.int64).cpu().numpy()
print("sweeps", sw.tolist())
print("per macro round (cycles): load | step A | step B | store | syncthreads")
for ir in range(5):
    print(ir, prof[ir * 8: ir * 8 + 5])
print("cluster_sync (+remote stores drain) per group round:", prof[200:216])
