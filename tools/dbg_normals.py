import numpy as np, torch
from sdepy_b200 import _lib, _cuda
n = 1 << 20
zf = torch.empty(2*n, dtype=torch.float64, device='cuda'); zl = torch.empty_like(zf)
_lib.check(_lib.lib.sdeb_test_normals(1234, n, _cuda.ptr(zf), _cuda.ptr(zl), _cuda.stream_ptr(zf.device)))
zf, zl = zf.cpu().numpy(), zl.cpu().numpy()
d = np.abs(zf - zl)
idx = np.argsort(-np.nan_to_num(d, nan=1e9))[:12]
for i in idx:
    print(i, i//2, zf[i], zl[i], d[i], 'r=', np.hypot(zl[2*(i//2)], zl[2*(i//2)+1]))
print('nan fast', np.isnan(zf).sum(), 'nan lib', np.isnan(zl).sum())
print('max err beyond corners', d[128:].max())
