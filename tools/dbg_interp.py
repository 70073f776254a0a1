import numpy as np, torch
import sdepy_b200 as m
from tests.cases import golden
g, s = golden('stats_cdf_chf'), golden('stats_lognorm')
dp = m.device_process(s['t'], torch.from_numpy(s['x']).cuda())
got = dp(g['tq']).cpu().numpy()
d = np.abs(got - g['interp'])
for k, t in enumerate(g['tq']):
    print(t, d[k].max(), (d[k] > 0).sum(), dp._bracket(float(t)))
hp = m.process(s['t'], x=s['x'])
print('host', np.abs(hp(g['tq']) - g['interp']).max())
