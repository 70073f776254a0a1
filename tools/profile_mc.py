"""Host-side cost of montecarlo(x, bins=100) on a resident 1e8 sample: cProfile of 30 calls."""
import cProfile
import os as _os
import pstats
import sys
sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import torch  # noqa: E402
import sdepy_b200 as sd  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
x = torch.empty(n, dtype=torch.float64, device='cuda').normal_()
for _ in range(3):
    sd.montecarlo(x, bins=100)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    sd.montecarlo(x, bins=100)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
