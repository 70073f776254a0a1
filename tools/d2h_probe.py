import time, torch
x = torch.empty((501, 1_000_000), dtype=torch.float64, device='cuda').normal_()
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); a = x.cpu().numpy(); t1 = time.perf_counter()
    print('pageable .cpu(): %.3f s  %.1f GB/s' % (t1 - t0, x.numel()*8/(t1 - t0)/1e9))
for rep in range(3):
    t0 = time.perf_counter()
    h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    t1 = time.perf_counter()
    h.copy_(x, non_blocking=True); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('pinned alloc %.3f s, copy %.3f s (%.1f GB/s)' % (t1 - t0, t2 - t1, x.numel()*8/(t2 - t1)/1e9))
    del h
