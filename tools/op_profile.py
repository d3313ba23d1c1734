import sys, os
sys.path.insert(0, '/root/repo')
import torch, bench
from torch.profiler import profile, ProfilerActivity
step, info = bench.make_step('kth_train_b32')
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted(ka, key=lambda e: -getattr(e, 'self_device_time_total', 0))
tot = sum(getattr(e, 'self_device_time_total', 0) for e in ka)
for e in rows[:45]:
    print('%-60s n=%5d self_dev=%9.1f us  %.3f' % (e.key[:60], e.count, e.self_device_time_total, e.self_device_time_total / tot))
