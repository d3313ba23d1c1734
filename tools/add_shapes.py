"""Which tensors do the strided gradient-accumulation adds of the training step touch?  (torch.profiler with shapes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from torch.profiler import profile, ProfilerActivity
step, info = bench.make_step('kth_train_b32')
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    step(); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ('aten::add', 'aten::add_', 'aten::sum', 'aten::copy_', 'aten::cat', 'aten::threshold_backward', 'aten::clamp_min', 'aten::relu')]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:28]:
    print('%-24s n=%4d dev=%8.1f us  %s' % (e.key, e.count, e.device_time_total, str(e.input_shapes)[:110]))
