"""coverage_c5 shard timing on one GPU (scratch): how does the prepared-grid call scale with the reference shard?"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from burg_toolkit_b200 import metrics
from burg_toolkit_b200.distributed import spatial_order
from burg_toolkit_b200.synthetic import coverage_sets

ref_all, qry = coverage_sets(1_000_000)
dq = torch.from_numpy(qry).cuda()
pq = metrics.prepare_query(dq, 15.0)
order = spatial_order(ref_all)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for label, rows in (('all rows', ref_all), ('1/2 spatial', ref_all[order[:500_000]]), ('1/2 contiguous', ref_all[:500_000]),
                    ('1/8 spatial', ref_all[order[:125_000]]), ('1/8 spatial (mid)', ref_all[order[500_000:625_000]]), ('1/8 contiguous', ref_all[:125_000])):
    d = torch.from_numpy(np.ascontiguousarray(rows)).cuda()
    for _ in range(3):
        metrics.covered_mask(d, pq, return_device=True)
    ts = []
    for s in range(8):
        flush.fill_(s)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        metrics.covered_mask(d, pq, return_device=True)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f'{label:22s} {np.median(ts):.3f} ms')
