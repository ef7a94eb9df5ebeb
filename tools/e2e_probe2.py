"""where does the end-to-end call spend its time?  (scratch; run on a GPU box)"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from burg_toolkit_b200 import sampling
from burg_toolkit_b200.gripper import ParallelJawGripper
from burg_toolkit_b200.synthetic import c3_mesh

ags = sampling.AntipodalGraspSampler()
ags.mesh, ags.gripper, ags.verbose = c3_mesh(), ParallelJawGripper(), False
N = 10000
batches = [[t.cpu().numpy() for t in ags.sample_surface(N, 7 + b)[:2]] for b in range(4)]
K = 40


def timed(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print(f'{label:58s} {(time.perf_counter() - t0) / K * 1e3:.3f} ms per batch')


def serial():
    for s in range(K):
        ags.evaluate_host(*batches[s % 4], seed=s + 1)


def pipelined(slots=2):
    def f():
        pend = []
        for s in range(K):
            pend.append(ags.submit_host(*batches[s % 4], seed=s + 1, slot=s % slots))
            if len(pend) >= slots:
                pend.pop(0).result()
        for p in pend:
            p.result()
    return f


def submit_only():
    # host cost of a submission: never wait (two slots would be overwritten, so use K slots' worth of time only)
    t0 = time.perf_counter()
    pend = [ags.submit_host(*batches[s % 4], seed=s + 1, slot=s % 2) for s in range(K)]
    host = time.perf_counter() - t0
    for p in pend:
        p.result()
    print(f'{"host time of submit_host (no waiting)":58s} {host / K * 1e3:.3f} ms per batch')


def device_only(streams=1):
    sts = [torch.cuda.Stream() for _ in range(streams)]

    def f():
        for s in range(K):
            with torch.cuda.stream(sts[s % streams]):
                ags.run_pipeline(n_ref=N, seed=s + 1, graph=True, slot=('dev', s % streams))
    return f


timed(serial, 'evaluate_host (serial)')
timed(pipelined(2), 'submit_host, 2 slots')
timed(pipelined(3), 'submit_host, 3 slots')
submit_only()
timed(device_only(1), 'device-resident graph replays, 1 stream')
timed(device_only(2), 'device-resident graph replays, 2 streams')
timed(device_only(3), 'device-resident graph replays, 3 streams')
