"""Hardware check of the one-sided peer gather (distributed.PeerGather) on REAL multi-GPU boxes -- not a pytest file: kernels
that wait for one another must not share a GPU, so this cannot run in the single-GPU test tier.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_check.py

Every rank runs STEPS batches of the AGS pipeline, several in flight on different streams, its compaction kernel storing
the kept rows straight into rank 0's buffer over NVLink (BT_PEER_TRANSPORT=direct) or its exchange stream pushing them there
from the local compaction output (BT_PEER_TRANSPORT=staged).  For EVERY step rank 0 compares what arrived (PeerGather.result)
with the ranks' locally computed rows sent through a plain NCCL gather, bit for bit.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

STEPS, STREAMS, N_REF = 14, 3, 2000


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    from burg_toolkit_b200 import distributed as bdist, mesh as bmesh, sampling
    from burg_toolkit_b200.gripper import ParallelJawGripper
    ags = sampling.AntipodalGraspSampler()
    ags.mesh = bmesh.create_bumpy_sphere(96, 96, radius=0.035, amplitude=0.15, seed=4)
    ags.gripper, ags.verbose, ags.device = ParallelJawGripper(), False, local
    peer = bdist.PeerGather(N_REF * 12, with_contacts=True, device=local, timeout_s=30.0)
    assert peer.one_sided
    streams = [torch.cuda.Stream() for _ in range(STREAMS)]
    ok, rows_seen = True, 0
    epochs = []
    for s in range(STEPS):
        k = s % STREAMS
        seed = 1000 * (rank + 1) + s
        with torch.cuda.stream(streams[k]):
            out = ags.run_pipeline(n_ref=N_REF, seed=seed, destinations=peer.destinations(slot=k), slot=k, graph=True)
            epochs.append((peer.complete(out['n_free'], rows=out.get('free_gs'), contacts=out.get('free_contacts'), slot=k), seed))
        # verify with a lag of one step, so that two steps are always in flight while rank 0 reads an older one
        if len(epochs) >= 2 or s == STEPS - 1:
            for e, sd in (epochs[:-1] if s < STEPS - 1 else epochs):
                got_rows, got_contacts = peer.result(e)
                local_out = ags.run_pipeline(n_ref=N_REF, seed=sd, slot='verify')
                n = int(local_out['n_free'].item())
                want_rows = bdist.gather_rows(local_out['free_gs'][:n].contiguous(), 0)
                want_contacts = bdist.gather_rows(local_out['free_contacts'][:n].contiguous(), 0)
                if rank == 0:
                    same = got_rows.shape == want_rows.shape and torch.equal(got_rows, want_rows) and torch.equal(got_contacts, want_contacts)
                    ok = ok and bool(same) and want_rows.shape[0] > 1000 * world
                    rows_seen += int(want_rows.shape[0])
            epochs = epochs[-1:] if s < STEPS - 1 else []
    timed_out = peer.close(raise_on_timeout=False)
    flag = torch.tensor([0.0 if (ok and not timed_out) else 1.0], device='cuda')
    dist.all_reduce(flag)
    if rank == 0:
        print(json.dumps({'check': 'PeerGather one-sided, multi-stream', 'transport': 'staged' if peer.staged else 'direct', 'world': world, 'steps': STEPS, 'streams': STREAMS,
                          'rows_compared': rows_seen, 'ok': bool(flag.item() == 0.0)}))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0.0 else 1)


if __name__ == '__main__':
    main()
