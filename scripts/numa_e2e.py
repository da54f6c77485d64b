"""GPU box, torchrun with N ranks: does binding every rank to its GPU's NUMA node lift the host-bound e2e rate?
Phase A: ranks run wherever the scheduler puts them; phase B: hostbind.bind_host_to_gpu before the pinned arrays are
allocated.  Per phase: pinned-copy rate with every rank copying, and quantize_batch on 512 pinned proteins."""
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch
from dctdomain_b200.hostbind import bind_host_to_gpu, gpu_numa_cpus

rank = int(os.environ.get('RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather_floats(vals):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world == 1:
        return [t.tolist()]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


if rank == 0:
    for cmd in (['nvidia-smi', 'topo', '-m'], ['lscpu']):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            if cmd[0] == 'lscpu':
                txt = '\n'.join(l for l in txt.splitlines() if any(k in l for k in ('Model name', 'Socket', 'NUMA', 'CPU(s):', 'Thread')))
            print(txt, flush=True)
        except Exception as e:     # noqa: BLE001
            print(cmd, 'failed:', e)
    print('affinity at start:', len(os.sched_getaffinity(0)), 'cpus', flush=True)
node, cpus = gpu_numa_cpus(local)
print(f'rank {rank}: gpu numa node {node}, {None if cpus is None else len(cpus)} cpus there', flush=True)

D, B = 1280, 512
lens = np.random.RandomState(777 + rank).randint(40, 501, size=B)
nbytes = int(sum(2 * int(L) * D * 4 for L in lens))


def phase(tag):
    gen = torch.Generator().manual_seed(99 + rank)
    host = [(f'p{i}', int(L), {15: torch.randn(int(L), D, generator=gen).pin_memory(),
                               21: torch.randn(int(L), D, generator=gen).pin_memory()}) for i, L in enumerate(lens)]
    big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    dbig = torch.empty_like(big, device=dev)
    dbig.copy_(big, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        dbig.copy_(big, non_blocking=True)
    torch.cuda.synchronize()
    link = 8 * big.numel() / (time.perf_counter() - t0) / 1e9

    def one():
        fps = [Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{Ln}'], quants={}) for pid, Ln, emb in host]
        quantize_batch(fps, [3, 80, 3, 80], device=dev)
    for _ in range(2):
        one()
    barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        one()
    torch.cuda.synchronize()
    mine = 6 * nbytes / (time.perf_counter() - t0) / 1e9
    barrier()
    rows = gather_floats([link, mine])
    if rank == 0:
        print(f'[{tag}] pinned copy, all ranks copying, GB/s per rank: ' + ' '.join(f'{r[0]:.1f}' for r in rows)
              + f' | sum {sum(r[0] for r in rows):.0f}')
        print(f'[{tag}] quantize_batch e2e, GB/s per rank:             ' + ' '.join(f'{r[1]:.1f}' for r in rows)
              + f' | sum {sum(r[1] for r in rows):.0f} = {sum(r[1] for r in rows) * 1e9 / (nbytes / B):.0f} fingerprints/s', flush=True)
    del host, big, dbig


phase('unbound')
torch._C._host_emptyCache()
info = bind_host_to_gpu(local)
print(f'rank {rank}: bound to node {info["numa_node"]}, {None if info["cpus"] is None else len(info["cpus"])} cpus '
      f'(of {len(info["previous"])}), nodes on the host: {info["n_nodes"]}', flush=True)
phase('bound')
if world > 1:
    dist.destroy_process_group()
