#!/usr/bin/env python3
"""Exchange micro-benchmark under torchrun: the copy-engine (or CTA, MLV_COPY_CTAS) exchange of the
inverse fields of a 16384^2 slab decomposition alone, no transforms: GB/s sent per GPU.
usage: torchrun --nproc-per-node N tools/dma_rate.py [nx nz reps]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from melvin.sharded import ShardedScalarStepper  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 2 else 16384
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
st = ShardedScalarStepper(nx, nz, 16.0 / 9.0, 1.0, 1e-6, 1e-6, mode="dma")
world, rank = st.world, st.rank
sent = 16 * st.inv_field * (world - 1)                       # bytes one field sends to its peers


def one_round():
    for f in range(3):
        st._dma(0, f, st._ev[f])
    st._dma_join(st._ev[3])


for _ in range(2):
    one_round()
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    one_round()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
t = torch.tensor([ms], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"DMA_RATE world={world} grid={nx}x{nz} copy_ctas={os.environ.get('MLV_COPY_CTAS', '0')} "
          f"streams={len(st._copy_streams)} block={16 * st.inv_field / 1e6:.1f}MB "
          f"round(3 fields)={float(t.item()):.3f} ms  -> {3 * sent / (float(t.item()) * 1e-3) / 1e9:.0f} GB/s sent per GPU")
st.close()
dist.destroy_process_group()
