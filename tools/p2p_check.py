#!/usr/bin/env python
"""Multi-GPU check of the fused Gram slab exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/p2p_check.py

Every rank builds the same synthetic X, the sharded Gram matrix is produced once with the peer-memory
exchange and once with NCCL broadcasts, and both must equal the single-GPU result bit for bit.  Prints one
line per stage (flushed) so that a hang can be located.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
from rtrec_b200.utils.synth import synth_events
from rtrec_b200 import device as D, pipeline as P

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
T0 = time.time()


def say(msg):
    print(f"[{time.time() - T0:6.2f}s rank {rank}] {msg}", file=sys.stderr, flush=True)


torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
say("process group up")
U, I, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4000, 3001, 300000)))
u, i, ts, r = synth_events(U, I, N, seed=21, rating="int")
X = D.DeviceMatrix.from_scipy(sp.csc_matrix((r.astype(np.float32), (u, i)), shape=(U, I)))
G1 = D.gram_full(X)
torch.cuda.synchronize()
say("single-GPU gram done")
for ex in ("nccl", "p2p", "p2p"):
    G = P.gram_sharded(X, rank=rank, world=world, exchange=ex, marks=lambda name: say(f"  {ex}: {name} queued"))
    torch.cuda.synchronize()
    same = bool(torch.equal(G, G1))
    say(f"{ex}: equal to single-GPU result = {same}")
    assert same
slabs = P.PeerSlabs.get(I, rank, world)
say(f"peer slabs: {'mapped' if slabs is not None else 'unavailable (NCCL fallback)'}")
dist.barrier()
say("done")
dist.destroy_process_group()
