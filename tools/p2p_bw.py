#!/usr/bin/env python
"""Peer-to-peer bandwidth between GPU 0 and GPU 1 of this box as seen by one process: cudaMemcpyPeer (torch
copy_) in both directions, and nvidia-smi's NVLink topology.  Development tool for the slab exchange."""
import subprocess, sys, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
n = torch.cuda.device_count()
print("devices", n, "peer access 0->1:", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
if n < 2:
    sys.exit(0)
nbytes = 1 << 30
a = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
b = torch.empty(nbytes, dtype=torch.uint8, device="cuda:1")
for name, dst, src in (("1->0", a, b), ("0->1", b, a)):
    for _ in range(2):
        dst.copy_(src)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(dst.device):
        e0.record()
        for _ in range(5):
            dst.copy_(src)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"memcpy peer {name}: {nbytes / ms / 1e6:.1f} GB/s ({ms:.3f} ms per GiB)")
