"""Peer-to-peer copy bandwidth between device 0 and 1 (torch.copy_, which enables peer access) -- context for the in-process
multi-GPU numbers: python scripts/p2p_probe.py"""
import subprocess, time, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("can access peer 0->1:", torch.cuda.can_device_access_peer(0, 1))
for mb in (32, 128, 1024, 4096):
    x = torch.empty(mb << 20, dtype=torch.uint8, device="cuda:0")
    y = torch.empty(mb << 20, dtype=torch.uint8, device="cuda:1")
    for _ in range(2): y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(5): y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 5
    print(f"{mb:5d} MiB 0->1: {dt*1e3:8.3f} ms  {mb/1024/dt:7.1f} GiB/s")
