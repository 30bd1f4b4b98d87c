"""Which transport does NCCL pick between two GPUs of this box, and how fast is a 2 GB send/recv?"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
x = torch.zeros(256 * 1024 * 1024, dtype=torch.float64, device="cuda")   # 2 GiB
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    if rank == 0: dist.recv(x, src=1)
    else: dist.send(x, dst=0)
    torch.cuda.synchronize(); dt = time.time() - t0
    if rank == 0: print("send/recv 2 GiB: %.1f ms = %.1f GB/s" % (dt * 1e3, x.numel() * 8 / dt / 1e9), flush=True)
y = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); dist.barrier(); t0 = time.time(); dist.all_reduce(y); torch.cuda.synchronize()
if rank == 0: print("all_reduce 512 MiB: %.1f ms" % ((time.time() - t0) * 1e3))
dist.destroy_process_group()
