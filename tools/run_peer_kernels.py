"""cfg3-sized bit-row pack / expand launches on one GPU (csrc/peer_exchange.cu) for ncu and compute-sanitizer:
8192 fp32 + 8192 int8 rows x 5640 spins -> bit rows -> the int8 Gram operand, checked against the spin extraction."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from image_generation_b200.dist import _DeviceOps as ops
from image_generation_b200.mmd_tc import pack_pair_i8

dev = torch.device("cuda:0")
rows, d = int(os.environ.get("ROWS", 8192)), int(os.environ.get("D", 5640))
g = torch.Generator(device=dev).manual_seed(1)
x = (torch.randint(0, 2, (rows, d), generator=g, device=dev) * 2 - 1).float()
y = (torch.randint(0, 2, (rows, d), generator=g, device=dev) * 2 - 1).to(torch.int8)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    ev[0].record()
    bits = ops.pack_bits(x, y)
    ev[1].record()
    z = ops.unpack_bits(bits.unsqueeze(0), rows, rows, d)
    ev[2].record()
torch.cuda.synchronize()
ok = torch.equal(z, pack_pair_i8(x, y).rows)
in_bytes, out_bytes = x.numel() * 4 + y.numel(), z.numel()
print(f"pack {ev[0].elapsed_time(ev[1]):.4f} ms ({in_bytes / ev[0].elapsed_time(ev[1]) / 1e6:.0f} GB/s read), "
      f"expand {ev[1].elapsed_time(ev[2]):.4f} ms ({out_bytes / ev[1].elapsed_time(ev[2]) / 1e6:.0f} GB/s written), equal={ok}")
sys.exit(0 if ok else 1)
