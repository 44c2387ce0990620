"""The e2m1 (tcgen05.mma.kind::mxf4) forward Gram pass against the int8 pass on one B200: Hamming histograms equal at a few
shapes, then both timed at cfg3 (8192 + 8192 rows, D = 5640; packing and histogram zeroing included).  SMALL=1 stops
after the small shapes (compute-sanitizer runs); B200GRBM_LIB selects an experiment build of the library."""
import os, sys, time
sys.path.insert(0, ".")
import torch
from image_generation_b200 import _lib, mmd_tc
if os.environ.get("B200GRBM_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["B200GRBM_LIB"])
dev = torch.device("cuda:0")
def hist(z, m_x, d, fp4, shard=(0,1)):
    os.environ["B200GRBM_MMD_FP4"] = "1" if fp4 else "0"
    return mmd_tc.mmd_histograms_i8(z, m_x, d, shard)
g = torch.Generator(device=dev).manual_seed(0)
for (m_x, m_y, d) in [(128, 256, 256), (300, 200, 77), (513, 640, 900), (1024, 256, 256)] + ([] if os.environ.get("SMALL") else [(2048, 2048, 5640)]):
    x = (torch.randint(0, 2, (m_x + m_y, d), generator=g, device=dev) * 2 - 1).to(torch.int8)
    zi, _ = mmd_tc.pack_rows_i8(x)
    a = hist(zi, m_x, d, False); b = hist(zi, m_x, d, True)
    torch.cuda.synchronize()
    print((m_x, m_y, d), "equal" if torch.equal(a, b) else f"DIFF {int((a-b).abs().sum())} of {int(a.sum())}", flush=True)
if os.environ.get("SMALL"):
    sys.exit(0)
m_x = m_y = 8192; d = 5640
x = (torch.randint(0, 2, (m_x + m_y, d), generator=g, device=dev) * 2 - 1).to(torch.int8)
zi, _ = mmd_tc.pack_rows_i8(x)
for fp4 in (False, True):
    os.environ["B200GRBM_MMD_FP4"] = "1" if fp4 else "0"
    z4 = mmd_tc.pack_fp4(zi) if fp4 else None
    h = torch.zeros((3, d + 1), dtype=torch.int64, device=dev)
    for it in range(3):
        hist(zi, m_x, d, fp4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        hh = hist(zi, m_x, d, fp4)
    e1.record(); torch.cuda.synchronize()
    print("fp4" if fp4 else "int8", "cfg3 hist incl. pack/zero:", e0.elapsed_time(e1) / 5, "ms")
    if fp4: print("equal to int8:", torch.equal(hh, ref))
    else: ref = hh
