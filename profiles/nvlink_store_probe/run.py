"""Experiment (one process, two GPUs): SM-driven stores into a peer's memory over NVLink, by store
pattern (see probe.cu). Prints GB/s for a local and for a peer destination."""
import ctypes
from pathlib import Path

import torch

lib = ctypes.CDLL(str(Path(__file__).with_name("libprobe.so")))
lib.nvlink_store_probe.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64,
                                   ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
size = 1 << 29
local = torch.empty(size // 4, dtype=torch.int32, device="cuda:0")
peer = torch.empty(size // 4, dtype=torch.int32, device="cuda:1")
peer.copy_(local)            # (torch enables peer access between the two devices here)
torch.cuda.synchronize(0)
torch.cuda.synchronize(1)
torch.cuda.set_device(0)
names = {0: "4 B/lane, rows 4 KiB apart (y-pass pattern)", 1: "4 B/lane, contiguous",
         2: "16 B/lane, contiguous", 3: "16 B/lane, 512 B pieces 4 KiB apart"}
for target_name, target in (("local", local), ("peer", peer)):
    for pattern in (0, 1, 2, 3):
        for blocks in (148 * 4, 148 * 8):
            ms = ctypes.c_float()
            code = lib.nvlink_store_probe(pattern, ctypes.c_void_p(target.data_ptr()), size, 4096,
                                          blocks, 5, ctypes.byref(ms))
            print(f"{target_name:5s} blocks {blocks:5d} {names[pattern]:45s} "
                  f"{size / ms.value / 1e6:8.1f} GB/s (code {code})", flush=True)
