"""Host -> device copy bandwidth of a 264 MB pinned buffer (the e2e step's input): regular pinned memory vs
write-combined pinned memory (cudaHostAllocWriteCombined), one copy vs two concurrent half copies."""
import ctypes
import sys

import torch

n_bytes = 264 * 1000 * 1000
dev = torch.device("cuda", 0)
dst = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
rt = ctypes.CDLL("libcudart.so.12")


def time_copy(src_ptr, label, split=1):
    streams = [torch.cuda.Stream() for _ in range(split)]
    part = n_bytes // split
    def once():
        for k, s in enumerate(streams):
            rt.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + k * part), ctypes.c_void_p(src_ptr + k * part),
                               ctypes.c_size_t(part), 1, ctypes.c_void_p(s.cuda_stream))
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(10):
        once()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label:40s} {n_bytes / ms / 1e6:6.1f} GB/s  ({ms:.2f} ms)")


for flags, name in ((0, "pinned (cudaHostAllocDefault)"), (4, "pinned write-combined")):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n_bytes), ctypes.c_uint(flags))
    if rc != 0:
        print(name, "cudaHostAlloc failed", rc)
        continue
    ctypes.memset(p, 1, n_bytes)
    time_copy(p.value, name)
    time_copy(p.value, name + ", 2 concurrent halves", split=2)
    rt.cudaFreeHost(p)
