"""H2D probe (development aid): is cudaMemcpyAsync from gg_host_alloc memory asynchronous, at what rate, and do two
streams add bandwidth?  Measured on the B200 box: call 5 us, 55 GB/s, no gain from a second stream (pageable: 13 GB/s,
blocking) -- the 63 MB upload of the e2e leg (1.15 ms + packing) is at link speed."""
import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
from cuda import cudart
from gasoline_b200.pkd import pinned_empty, load_library
load_library()
n = 8_000_000
err, d = cudart.cudaMalloc(n * 8)
err, st = cudart.cudaStreamCreateWithFlags(cudart.cudaStreamNonBlocking)
err, st2 = cudart.cudaStreamCreateWithFlags(cudart.cudaStreamNonBlocking)
err, d2 = cudart.cudaMalloc(n * 8)
for name, arr in (("gg_host_alloc", pinned_empty(n)), ("pageable", np.zeros(n))):
    arr[:] = 1.0
    for rep in range(3):
        cudart.cudaStreamSynchronize(st)
        t0 = time.perf_counter()
        cudart.cudaMemcpyAsync(d, arr.ctypes.data, n * 8, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
        t1 = time.perf_counter()
        cudart.cudaStreamSynchronize(st)
        t2 = time.perf_counter()
    print(f"{name}: call {1e3*(t1-t0):.3f} ms, sync {1e3*(t2-t1):.3f} ms, {n*8/(t2-t0)/1e9:.1f} GB/s")
a, b = pinned_empty(n), pinned_empty(n)
a[:] = 1; b[:] = 2
for rep in range(3):
    cudart.cudaDeviceSynchronize()
    t0 = time.perf_counter()
    cudart.cudaMemcpyAsync(d, a.ctypes.data, n * 8, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
    cudart.cudaMemcpyAsync(d2, b.ctypes.data, n * 8, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st2)
    t1 = time.perf_counter()
    cudart.cudaDeviceSynchronize()
    t2 = time.perf_counter()
print(f"two streams: calls {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms, {2*n*8/(t2-t0)/1e9:.1f} GB/s")
