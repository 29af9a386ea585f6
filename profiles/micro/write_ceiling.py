"""Times write_ceiling.cu (see there) for the rollout kernel's geometry and a few others; prints
µs per frame and TB/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared
-Xcompiler -fPIC -o profiles/micro/libwrite_ceiling.so profiles/micro/write_ceiling.cu"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libwrite_ceiling.so"))
lib.write_ceiling.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_int] * 7 + [ctypes.c_void_p]


def run(n_envs, envs_per_cta, block, store_warps, frames, ring=3, reps=20, interleaved=0):
    frame_bytes = n_envs * 1616
    grid = n_envs // envs_per_cta
    buf = torch.empty(ring * frame_bytes, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launch():
        rc = lib.write_ceiling(buf.data_ptr(), frame_bytes, grid, block, envs_per_cta * 1616, frames,
                               ring, store_warps, interleaved, st)
        assert rc == 0, rc
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        launch()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / (reps * frames)
    print("envs %8d  grid %6d x %3d threads, %d storing warps, %2d frames/launch%s: %7.2f us/frame  %.2f TB/s"
          % (n_envs, grid, block, store_warps, frames, ("", ", chunk-interleaved", ", frames innermost")[interleaved], us,
             frame_bytes / us / 1e6))


if __name__ == "__main__":
    # ring == frames: no line is rewritten while it may still be dirty in L2 (with a short ring a
    # short-lived CTA returns to the same lines within microseconds, L2 merges the writes and the
    # "bandwidth" exceeds what HBM can do: 8.4 TB/s was seen that way)
    for epc, block, sw in ((64, 128, 2), (32, 64, 1), (32, 32, 1), (16, 32, 1), (8, 32, 1), (4, 32, 1),
                           (16, 128, 4), (64, 256, 8), (128, 256, 4)):
        run(65536, epc, block, sw, 8, ring=8)
    run(65536, 64, 128, 2, 8, ring=8, interleaved=1)
    run(65536, 64, 128, 2, 8, ring=8, interleaved=2)
    run(65536, 64, 128, 2, 40, ring=40)
    for epc, block, sw in ((64, 128, 2), (32, 64, 1), (16, 32, 1), (4, 32, 1)):
        run(1048576, epc, block, sw, 8, ring=8, reps=5)
