"""CPU check of the host side of the pinned staging ring (csrc/host_staging.hpp): the copy pool that moves pageable
caller memory into / out of the pinned blocks must be an exact memcpy for every size, alignment and thread count
(streaming-store body, memcpy head and tail, pieces handed to worker threads).  The GPU side of the ring is covered by
test_pageable_batches_go_through_the_pinned_ring."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_copy_pool_is_an_exact_memcpy():
    exe = os.path.join(HERE, "_build", "host_copy_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cuda_inc = "/usr/local/cuda/include"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", cuda_inc, "-o", exe,
                           os.path.join(HERE, "host", "host_copy_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 bad" in out.stdout
